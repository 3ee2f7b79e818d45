"""ncu per-launch DRAM bytes (profiles/r02_traffic.sh) -> per-kernel bytes per engine pass, written to
gpurun_out/r02_traffic.json (copy it to profiles/r02_traffic.json to have bench.py report it)."""
import collections, csv, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}
K1 = ("simplsKspaceKernel", "simplsFitFast", "simplsFitKernel", "simplsFitTmem", "simplsScoreKernel", "gramPackKernel", "gramPackVectorsKernel", "predictBlocksKernel")
out = {}
for wl in sys.argv[1:]:
    rows = [r for r in csv.reader(open("gpurun_out/r02_traffic_%s.csv" % wl)) if len(r) > 5 and r[0].isdigit()]
    per = collections.OrderedDict()
    passes = 0
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").replace("gsb::<unnamed>::", "").split("<")[0]
        metric, unit, val = r[-3], r[-2], float(r[-1].replace(",", ""))
        d = per.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "ns": 0.0})
        if metric == "gpu__time_duration.sum":
            d["ns"] += val * UNIT.get(unit, 1.0)
            d["launches"] += 1
            passes += name == "cvReduceKernel"
        else:
            d["dram_bytes"] += val * UNIT.get(unit, 1.0)
    passes = max(passes, 1)
    k1 = sum(d["dram_bytes"] for n, d in per.items() if n in K1) / passes
    pop = bench.WORKLOADS[wl]["pop"]
    out["%s/%d" % (wl, pop)] = {
        "k1_dram_bytes_per_step": k1, "engine_passes_in_capture": passes,
        "source": "profiles/r02_traffic.json: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over `python bench.py --workload %s "
                  "--steps 1 --warmup 3`, K1 launches summed, / %d engine passes (L2 flushed before every pass)" % (wl, passes),
        "per_kernel_per_step": {n: {"launches": d["launches"] / passes, "dram_MB": d["dram_bytes"] / passes / 1e6, "ms": d["ns"] / passes / 1e6}
                                for n, d in per.items() if d["launches"]}}
    print(wl, "K1 DRAM MB/step %.1f" % (k1 / 1e6))
    for n, d in out["%s/%d" % (wl, pop)]["per_kernel_per_step"].items():
        print("   %-28s %6.1f launches  %9.1f MB  %8.3f ms" % (n, d["launches"], d["dram_MB"], d["ms"]))
json.dump(out, open("gpurun_out/r02_traffic.json", "w"), indent=1)
