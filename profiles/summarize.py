"""Turn the ncu outputs of a round (gpurun_out/) into the small, committed summaries under profiles/.

usage: python profiles/summarize.py r01 ["the profiled command"]      (round 2: r02_cfg3, r02_cfg4 -- profiles/r02_capture.sh)
  gpurun_out/<round>_launches.csv      (ncu --metrics gpu__time_duration.sum ... python bench.py ...)
  gpurun_out/<round>_k1_full.ncu-rep   (ncu --set full ... -k regex:simpls ... python bench.py ...)
"""
import collections, csv, io, os, subprocess, sys

rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
command = sys.argv[2] if len(sys.argv) > 2 else "python bench.py --steps 3 --warmup 3 --no-cpu"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = []

path = os.path.join(root, "gpurun_out", rnd + "_launches.csv")
if os.path.exists(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(io.StringIO("\n".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        name = row["Kernel Name"].split("(")[0].replace("void gsb::<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out.append("== launch list (ncu --metrics gpu__time_duration.sum, cold-cache serialised launches; compare SHARES) ==")
    out.append("command: %s" % command)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-44s launches=%4d  total=%9.3f ms  share=%5.1f%%" % (k[:44], v[0], v[1], 100 * v[1] / tot))
    with open(os.path.join(root, "profiles", rnd + "_launches.csv"), "w") as fh:
        fh.write("\n".join(lines) + "\n")

rep = os.path.join(root, "gpurun_out", rnd + "_k1_full.ncu-rep")
rawcsv = os.path.join(root, "gpurun_out", rnd + "_k1_raw.csv")  # exported on the GPU box (the report itself stays there)
if os.path.exists(rep) or os.path.exists(rawcsv):
    raw = open(rawcsv).read() if os.path.exists(rawcsv) else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    raw = "\n".join(l for l in raw.splitlines() if l.startswith('"'))
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "lts__t_sectors_op_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(os.path.join(root, "profiles", rnd + "_k1_metrics.csv"), "w") as fh:
        w = csv.writer(fh)
        w.writerow([n for n, _ in idx] + ["unit row: " + " | ".join(rows[1][i] for _, i in idx)])
        for r in rows[2:]:
            w.writerow([r[i].replace("void gsb::<unnamed>::", "") for _, i in idx])
    out.append("")
    out.append("== K1 launches of one generation (kernel-space kernel, per-fit path per size class), ncu --set full --clock-control none ==")
    out.append("%-34s %-10s %-8s %9s %10s %10s %6s %7s %7s %7s %7s %7s %7s" % ("kernel", "block", "grid", "time ms", "dram rd MB", "dram wr MB", "regs", "smem KB", "warps%", "issue%", "fp64%", "lds%", "dmma%"))
    g = dict(idx)
    unit_of = {n: rows[1][i] for n, i in idx}
    t_ms = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit_of.get("gpu__time_duration.sum", "ms"), 1.0)
    b_mb = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    rd_mb = b_mb.get(unit_of.get("dram__bytes_read.sum", "Mbyte"), 1.0)
    wr_mb = b_mb.get(unit_of.get("dram__bytes_write.sum", "Mbyte"), 1.0)
    sm_kb = {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}.get(unit_of.get("launch__shared_mem_per_block_dynamic", "Kbyte"), 1.0)
    tot_t = tot_rd = tot_wr = 0.0
    for r in rows[2:]:
        def f(name, d=0.0):
            try:
                return float(r[g[name]].replace(",", ""))
            except (KeyError, ValueError):
                return d
        tot_t += f("gpu__time_duration.sum") * t_ms; tot_rd += f("dram__bytes_read.sum") * rd_mb; tot_wr += f("dram__bytes_write.sum") * wr_mb
        out.append("%-34s %-10s %-8s %9.3f %10.1f %10.1f %6.0f %7.1f %7.1f %7.1f %7.1f %7.1f %7.1f" % (
            r[g["Kernel Name"]].replace("void gsb::<unnamed>::", "").split("(")[0][:34], r[g["Block Size"]].replace(" ", ""), r[g["Grid Size"]].split(",")[0].strip("("), f("gpu__time_duration.sum") * t_ms,
            f("dram__bytes_read.sum") * rd_mb, f("dram__bytes_write.sum") * wr_mb, f("launch__registers_per_thread"),
            f("launch__shared_mem_per_block_dynamic") * sm_kb, f("sm__warps_active.avg.pct_of_peak_sustained_active"),
            f("smsp__issue_active.avg.pct_of_peak_sustained_active"), f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
            f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")))
    out.append("sum over the captured launches: time %.3f ms (serialised under ncu), dram read %.1f MB, dram write %.1f MB" % (tot_t, tot_rd, tot_wr))

text = "\n".join(out) + "\n"
with open(os.path.join(root, "profiles", rnd + "_summary.txt"), "w") as fh:
    fh.write(text)
print(text)
