"""Key numbers of ncu --set full reports (one kernel each): duration, launch shape, pipe activity, shared-memory wavefronts,
DRAM bytes, L2 hit rate and the warp-stall shares.   usage: python profiles/ncu_extract.py REPORT.ncu-rep [...] > profiles/r02_ncu_full.txt"""
import csv, io, subprocess, sys
NAMES = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
         "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
         "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO("\n".join(l for l in raw.splitlines() if l.startswith('"')))))
    hdr, unit = rows[0], rows[1]
    for r in rows[2:]:
        print("== %s: %s  block %s grid %s" % (rep.split("/")[-1], r[hdr.index("Kernel Name")], r[hdr.index("Block Size")], r[hdr.index("Grid Size")]))
        for n in NAMES:
            if n in hdr:
                print("   %-82s %14s %s" % (n, r[hdr.index(n)], unit[hdr.index(n)]))
        st = [(h, float(r[i] or 0)) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        tot = sum(v for _, v in st) or 1.0
        print("   warp stalls (share of issue-stall samples): " + ", ".join("%s %.1f %%" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), 100 * v / tot)
                                                                          for h, v in sorted(st, key=lambda x: -x[1])[:8]))
