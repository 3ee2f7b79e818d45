#!/bin/bash
# round 2, final build: per workload the ncu launch list of a bench command that has just exited 0 without ncu, the
# selected raw metrics of the K1 launches of ONE generation, one ncu --set full capture (with source) of the dominant
# kernel of each workload, and the DRAM-traffic pass bench.py reads (roofline.traffic).  The multi-launch reports are
# exported to CSV on the box and deleted (gpurun brings back at most 64 MiB).
# Summaries: python profiles/summarize.py r02_cfg3 / r02_cfg4 (here, after the call) -> profiles/r02_cfg*_summary.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum
for wl in cfg3 cfg4; do
  python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_${wl}_plain.json 2> gpurun_out/r02_${wl}_plain.err || { echo "$wl: plain run failed"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_${wl}_launches.csv \
      python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_${wl}_ncu_l.log 2>&1
done
# one generation of K1 launches, skipping the first four generations: cfg3 = K1k + 4 per-fit classes (pack, vectors, fit);
# cfg4 = 12 classes x (pack, vectors, fit, predict)
ncu --metrics $M --clock-control none -k regex:'simpls|gramPack|predictBlocks' -s 52 -c 13 -o /tmp/r02_cfg3_k1 -f \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_cfg3_ncu_f.log 2>&1; tail -1 gpurun_out/r02_cfg3_ncu_f.log
ncu -i /tmp/r02_cfg3_k1.ncu-rep --page raw --csv > gpurun_out/r02_cfg3_k1_raw.csv
ncu --metrics $M --clock-control none -k regex:'simpls|gramPack|predictBlocks' -s 192 -c 48 -o /tmp/r02_cfg4_k1 -f \
    python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_cfg4_ncu_f.log 2>&1; tail -1 gpurun_out/r02_cfg4_ncu_f.log
ncu -i /tmp/r02_cfg4_k1.ncu-rep --page raw --csv > gpurun_out/r02_cfg4_k1_raw.csv
# the dominant kernel of each workload with everything (source attached): K1k, and the fit kernel's largest class at config 4
ncu --set full --clock-control none --import-source on -k regex:simplsKspace -s 4 -c 1 -o gpurun_out/r02_k1k_full -f \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_k1k_ncu.log 2>&1; tail -1 gpurun_out/r02_k1k_ncu.log
ncu --set full --clock-control none --import-source on -k regex:simplsFitFast -s 48 -c 1 -o gpurun_out/r02_fit4_full -f \
    python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_fit4_ncu.log 2>&1; tail -1 gpurun_out/r02_fit4_ncu.log
bash profiles/r02_traffic.sh cfg3 cfg4
ls -la gpurun_out/*.ncu-rep | tail -4; du -sm gpurun_out
