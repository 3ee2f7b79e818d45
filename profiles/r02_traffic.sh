#!/bin/bash
# DRAM bytes and durations of every launch of one bench command (one cheap ncu pass, no replay of the full set), summed per
# engine pass -> profiles/r02_traffic.json (read by bench.py: roofline.traffic).  usage: r02_traffic.sh WORKLOAD [WORKLOAD ...]
for wl in "$@"; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 4000 --csv \
      --log-file gpurun_out/r02_traffic_${wl}.csv python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_traffic_${wl}.log 2>&1
done
python profiles/traffic_summary.py "$@"
