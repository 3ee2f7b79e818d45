#!/bin/bash
# ncu launch list (gpu__time_duration per launch, serialised, cold caches) of one bench workload; usage: r02_launches.sh WORKLOAD TAG
wl=${1:-cfg4}; tag=${2:-$wl}
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_${tag}_launches.csv python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_${tag}_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_${tag}_launches.csv
