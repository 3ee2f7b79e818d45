#!/bin/bash
# round 2, last build: r02_final.sh (tests, smoke, every workload's bench line, reference arm, sweep, GA, probes, the raw
# metrics of one config-4 generation), then the config-4 launch list and the DRAM-traffic pass of both workloads.
bash profiles/r02_final.sh
python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_cfg4_plain.json 2> gpurun_out/r02_cfg4_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_cfg4_launches.csv \
    python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_cfg4_ncu_l.log 2>&1
python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_cfg3_plain.json 2> gpurun_out/r02_cfg3_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_cfg3_launches.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_cfg3_ncu_l.log 2>&1
bash profiles/r02_traffic.sh cfg3 cfg4
du -sm gpurun_out
