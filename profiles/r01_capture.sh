#!/bin/bash
# One gpurun call: bench line, reference arm, ncu launch list, ncu --set full of the K1 launches of one generation.
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2>> gpurun_out/r01_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r01_bench_plain.json 2>> gpurun_out/r01_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_l.log 2>&1
# K1 launches of one generation: the kernel-space kernel + 4 per-fit size classes per step; skip the first 4 generations
ncu --set full --clock-control none --import-source on -k regex:simpls -s 20 -c 5 -o gpurun_out/r01_k1_full -f \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
cat gpurun_out/r01_bench.json
