#!/bin/bash
# round 2, first call: FP64 peaks (BASELINE.md section 3 protocol), DMMA/DFMA microbenchmark, cfg4 baseline + phase stamps of the per-fit kernel
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_box.txt; nproc >> gpurun_out/r02_box.txt
python profiles/fp64_peak.py > gpurun_out/r02_fp64_peak.json 2> gpurun_out/r02_fp64_peak.err
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_probe profiles/micro/dmma_probe.cu && /tmp/dmma_probe > gpurun_out/r02_dmma_probe_b200.txt 2>&1
for k in 24 40 64 110 160 200; do GSB_TRACE=1 python profiles/k1_probe4.py $k 148; done > gpurun_out/r02_k1_probe4.txt 2>&1
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_cfg4_base.json 2> gpurun_out/r02_cfg4_base.err
cat gpurun_out/r02_fp64_peak.json gpurun_out/r02_k1_probe4.txt gpurun_out/r02_cfg4_base.json
