#!/bin/bash
# BASELINE config 5 on N GPUs of one box (multi-device engine, one process) next to one GPU.  usage: r02_sweep8.sh N
n=${1:-8}
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k multi_device 2>&1 | tail -1
for d in 1 2 4 $n; do [ $d -le $n ] && python profiles/sweep.py --devices $d > gpurun_out/r02_sweep_${d}gpu.txt 2>&1; cat gpurun_out/r02_sweep_${d}gpu.txt; done
