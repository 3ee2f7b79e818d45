#!/bin/bash
# ncu --set full of the per-fit kernel (one launch) at cfg4 shape, one subset size; usage: r02_ncu_fit.sh K [KERNEL_REGEX] [TAG]
k=${1:-200}; rx=${2:-simplsFit}; tag=${3:-fit_k$k}
ncu --set full --import-source on --clock-control none -k regex:$rx --launch-skip 1 --launch-count 1 -o gpurun_out/r02_$tag -f python profiles/k1_probe4.py $k 148 > gpurun_out/r02_${tag}_ncu.log 2>&1
tail -3 gpurun_out/r02_${tag}_ncu.log
