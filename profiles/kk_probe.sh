#!/bin/bash
# K1k (kernel-space SIMPLS) per subset size; GSB_KSPACE_SPLIT forces the CTAs per chromosome
for k in 64 100 160 200; do
  python profiles/k1_probe.py $k 296 3 | grep -v phase
done
for sp in 1 2 3 4; do echo "split=$sp $(GSB_KSPACE_SPLIT=$sp python profiles/k1_probe.py 128 296 3 | grep -v phase)"; done
GSB_TRACE=1 python profiles/k1_probe.py 128 296 2
