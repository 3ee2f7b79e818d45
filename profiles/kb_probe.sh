#!/bin/bash
# per-subset-size timing of K1: score-space tensor-core kernel (K1x); GSB_NO_SCORES=1 gives the per-fit kernel
for k in 16 32 48 64 100 128 160 200 300; do
  python profiles/k1_probe.py $k 296 3 | grep -v phase
done
