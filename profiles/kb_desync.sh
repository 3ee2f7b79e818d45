#!/bin/bash
# 16-fit single CTA: CTA-wide barriers (0) against the two 8-fit tiles on named barriers (1), started out of phase (2, 3)
for k in 48 64; do
  for m in 0 1 2 3; do
    echo "desync=$m $(GSB_DESYNC=$m GSB_SCORE_CFG=16,1,0 python profiles/k1_probe.py $k 296 3 | grep -v phase | grep -v tile)"
  done
done
for k in 96 112; do
  for m in 0 1 2 3; do
    echo "desync=$m $(GSB_DESYNC=$m GSB_SCORE_CFG=16,1,1 python profiles/k1_probe.py $k 296 3 | grep -v phase | grep -v tile)"
  done
done
GSB_DESYNC=2 GSB_TRACE=1 GSB_SCORE_CFG=16,1,0 python profiles/k1_probe.py 64 296 2
GSB_DESYNC=2 python -m pytest tests/test_gpu_scores.py -x -q 2>&1 | tail -3
