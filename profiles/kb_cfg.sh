#!/bin/bash
# K1x configuration sweep: GSB_SCORE_CFG=nf,cs,vg forces fits per CTA, cluster size, loadings in the global scratch
for k in ${KS:-64 96 112 128 160 200}; do
  for cfg in ${CFGS:-16,1,0 16,1,1 8,1,0 8,1,1 16,2,0 16,2,1 8,2,0 8,2,1}; do
    out=$(GSB_SCORE_CFG=$cfg python profiles/k1_probe.py $k 296 3 | grep -v phase | grep -v tile)
    echo "cfg=$cfg $out"
  done
done
