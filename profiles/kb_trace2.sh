for k in 16 64; do GSB_SCORE_CFG=16,1,0 GSB_TRACE=1 python profiles/k1_probe.py $k 296 2; done
for k in 64 96; do GSB_SCORE_CFG=8,1,0 GSB_TRACE=1 python profiles/k1_probe.py $k 296 2; done
for k in 160 ; do GSB_SCORE_CFG=8,2,0 GSB_TRACE=1 python profiles/k1_probe.py $k 296 2; done
for k in 200 ; do GSB_SCORE_CFG=16,2,1 GSB_TRACE=1 python profiles/k1_probe.py $k 296 2; done
