#!/bin/bash
# round 2: per-fit kernel with the Gram in tensor memory (K1t) + tensor-core prediction: tests, cfg3/cfg4 bench, phase stamps, launch lists
tag=${1:-wip3}
python -m pytest tests -m gpu -x -q > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -5 gpurun_out/r02_${tag}_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_${tag}_cfg3.json 2> gpurun_out/r02_${tag}_cfg3.err; cut -c1-330 gpurun_out/r02_${tag}_cfg3.json
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_${tag}_cfg4.json 2> gpurun_out/r02_${tag}_cfg4.err; cut -c1-330 gpurun_out/r02_${tag}_cfg4.json
for k in 24 64 110 160 200 230; do GSB_TRACE=1 python profiles/k1_probe4.py $k 148; done > gpurun_out/r02_${tag}_probe4.txt 2>&1; cat gpurun_out/r02_${tag}_probe4.txt
bash profiles/r02_ncu_probe4.sh 200 | tail -5
bash profiles/r02_ncu_probe4.sh 110 | tail -5
bash profiles/r02_ncu_probe4.sh 40 | tail -5
