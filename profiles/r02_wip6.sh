#!/bin/bash
# round 2: early stop of the outer-only fits in K1k + near-tie flag: GPU tests, cfg3 / cfg4 bench lines, A/B against GSB_NO_EARLYSTOP
tag=${1:-wip6}
python -m pytest tests -m gpu -x -q > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -5 gpurun_out/r02_${tag}_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_${tag}_cfg3.json 2> gpurun_out/r02_${tag}_cfg3.err; cut -c1-250 gpurun_out/r02_${tag}_cfg3.json
GSB_NO_EARLYSTOP=1 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_${tag}_cfg3_noearly.json 2> gpurun_out/r02_${tag}_cfg3_noearly.err; cut -c1-250 gpurun_out/r02_${tag}_cfg3_noearly.json
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_${tag}_cfg4.json 2> gpurun_out/r02_${tag}_cfg4.err; cut -c1-250 gpurun_out/r02_${tag}_cfg4.json
