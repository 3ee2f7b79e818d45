#!/bin/bash
# round 2: per-generation packed Grams (gramPackKernel) in front of the per-fit kernel: tests, cfg3/cfg4 bench, cfg4 phase stamps
python -m pytest tests -m gpu -x -q > gpurun_out/r02_wip1_pytest.log 2>&1; tail -5 gpurun_out/r02_wip1_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_wip1_cfg3.json 2> gpurun_out/r02_wip1_cfg3.err; cat gpurun_out/r02_wip1_cfg3.json
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_wip1_cfg4.json 2> gpurun_out/r02_wip1_cfg4.err; cat gpurun_out/r02_wip1_cfg4.json
for k in 24 64 110 200; do GSB_TRACE=1 python profiles/k1_probe4.py $k 148; done > gpurun_out/r02_wip1_probe4.txt 2>&1; cat gpurun_out/r02_wip1_probe4.txt
