#!/bin/bash
# round 2: fit-only kernel + tensor-core prediction per (chromosome, held-out block): tests, cfg3/cfg4 bench, cfg4 phase stamps
python -m pytest tests -m gpu -x -q > gpurun_out/r02_wip2_pytest.log 2>&1; tail -5 gpurun_out/r02_wip2_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_wip2_cfg3.json 2> gpurun_out/r02_wip2_cfg3.err; cut -c1-400 gpurun_out/r02_wip2_cfg3.json
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_wip2_cfg4.json 2> gpurun_out/r02_wip2_cfg4.err; cut -c1-400 gpurun_out/r02_wip2_cfg4.json
for k in 24 64 110 200; do GSB_TRACE=1 python profiles/k1_probe4.py $k 148; done > gpurun_out/r02_wip2_probe4.txt 2>&1; cat gpurun_out/r02_wip2_probe4.txt
