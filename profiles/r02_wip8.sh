#!/bin/bash
tag=${1:-wip8}
python -m pytest tests -m gpu -x -q > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -3 gpurun_out/r02_${tag}_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_${tag}_cfg3.json 2> gpurun_out/r02_${tag}_cfg3.err; cut -c1-250 gpurun_out/r02_${tag}_cfg3.json
python profiles/ga_bench.py cfg3 50 --skip-1core --no-ref > gpurun_out/r02_${tag}_ga_cfg3.txt 2>&1; cat gpurun_out/r02_${tag}_ga_cfg3.txt
python profiles/sweep.py > gpurun_out/r02_${tag}_sweep.txt 2>&1; tail -12 gpurun_out/r02_${tag}_sweep.txt
