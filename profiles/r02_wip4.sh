#!/bin/bash
tag=${1:-wip4}
python -m pytest tests -m gpu -x -q > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -3 gpurun_out/r02_${tag}_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_${tag}_cfg3.json 2> gpurun_out/r02_${tag}_cfg3.err; cut -c1-250 gpurun_out/r02_${tag}_cfg3.json
python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_${tag}_cfg4.json 2> gpurun_out/r02_${tag}_cfg4.err; cut -c1-250 gpurun_out/r02_${tag}_cfg4.json
bash profiles/r02_launches.sh cfg4 ${tag}_cfg4 | head -8
