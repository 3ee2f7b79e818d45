#!/bin/bash
# round 2: the reference's GA drivers on the engine (drop-in tests) and end-to-end genAlg() timing
tag=${1:-ga}
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_kspace.py -m gpu -x -q > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -3 gpurun_out/r02_${tag}_pytest.log
python profiles/ga_bench.py cfg3 50 --skip-1core > gpurun_out/r02_${tag}_cfg3.txt 2>&1; cat gpurun_out/r02_${tag}_cfg3.txt
python profiles/ga_bench.py cfg1 20 > gpurun_out/r02_${tag}_cfg1.txt 2>&1; cat gpurun_out/r02_${tag}_cfg1.txt
