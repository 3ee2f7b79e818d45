#!/bin/bash
# the tables of profiles/README.md: per-configuration timing, config-5 sweep, end-to-end GA, K1 per subset size
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python profiles/cfg_probe.py > gpurun_out/cfg_probe.txt 2>&1; cat gpurun_out/cfg_probe.txt
python profiles/sweep.py > gpurun_out/sweep.txt 2>&1; cat gpurun_out/sweep.txt
bash profiles/kb_probe.sh > gpurun_out/k1_probe.txt 2>&1; cat gpurun_out/k1_probe.txt
(python profiles/ga_bench.py cfg1 20; python profiles/ga_bench.py cfg3 50 --skip-1core) > gpurun_out/ga_bench.txt 2>&1; tail -30 gpurun_out/ga_bench.txt
