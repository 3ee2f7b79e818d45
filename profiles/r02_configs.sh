#!/bin/bash
# round 2: one bench.py line per BASELINE configuration (own arm, with the CPU baseline of the same run), the reference arm of
# the headline configuration, and the GPU test suite -> gpurun_out/r02_final_* (copied to profiles/r02/)
tag=${1:-final}
python -m pytest tests -m gpu -q > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -2 gpurun_out/r02_${tag}_pytest.log
for cfg in cfg3 cfg1 cfg2 cfg4 cfg4w; do
  steps=20; [ $cfg = cfg4 ] && steps=5; [ $cfg = cfg4w ] && steps=3
  python bench.py --workload $cfg --steps $steps --warmup 3 --cpu-seconds 8 > gpurun_out/r02_${tag}_${cfg}.json 2> gpurun_out/r02_${tag}_${cfg}.err || tail -3 gpurun_out/r02_${tag}_${cfg}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_${tag}_${cfg}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-6s value %9.0f evals/s  e2e %9.0f  ms/step %7.3f  roofline %s %.3f (%.1f of %.1f %s)  cpu %s" % ("${cfg}", d["value"], d["e2e"]["value"], d["ms_per_step"], r["bound"], r["frac"], r["achieved"], r["peak"], r["unit"], d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 1)))
except Exception as exc:
    print("${cfg}: no line (%r)" % exc)
PY
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_${tag}_cfg3_reference.json 2> gpurun_out/r02_${tag}_cfg3_reference.err; cut -c1-160 gpurun_out/r02_${tag}_cfg3_reference.json
