#!/bin/bash
# round 2, final build: everything the tables of profiles/README.md and DESIGN.md quote, in one gpurun call.
# Outputs: gpurun_out/r02_final_* (the JSON lines and texts are then copied to profiles/r02/)
tag=final
python -m pytest tests -m gpu -q > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -2 gpurun_out/r02_${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_${tag}_smoke.txt 2>&1; tail -1 gpurun_out/r02_${tag}_smoke.txt
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
python profiles/sweep.py > gpurun_out/r02_${tag}_sweep.txt 2>&1; tail -6 gpurun_out/r02_${tag}_sweep.txt
python profiles/ga_bench.py cfg3 50 --skip-1core > gpurun_out/r02_${tag}_ga_cfg3.txt 2>&1; cat gpurun_out/r02_${tag}_ga_cfg3.txt
python profiles/ga_bench.py cfg1 20 > gpurun_out/r02_${tag}_ga_cfg1.txt 2>&1; cat gpurun_out/r02_${tag}_ga_cfg1.txt
python profiles/stress.py > gpurun_out/r02_${tag}_stress.txt 2>&1; cat gpurun_out/r02_${tag}_stress.txt
for a in "62 100" "100 100" "200 100" "300 100" "448 100" "500 100"; do python profiles/split_probe.py $a 1; done > gpurun_out/r02_${tag}_split.txt 2>&1; cat gpurun_out/r02_${tag}_split.txt
python profiles/pred_probe.py cfg3 500 > gpurun_out/r02_${tag}_pred.txt 2>&1; python profiles/pred_probe.py cfg4 148 >> gpurun_out/r02_${tag}_pred.txt 2>&1; cat gpurun_out/r02_${tag}_pred.txt
python profiles/wide_probe.py 74 > gpurun_out/r02_${tag}_wide.txt 2>&1; cat gpurun_out/r02_${tag}_wide.txt
# the K1 launches of ONE config-4 generation (15 classes x pack, vectors, fit, predict), selected raw metrics
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:'simpls|gramPack|predictBlocks' -s 240 -c 60 -o /tmp/r02_cfg4_k1 -f \
    python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_cfg4_ncu_f.log 2>&1; tail -1 gpurun_out/r02_cfg4_ncu_f.log
ncu -i /tmp/r02_cfg4_k1.ncu-rep --page raw --csv > gpurun_out/r02_cfg4_k1_raw.csv
du -sm gpurun_out
