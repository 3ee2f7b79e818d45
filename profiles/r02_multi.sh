#!/bin/bash
# round 2, N GPUs of one box: the multi-device engine of the C ABI (test, single-process bench) next to the one-process-per-GPU path
# (weak and strong scaling, gathered vector checked bit for bit).  usage: r02_multi.sh N
n=${1:-2}
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k multi_device > gpurun_out/r02_multi${n}_pytest.log 2>&1; tail -2 gpurun_out/r02_multi${n}_pytest.log
run() { tag=$1; shift; "$@" > gpurun_out/r02_multi${n}_${tag}.json 2> gpurun_out/r02_multi${n}_${tag}.err || tail -5 gpurun_out/r02_multi${n}_${tag}.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_multi${n}_${tag}.json").read().strip().splitlines()[-1])
    print("%-22s n_gpus %d  %s  value %9.0f evals/s  e2e %9.0f  ms/step %.3f  check %s" % ("${tag}", d["n_gpus"], d["scaling"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("gather_check")))
except Exception as exc:
    print("${tag}: no line (%r)" % exc)
PY
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"
run cfg3_1gpu python bench.py --steps 20 --warmup 3 --no-cpu
run cfg3_weak $TR bench.py --gpus $n --steps 20 --warmup 3
run cfg3_strong $TR bench.py --gpus $n --steps 20 --warmup 3 --scaling strong
run cfg3_single_weak python bench.py --gpus $n --steps 20 --warmup 3 --single-process --no-cpu
run cfg3_single_strong python bench.py --gpus $n --steps 20 --warmup 3 --single-process --scaling strong --no-cpu
run cfg4_strong1 python bench.py --workload cfg4 --steps 3 --warmup 3 --scaling strong --no-cpu
run cfg4_strong $TR bench.py --workload cfg4 --gpus $n --steps 5 --warmup 3 --scaling strong
run cfg4_single_strong python bench.py --workload cfg4 --gpus $n --steps 5 --warmup 3 --single-process --scaling strong --no-cpu
