#!/bin/bash
# weak scaling of the bench workload on one box (run under gpurun --gpus 8): N = 1, 2, 4, 8; config 4 on 8 GPUs
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --workload cfg4 > gpurun_out/bench_n8_cfg4.json 2> gpurun_out/bench_n8_cfg4.err
for f in gpurun_out/bench_n1.json gpurun_out/bench_n2.json gpurun_out/bench_n4.json gpurun_out/bench_n8.json gpurun_out/bench_n8_cfg4.json; do python - "$f" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], "GPUs", round(d["value"]), "evals/s", round(d["ms_per_step"], 3), "ms/step e2e", round(d["e2e"]["value"]))
except Exception as ex:
    print(sys.argv[1], "failed", ex)
P
done
