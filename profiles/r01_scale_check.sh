#!/bin/bash
# 1 / 2 / 4 / 8 GPUs on one box (run under gpurun --gpus 8), 50 steps each
python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 50 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29619 bench.py --gpus 8 --steps 10 --warmup 3 --workload cfg4 > gpurun_out/bench_n8_cfg4.json 2> gpurun_out/bench_n8_cfg4.err
for f in gpurun_out/bench_n1.json gpurun_out/bench_n2.json gpurun_out/bench_n4.json gpurun_out/bench_n8.json gpurun_out/bench_n8_cfg4.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], 'GPUs', round(d['value']), 'evals/s', round(d['ms_per_step'],3), 'ms/step e2e', round(d['e2e']['value']), 'clock samples', d['clocks']['samples'], d['clocks']['reasons'])"; done
