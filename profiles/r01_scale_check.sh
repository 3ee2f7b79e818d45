for n in 8 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 50 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
done
python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
for f in gpurun_out/bench_n1.json gpurun_out/bench_n2.json gpurun_out/bench_n8.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['samples'])"; done
