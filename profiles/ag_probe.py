"""Where do the 0.13 ms per step go at N >= 2?  Times, on every rank: the step without the collective, with the NCCL
all-gather, and the all-gather alone.  run: torchrun --nproc-per-node 2 profiles/ag_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from gaselect_b200 import synth, evaluators as E, _capi, multigpu

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
if os.environ.get("AG_LAZY"):
    dist.init_process_group("nccl")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
if os.environ.get("AG_WARM"):  # create the communicator now
    _t = torch.zeros(8, device="cuda"); dist.all_reduce(_t); torch.cuda.synchronize()
X, y = synth.nir_spectra(150, 1000, seed=777)
gen = synth.random_subsets(1000, 500 * world, 10, 200, seed=42)
sizes = [len(s) for s in gen]
mine = multigpu.deal(sizes, world, cost=multigpu.kspace_cost(sizes), equal_counts=True)[rank]
idx, off = _capi.to_csr([gen[i] for i in mine])
ev = E.PLSEvaluator(X, y, 123, device=lr, numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5)
ev.prepare()
stream = torch.cuda.Stream(device=lr); torch.cuda.set_stream(stream); ev.set_stream(stream.cuda_stream)
ev.upload_population(idx, off)
f_ptr, s_ptr, npop = ev.device_results()
class _Dev:
    def __init__(self, ptr, n): self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
d_fit = torch.as_tensor(_Dev(f_ptr, npop), device=torch.device("cuda", lr))
gathered = torch.empty(world * npop, dtype=torch.float64, device=d_fit.device)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d_fit.device)

def run(fn, steps=50):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps): fn()
    e1.record(stream); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=d_fit.device); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])

def compute(): flush.zero_(); ev.enqueue_resident()
def both(): compute(); dist.all_gather_into_tensor(gathered, d_fit)
def gather_only(): dist.all_gather_into_tensor(gathered, d_fit)
a, b, c = run(compute), run(both), run(gather_only, 200)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): ev.enqueue_resident()
host_enq = (time.perf_counter() - t0) / 50 * 1e3
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): flush.zero_()
host_flush = (time.perf_counter() - t0) / 50 * 1e3
torch.cuda.synchronize()
def noflush(): ev.enqueue_resident()
d = run(noflush)
ev.evaluate_resident(); k1 = ev.timing()["fit_kernel_ms"]
ks = torch.tensor([k1], device=d_fit.device); lo = ks.clone(); dist.all_reduce(ks, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
if rank == 0:
    print("world %d: step without collective %.3f ms, with all-gather %.3f ms, all-gather alone %.3f ms; K1 per rank %.3f .. %.3f ms" % (world, a, b, c, float(lo[0]), float(ks[0])))
    print("host time to enqueue one step %.3f ms, one flush %.3f ms; step without flush %.3f ms" % (host_enq, host_flush, d))
dist.destroy_process_group()
