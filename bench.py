#!/usr/bin/env python
"""bench.py -- chromosome fitness evaluations per second (repeated-CV PLS) on B200.

A "step" is ONE generation's fitness pass: every chromosome of the population is scored by
R*O*(I+1) cross-validated SIMPLS fits (BASELINE.json config "evaluatorPLS repeated double CV
(10 outer reps, 5/4 segments, <=10 components), n=150 p=1000 NIR-like spectra, population 500").
At N GPUs every rank scores its own 500-chromosome shard of a 500*N population (weak scaling; the
units are independent, the only exchange is the all-gather of the fitness vectors over NCCL);
--scaling strong splits ONE generation (500 at cfg3, 2000 at cfg4) over the ranks instead.  Rank 0
checks once, outside the timed region, that the un-permuted gathered vector equals its own
single-GPU evaluation of the whole generation bit for bit.  --single-process drives the same N
GPUs from ONE process through the multi-device engine of the C ABI (gsb_config.devices).

  value     device-resident throughput: population already in HBM, K1+K2 enqueued per step
  e2e       through the public call (gsb_evaluate_indices) with HOST buffers: per step the CSR
            lists go host->device and fitness/status come back
  roofline  K1 of the workload: cfg3 = simplsKspaceKernel (FP64 tensor pipe: algorithmic FLOPs of SURVEY.md 8d and the
            DMMA FLOPs issued, against the measured cuBLAS DGEMM rate), cfg4 = the per-fit Gram-space path
            (gramPack + simplsFitFast + predictBlocks; HBM: algorithmic bytes of SURVEY.md 8d / its CUDA-event time);
            traffic = DRAM bytes of those launches from the committed ncu pass (profiles/r02_traffic.json) or null
  L2        the path's inputs are small by nature (Z is 1.2 MB): a 256 MB buffer is overwritten before every step,
            inside the timed region, so that no step finds its inputs in L2
  cpu_baseline / --impl reference: the reference's own evaluator code (oracle/_ref, built from
            /root/reference/src) on the box's host cores, on a bounded sample of the same population
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gaselect_b200 import synth  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[0]: evaluatorPLS defaults (R = 30 srCV, as many components as the subset allows), Gaussian data
    "cfg1": dict(n=100, p=200, pop=100, strong_pop=100, kmin=5, kmax=30, data_seed=777, data="gaussian",
                 kw=dict(numReplications=30, maxNComp=0, innerSegments=7, outerSegments=1)),
    # BASELINE.json configs[1]: evaluatorLM, BIC fitness (K3: OLS by Cholesky + explicit residuals)
    "cfg2": dict(n=200, p=500, pop=200, strong_pop=200, kmin=5, kmax=50, data_seed=777, data="gaussian", evaluator="lm", kw=dict(statistic="BIC")),
    # BASELINE.json configs[2]: the configuration the metric (repeated-CV PLS) is quoted on
    "cfg3": dict(n=150, p=1000, pop=500, strong_pop=500, kmin=10, kmax=200, data_seed=777,
                 kw=dict(numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5)),
    # BASELINE.json configs[3] shape, per-GPU shard of the 2000-chromosome population
    "cfg4": dict(n=500, p=5000, pop=250, strong_pop=2000, kmin=20, kmax=200, data_seed=778,
                 kw=dict(numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5)),
    # the same with SURVEY.md 8(d)'s subset sizes 20..400 (beyond 232 columns the packed Gram no longer fits in shared memory)
    "cfg4w": dict(n=500, p=5000, pop=250, strong_pop=2000, kmin=20, kmax=400, data_seed=778,
                  kw=dict(numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5)),
}
GA_SEED = 123
METRIC = "chromosome fitness evals/sec (repeated-CV PLS)"


def bytes_eval(k, seg_lengths):
    """algorithmic bytes of one evaluation.  seg_lengths: list over (rep, outer)
    groups of (list of inner test lengths, outer test length)."""
    def bytes_fit(k, nte):
        return 8.0 * (k * (k + 1) / 2 + 2 * k + 2) + 8.0 * nte * (k + 1)
    total = 16.0
    for inner, outer in seg_lengths:
        total += sum(bytes_fit(k, t) for t in inner) + bytes_fit(k, outer)
    return total


def flops_eval(k, seg_lengths, max_ncomp):
    """algorithmic FLOPs of one evaluation in Gram space,
    flops_fit = A (2 k^2 + 8 k) + 2 k A (A - 1) + 2 n_te k A + 3 n_te A, over the R*O*(I+1) fits the reference runs."""
    A = min(max_ncomp, k)
    def flops_fit(nte):
        return A * (2.0 * k * k + 8.0 * k) + 2.0 * k * A * (A - 1) + 2.0 * nte * k * A + 3.0 * nte * A
    return sum(sum(flops_fit(t) for t in inner) + flops_fit(outer) for inner, outer in seg_lengths)


def lm_bytes(k, n):
    """K3 (LMEvaluator): packed centred Gram entries, X'y and column means gathered, then the n x (k + 1) rows for the residuals."""
    return 8.0 * (k * (k + 1) / 2 + 2 * k) + 8.0 * n * (k + 1) + 16.0


def lm_flops(k, n):
    """K3: Cholesky k^3 / 3, two triangular solves 2 k^2, residuals 2 n k (src/LMEvaluator.cpp:34-37 does the same by QR: 2 n k^2)."""
    return k ** 3 / 3.0 + 2.0 * k * k + 2.0 * n * k + 3.0 * n


def make_data(w):
    if w.get("data") == "gaussian":
        return synth.gaussian(w["n"], w["p"], seed=w["data_seed"])
    return synth.nir_spectra(w["n"], w["p"], seed=w["data_seed"], noise=1e-2)


def group_lengths(seg, inner):
    per = 2 * (inner + 1)
    out = []
    for g in range(len(seg) // per):
        base = g * per
        out.append(([len(seg[base + 2 * j + 1]) for j in range(inner)], len(seg[base + 2 * inner + 1])))
    return out


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML while the timed region runs (the
    nvidia-smi clocks line of B200_PROFILING.md, without spawning a process next to the GPU work)."""

    def __init__(self, device, period=0.01):
        self.device, self.period, self.rows, self.thread, self.h = device, period, [], None, None
        self._stop = threading.Event()
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # no NVML: report that instead of a number
            self.err = repr(exc)
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                  int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.h is not None:
            self.rows = []
            self._stop.clear()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: " + getattr(self, "err", "")]}
        self._stop.set()
        self.thread.join(timeout=2)
        nv = self.nv
        masks = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(name for name, m in masks.items() if any(r[1] & m for r in self.rows))
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "power_w_max": max([r[2] for r in self.rows], default=None)}


def reference_library():
    from oracle import bindings
    if os.path.isdir("/root/reference/src"):
        bindings.build("ref")
    if bindings.have_ref():
        return bindings.RefLib(), "reference"
    bindings.build("port")
    return bindings.OracleLib(), "port"


def cpu_leg(w, X, y, subsets, budget_s):
    """The reference's own evaluator on the host cores: the generation's chromosomes dealt to all host
    threads (cloned evaluators over contiguous slices), repeated until about budget_s seconds of work."""
    lib, kind = reference_library()
    cores = os.cpu_count() or 1
    sv = lib.make_seedvec(GA_SEED)
    ev = lib.lm_evaluator(X, y, **w["kw"]) if w.get("evaluator") == "lm" else lib.pls_evaluator(X, y, sv, **w["kw"])
    ev.evaluate(subsets[:cores], threads=cores)  # warm-up
    sample = subsets[: max(cores, (len(subsets) // cores) * cores)]
    done, spent, rounds = 0, 0.0, 0
    while spent < budget_s and rounds < 1000:
        _, st, t = ev.evaluate(sample, threads=cores, return_time=True)
        done += len(sample)
        spent += t
        rounds += 1
    return {"value": done / spent, "unit": "evals/s", "cores": cores, "kind": kind,
            "sample": "%d x %d of the %d chromosomes of one generation (k %d..%d), %d host threads, %.1f s; linear algebra under "
                      "the reference code is the repo's stand-in for Armadillo/BLAS (no R in the image)" %
                      (rounds, len(sample), len(subsets), w["kmin"], w["kmax"], cores, spent)}, done, spent


def measured_traffic(workload, pop):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the K1 launches of ONE step, from the committed ncu
    pass over this command (profiles/r02_traffic.sh -> profiles/r02_traffic.json); None when there is no capture of
    this workload and population."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get("%s/%d" % (workload, pop))
    except (OSError, ValueError):
        return None, None
    if not rec:
        return None, None
    return float(rec["k1_dram_bytes_per_step"]), rec["source"]


def fp64_peak():
    """FP64 tensor peak by BASELINE.md section 3's protocol (cuBLAS DGEMM 8192^3, best of 10 and 4 s sustained:
    profiles/fp64_peak.py -> profiles/r02_fp64_peak.json); the DMMA issue-rate figure of the microbenchmark
    (profiles/micro/dmma_probe_b200.txt: one DMMA.8x8x4 per 16 cycles per sub-partition = 37.2 TFLOP/s) as the fallback."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r02_fp64_peak.json")))
        return float(rec["fp64_dgemm_tflops_sustained"]), "measured: cuBLAS DGEMM 8192^3 sustained, profiles/r02_fp64_peak.json (burst %.1f)" % rec["fp64_dgemm_tflops"]
    except (OSError, ValueError, KeyError):
        return 37.2, "DMMA issue rate, profiles/micro/dmma_probe_b200.txt"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--pop", type=int, default=0, help="chromosomes per GPU (default: the workload's)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's population per GPU; strong: ONE generation (cfg3 500, cfg4 2000) split over the GPUs")
    ap.add_argument("--single-process", action="store_true",
                    help="drive --gpus N devices from this ONE process through the multi-device engine of the C ABI")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ndev = args.gpus if args.single_process else 1   # devices driven by this process
    units = world * ndev                              # GPUs of the job
    w = dict(WORKLOADS[args.workload])
    if args.pop:
        w["pop"] = args.pop
        w["strong_pop"] = args.pop
    global_pop = w["strong_pop"] if args.scaling == "strong" else w["pop"] * units
    w["pop"] = -(-global_pop // units)               # chromosomes per GPU (the last GPUs may hold one less)
    X, y = make_data(w)
    lm = w.get("evaluator") == "lm"
    if lm:
        what = "%s: evaluatorLM %s, Gaussian n=%d p=%d, population %d per GPU, subset sizes U[%d,%d]" % (
            args.workload, w["kw"]["statistic"], w["n"], w["p"], w["pop"], w["kmin"], w["kmax"])
    else:
        what = "%s: evaluatorPLS %s R=%d outer=%d inner=%d maxNComp=%d, %s n=%d p=%d, population %d per GPU, subset sizes U[%d,%d]" % (
            args.workload, "rdCV" if w["kw"]["outerSegments"] > 1 else "srCV", w["kw"]["numReplications"], w["kw"]["outerSegments"],
            w["kw"]["innerSegments"], w["kw"]["maxNComp"], "Gaussian" if w.get("data") == "gaussian" else "NIR-like (noise 1e-2)",
            w["n"], w["p"], w["pop"], w["kmin"], w["kmax"])
    config = {"workload": what,
              "population_per_gpu": w["pop"], "global_population": global_pop, "ga_seed": GA_SEED,
              "parallelism": ("generation of %d chromosomes dealt to %d GPU(s) (equal counts, longest-processing-time-first on "
                              "the kernels' measured cost); " % (global_pop, units)) +
                             ("ONE process, multi-device engine of the C ABI (gsb_config.devices): peer-copy gather" if args.single_process
                              else "one process per GPU, fitness all-gather over NCCL")}

    if args.impl == "reference":
        if rank != 0:
            return
        subsets = synth.random_subsets(w["p"], global_pop, w["kmin"], w["kmax"], seed=42)[: max(w["pop"], 1)]
        lib, kind = reference_library()
        cores = os.cpu_count() or 1
        ev = lib.lm_evaluator(X, y, **w["kw"]) if lm else lib.pls_evaluator(X, y, lib.make_seedvec(GA_SEED), **w["kw"])
        _, _, t_probe = ev.evaluate(subsets[:cores], threads=cores, return_time=True)
        # bounded sample of the generation per step: the whole generation when the run stays under ~2.5 min
        afford = int(150.0 / (args.steps + args.warmup) / max(t_probe, 1e-4)) * cores
        per_step = max(cores, min(len(subsets), afford))
        for i in range(args.warmup):
            ev.evaluate(subsets[:per_step], threads=cores)
        total_t, total_n = 0.0, 0
        for s in range(args.steps):
            lo = (s * per_step) % max(1, len(subsets) - per_step + 1)
            _, _, t = ev.evaluate(subsets[lo:lo + per_step], threads=cores, return_time=True)
            total_t += t
            total_n += per_step
        v = total_n / total_t
        sample = "%d chromosomes of the generation per step, %d host threads" % (per_step, cores)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v,
                          "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": args.scaling,
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from gaselect_b200 import _capi, evaluators as E

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # no device_id: the NCCL communicator is then created by the first collective, i.e. in the warm-up steps, AFTER
        # the engine and the flush buffer have allocated their device memory.  Measured (profiles/ag_probe.py,
        # profiles/ag_env.sh): with the communicator created first every step takes 2.28 instead of 2.20 ms, K1 itself
        # unchanged -- the all-gather costs 12 us, the rest of the 1 -> N GPU gap was this allocation order
        dist.init_process_group("nccl")

    # the generation is dealt to the ranks (equal counts, balanced on the kernels' measured cost)
    from gaselect_b200 import multigpu
    generation = synth.random_subsets(w["p"], global_pop, w["kmin"], w["kmax"], seed=42)
    sizes = [len(s) for s in generation]
    cost = multigpu.kspace_cost(sizes) if (w["n"] <= 160 and not lm and 0 < w["kw"]["maxNComp"] <= 10) else multigpu.gram_cost(sizes)
    shares = multigpu.deal(sizes, world, cost=cost, equal_counts=True)
    subsets = [generation[i] for i in shares[rank]]  # (single process: everything; the engine deals it to its devices)
    idx, offsets = _capi.to_csr(subsets)
    ks = np.diff(offsets.astype(np.int64))

    def make_evaluator(**where):
        return E.LMEvaluator(X, y, **w["kw"], **where) if lm else E.PLSEvaluator(X, y, GA_SEED, **w["kw"], **where)

    ev = make_evaluator(devices=list(range(ndev))) if args.single_process else make_evaluator(device=local_rank)
    info = ev.prepare()
    # the engine works on torch's current stream, so torch.cuda.Event brackets see its kernels
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    if not args.single_process:
        ev.set_stream(stream.cuda_stream)
    all_ks = np.asarray(sizes, dtype=np.int64)
    if lm:
        algo_bytes = float(sum(lm_bytes(int(k), w["n"]) for k in all_ks)) / units
        algo_flops = float(sum(lm_flops(int(k), w["n"]) for k in all_ks)) / units
    else:
        seg = ev.getSegmentation()
        lengths = group_lengths(seg, info["inner_segments"])
        algo_bytes = float(sum(bytes_eval(int(k), lengths) for k in all_ks)) / units      # per GPU and step (mean over the GPUs)
        algo_flops = float(sum(flops_eval(int(k), lengths, info["max_ncomp"]) for k in all_ks)) / units

    fitness = np.zeros(len(subsets), dtype=np.float64)
    status = np.zeros(len(subsets), dtype=np.int32)
    ev.upload_population(idx, offsets)
    cap = max(len(s) for s in shares)
    dev = torch.device("cuda", local_rank)
    d_fit = send = gathered = None
    if not args.single_process:
        f_ptr, s_ptr, npop = ev.device_results()

        class _Dev:  # zero-copy view of the engine's result buffer for NCCL
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        d_fit = torch.as_tensor(_Dev(f_ptr, npop), device=dev)
        # equal-sized all-gather pieces: the engine's buffer itself when every rank holds `cap` chromosomes
        send = d_fit if npop == cap else torch.zeros(cap, dtype=torch.float64, device=dev)
        gathered = torch.empty(world * cap, dtype=torch.float64, device=dev) if world > 1 else None

    flush_bufs = []
    for d in range(ndev):  # 2 x L2 on every device this process drives
        flush_bufs.append(torch.empty(256 << 20, dtype=torch.uint8, device=torch.device("cuda", local_rank + d)))
    flush_buf = flush_bufs[0]

    def flush():
        flush_buf.zero_()  # L2 flush, on the timed stream
        for b in flush_bufs[1:]:
            b.zero_()

    def all_gather():
        if world > 1:
            if send is not d_fit:
                send[: d_fit.numel()].copy_(d_fit)
            dist.all_gather_into_tensor(gathered, send)

    def step_resident():
        flush()
        if args.single_process:
            ev.evaluate_resident()  # the replicas run on their own streams: the call returns when every device is done
        else:
            ev.enqueue_resident()
        all_gather()

    def barrier():
        if world > 1:
            dist.barrier()
        for d in range(ndev):
            torch.cuda.synchronize(local_rank + d)

    # ---- value: device-resident -------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    gather_check = None
    if world > 1:
        # chromosomes the kernel-space kernel declined (GSB_STATUS_RETRY; none on the bench populations) are re-scored
        # and patched into the device buffers by the download: gather once more so that the check below sees them
        ev.download_results(fitness, status)
        all_gather()
        barrier()
    if world > 1 and rank == 0:
        # the gathered vector, un-permuted, against this GPU's own evaluation of the WHOLE generation: bit for bit
        got = np.zeros(len(generation), dtype=np.float64)
        g = gathered.cpu().numpy().reshape(world, cap)
        for r, share in enumerate(shares):
            got[share] = g[r, : len(share)]
        with make_evaluator(device=local_rank) as whole:
            want, wst = whole.evaluate_population(generation)
        gather_check = {"chromosomes": len(generation), "bit_identical": bool((got == want).all() and np.isin(wst, (0, 4)).all())}
        assert gather_check["bit_identical"], "the gathered fitness vector differs from the single-GPU evaluation"
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    # one process per GPU: CUDA events on the stream the kernels run on.  Single process, several devices: the engine's
    # replicas run on their own streams, so the host clock around the synchronous calls is the measure
    ms = 1e3 * (time.perf_counter() - t_host) if args.single_process else e0.elapsed_time(e1)
    launches = ev.timing()["kernel_launches"] * args.steps
    # K1 timed live with the engine's CUDA events on the same stream, per step
    k1_ms, k2_ms = [], []
    tensor_gflop = 0.0
    for _ in range(args.steps):
        flush()
        ev.evaluate_resident()
        t = ev.timing()
        k1_ms.append(t["fit_kernel_ms"])
        k2_ms.append(t["reduce_kernel_ms"])
        tensor_gflop = t.get("fit_tensor_gflop", 0.0)
    clocks = sampler.stop() if rank == 0 else None
    ev.download_results(fitness, status)
    retries = ev.timing()["retries"]
    # 0 = ok, 4 = GSB_STATUS_FLAGGED (valid fitness, near-tie in the one-SE rule)
    assert np.isin(status, (0, 4)).all() and np.isfinite(fitness).all()

    # ---- e2e: public call with host buffers ----------------------------------------------------
    for _ in range(2):
        ev.evaluate_csr(idx, offsets)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        flush()
        f2, s2 = ev.evaluate_csr(idx, offsets)
        all_gather()
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
    assert (f2 == fitness).all()
    h2d = int(idx[: int(offsets[-1])].nbytes + offsets.nbytes + 4 * len(subsets))
    d2h = int(fitness.nbytes + status.nbytes)

    if args.single_process:
        single_check = None
        if ndev > 1:  # the multi-device engine against one device, bit for bit
            with make_evaluator(device=local_rank) as whole:
                want, wst = whole.evaluate_population(subsets)
            single_check = {"chromosomes": len(subsets), "bit_identical": bool((fitness == want).all() and (wst == status).all())}
            assert single_check["bit_identical"], "the multi-device engine differs from the single-device engine"
        gather_check = single_check
    if world > 1:
        tt = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tt[0]), float(tt[1])
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        k1 = float(np.mean(k1_ms))
        achieved = algo_bytes / (k1 * 1e-3) / 1e9
        peak64, peak64_src = fp64_peak()
        traffic, traffic_src = measured_traffic(args.workload, w["pop"])
        kspace = tensor_gflop > 0.0  # the kernel-space kernel ran (n <= 160, <= 10 components): the FP64 tensor pipe binds
        hbm = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
               "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
               "algorithmic_bytes_per_step": algo_bytes,
               "note": "SURVEY 8(d)'s algorithmic bytes (the logical inputs of every fit in the Gram formulation) of one GPU's share / K1 time"}
        tensor = {"bound": "tensor", "achieved": algo_flops * 1e-9 / max(k1, 1e-9), "peak": peak64, "unit": "TFLOP/s",
                  "frac": algo_flops * 1e-9 / max(k1, 1e-9) / peak64, "peak_source": peak64_src,
                  "algorithmic_gflop_per_step": algo_flops * 1e-9,
                  "issued_dmma_gflop_per_step": tensor_gflop, "issued_tflops": tensor_gflop / max(k1, 1e-9),
                  "issued_frac": tensor_gflop / max(k1, 1e-9) / peak64,
                  "note": "achieved = SURVEY 8(d)'s ALGORITHMIC Gram-space FLOPs (flops_fit over the R*O*(I+1) fits the reference runs) / K1 "
                          "time; issued = DMMA.8x8x4 x 512 FLOP the kernel-space kernel really executes (2 n^2 per component and fit, "
                          "independent of the subset size; the products are counted by the kernel, the K builds from the launch geometry)"}
        if kspace:
            main, other, other_key = tensor, hbm, "hbm_model"
            kernel = ("K1 = simplsKspaceKernel (kernel-space SIMPLS on the FP64 tensor cores: one K = Z Z' per chromosome in shared memory, "
                      "subsets of more than 40 columns) + gramPack / simplsFitFast / predictBlocks (per-fit Gram-space path; the rest), "
                      "all launches of a step")
            other = dict(other, note=other["note"] + "; the kernel-space kernel does not move those bytes (see traffic): kept for comparison "
                                                        "with round 1, not as the binding roofline")
        elif lm:
            main, other, other_key = hbm, dict(tensor, bound="fp64", note="K3's algorithmic FLOPs (Cholesky + solves + residuals) / its time against the FP64 tensor peak: not the binding resource"), "fp64"
            main = dict(main, note="K3's algorithmic bytes (gathered packed Gram entries, X'y, means; the n x (k + 1) rows for the explicit residuals) / its time")
            kernel = "K3 = olsFitKernel (one CTA per chromosome: Cholesky of the gathered centred Gram, explicit residuals), all launches of a step"
        elif info["max_ncomp"] > 10:
            main, other, other_key = dict(tensor, bound="fp64", note="SURVEY 8(d): with unconstrained components (A = k) the per-fit path is bound by the FP64 pipe, not HBM: "
                                          "algorithmic Gram-space FLOPs / K1 time against the measured FP64 (DGEMM) peak"), hbm, "hbm_model"
            kernel = "K1 = gramPackKernel + simplsFitKernel (generic Gram-space SIMPLS, more than 10 components, CUDA-core FP64), all launches of a step"
        else:
            main, other, other_key = hbm, tensor, "fp64"
            kernel = ("K1 = gramPackKernel (per-generation packed Grams, HBM-bound) + simplsFitFast (Gram-space SIMPLS, one CTA per "
                      "(chromosome, training set)) + predictBlocksKernel (held-out prediction on the FP64 tensor cores), all launches of a step")
        roof = dict(main, kernel=kernel, k1_ms_per_step=k1, k2_ms_per_step=float(np.mean(k2_ms)), traffic=traffic,
                    traffic_source=traffic_src)
        roof[other_key] = other
        line = {
            "metric": METRIC if not lm else "chromosome fitness evals/sec (OLS / BIC)",
            "value": global_pop * args.steps / (ms * 1e-3), "unit": "evals/s",
            "n_gpus": units, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config, l2="explicit flush: a 256 MB buffer (2 x L2) is overwritten before every step, inside the "
                                      "timed region (the path's small shared inputs, %.1f MB of centred data, would otherwise stay in L2)"
                                      % (8e-6 * ((w["n"] + 3) // 4 * 4) * (w["p"] + 2)),
                           fits_per_evaluation=info["fits_per_evaluation"], distinct_fits=info["distinct_fits"],
                           setup_ms=info["setup_ms"], gram_dmma_ms=info["gram_dmma_ms"],
                           gram_tflops=info["gram_flops"] / max(info["gram_dmma_ms"], 1e-9) / 1e9,
                           timing=("host clock around synchronous multi-device engine calls" if args.single_process
                                   else "CUDA events on the launching stream, max over ranks"),
                           gather_check=gather_check, flagged_near_ties=int((status == 4).sum()), retried_on_gram_kernel=int(retries)),
            "e2e": {"value": global_pop * args.steps / (e2e_ms * 1e-3), "unit": "evals/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
        }
        if not args.no_cpu and units == 1:
            cb, _, _ = cpu_leg(w, X, y, subsets, args.cpu_seconds)
            line["cpu_baseline"] = cb
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    ev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
