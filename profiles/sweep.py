"""BASELINE config 5: single-generation fitness throughput sweep, population x mean subset size, on one GPU or on the
multi-device engine of the C ABI (one process, --devices N: every generation dealt over N GPUs, device time = the
slowest GPU's kernels).
usage: python profiles/sweep.py [--quick] [--devices N]   (config-3 data and CV; subset sizes k ~ round(U(0.5,1.5) * kbar))"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gaselect_b200 import synth, evaluators as E, _capi

quick = "--quick" in sys.argv
ndev = int(sys.argv[sys.argv.index("--devices") + 1]) if "--devices" in sys.argv else 1
X, y = synth.nir_spectra(150, 1000, seed=777)
ev = E.PLSEvaluator(X, y, 123, numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5,
                    devices=list(range(ndev)) if ndev > 1 else None)
ev.prepare()
pops = [100, 500, 2000] if quick else [100, 500, 2000, 5000, 20000]
print("%d GPU(s): device evals/s (K1+K2, population resident) | e2e evals/s (host buffers in, fitness out)" % ndev)
print("%8s" % "pop\\kbar" + "".join("%22d" % kb for kb in (10, 20, 50, 100, 200)))
for pop in pops:
    row = "%8d" % pop
    for kbar in (10, 20, 50, 100, 200):
        if kbar == 200:  # keep k <= 256 (the fast path); U(0.5, 1.5) * 200 reaches 300
            subsets = [s[:256] for s in synth.sweep_subsets(1000, pop, kbar, seed=kbar)]
        else:
            subsets = synth.sweep_subsets(1000, pop, kbar, seed=kbar)
        idx, off = _capi.to_csr(subsets)
        ev.evaluate_csr(idx, off)
        best_e, best_k = 1e9, 1e9
        for _ in range(3):
            t0 = time.perf_counter(); ev.evaluate_csr(idx, off); best_e = min(best_e, time.perf_counter() - t0)
            t = ev.timing(); best_k = min(best_k, (t["fit_kernel_ms"] + t["reduce_kernel_ms"]) * 1e-3)
        row += "%11.0f|%10.0f" % (pop / best_k, pop / best_e)
    print(row, flush=True)
