import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from gaselect_b200 import synth, evaluators as E, _capi
X, y = synth.nir_spectra(150, 1000, seed=777)
subsets = synth.random_subsets(1000, 500, 10, 200, seed=42)
idx, off = _capi.to_csr(subsets)
with E.PLSEvaluator(X, y, 123, numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5) as ev:
    for _ in range(5):
        ev.evaluate_csr(idx, off)
    ts = []
    for _ in range(50):
        t0 = time.perf_counter(); ev.evaluate_csr(idx, off); ts.append(time.perf_counter() - t0)
    t = ev.timing()
    print("e2e call: median %.3f ms, min %.3f ms; device: h2d %.3f fit %.3f reduce %.3f d2h %.3f total %.3f ms" % (
        np.median(ts) * 1e3, min(ts) * 1e3, t["h2d_ms"], t["fit_kernel_ms"], t["reduce_kernel_ms"], t["d2h_ms"], t["total_ms"]))
    fitness = np.zeros(500); status = np.zeros(500, dtype=np.int32)
    ts2 = []
    for _ in range(50):
        t0 = time.perf_counter(); ev.upload_population(idx, off); t1 = time.perf_counter(); ev.evaluate_resident(); t2 = time.perf_counter(); ev.download_results(fitness, status); t3 = time.perf_counter()
        ts2.append((t1 - t0, t2 - t1, t3 - t2))
    a = np.median(np.array(ts2), axis=0) * 1e3
    print("three calls: upload %.3f ms, evaluate %.3f ms, download %.3f ms" % tuple(a))
