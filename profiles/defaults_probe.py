"""evaluatorPLS defaults (maxNComp = 0: ~106 components) and maxNComp = 20 next to config 3 proper, on config-3 data: the generic
per-fit kernel with its stored loadings in a global scratch.   usage: python profiles/defaults_probe.py"""
import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from gaselect_b200 import synth, evaluators as E, _capi
X, y = synth.nir_spectra(150, 1000, seed=777)
subsets = synth.random_subsets(1000, 100, 10, 200, seed=42)
idx, off = _capi.to_csr(subsets)
for kw in (dict(), dict(maxNComp=20), dict(numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5)):
    with E.PLSEvaluator(X, y, 123, **kw) as ev:
        info = ev.prepare()
        ev.evaluate_csr(idx, off)
        t0 = time.perf_counter(); f, s = ev.evaluate_csr(idx, off); dt = time.perf_counter() - t0
        t = ev.timing()
        print(kw, "max_ncomp", info["max_ncomp"], "fits/eval", info["fits_per_evaluation"], "100 chromosomes: %.2f ms (K1 %.2f) -> %.0f evals/s" % (dt * 1e3, t["fit_kernel_ms"], 100 / dt), "launches", t["kernel_launches"], "status", np.unique(s), flush=True)
