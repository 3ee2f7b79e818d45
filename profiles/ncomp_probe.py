import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from gaselect_b200 import synth, evaluators as E, _capi
X, y = synth.nir_spectra(150, 1000, seed=777)
for k in (60, 200):
    subsets = synth.random_subsets(1000, 74, k, k, seed=42)
    idx, off = _capi.to_csr(subsets)
    for A in (10, 11, 20, 40, 80, 0):
        with E.PLSEvaluator(X, y, 123, maxNComp=A) as ev:
            info = ev.prepare()
            ev.evaluate_csr(idx, off)
            ev.evaluate_csr(idx, off)
            t = ev.timing()
            print("k=%d maxNComp=%d (effective %d): K1 %.2f ms for 74 chromosomes x %d fits, launches %d" % (k, A, min(info["max_ncomp"], k), t["fit_kernel_ms"], info["fits_per_evaluation"], t["kernel_launches"]), flush=True)
