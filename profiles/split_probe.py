"""K1k CTAs-per-chromosome policy probe: device time of one generation for a population / mean subset size under the
default policy and forced splits.   usage: python profiles/split_probe.py POP KBAR"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gaselect_b200 import synth, evaluators as E, _capi

pop, kbar = int(sys.argv[1]), int(sys.argv[2])
first = int(sys.argv[3]) if len(sys.argv) > 3 else 99  # only the first N configurations
X, y = synth.nir_spectra(150, 1000, seed=777)
subsets = [s[:256] for s in synth.sweep_subsets(1000, pop, kbar, seed=kbar)]
idx, off = _capi.to_csr(subsets)
for env in ({}, {"GSB_NO_EARLYSTOP": "1"}, {"GSB_KSPACE_SPLIT": "1"}, {"GSB_KSPACE_SPLIT": "2"}, {"GSB_KSPACE_SPLIT": "3"}, {"GSB_KSPACE_SPLIT": "4"},
            {"GSB_KSPACE_SPLIT": "2", "GSB_NO_EARLYSTOP": "1"})[:first]:
    for k in ("GSB_NO_EARLYSTOP", "GSB_KSPACE_SPLIT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    with E.PLSEvaluator(X, y, 123, numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5) as ev:
        ev.evaluate_csr(idx, off)
        best = 1e9
        for _ in range(5):
            ev.evaluate_csr(idx, off)
            t = ev.timing()
            best = min(best, t["fit_kernel_ms"] + t["reduce_kernel_ms"])
        print("pop %5d kbar %3d %-50s K1+K2 %.3f ms  %8.0f evals/s  CTAs %d  tensor GFLOP %.2f" % (pop, kbar, env, best, pop / best * 1e3, t["fit_ctas"], t["fit_tensor_gflop"]), flush=True)
