"""Per-fit path at config-4 shape for WIDE subsets (the spill path above 212 columns, the in-place path above 384):
device time per subset size.   usage: python profiles/wide_probe.py [POP]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gaselect_b200 import synth, evaluators as E, _capi

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 74
X, y = synth.nir_spectra(500, 5000, seed=778)
rng = np.random.default_rng(1)
with E.PLSEvaluator(X, y, 123, numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5) as ev:
    for k in (200, 212, 224, 256, 300, 352, 384, 400):
        subsets = [np.sort(rng.choice(5000, size=k, replace=False)).astype(np.uint32) for _ in range(pop)]
        idx, off = _capi.to_csr(subsets)
        ev.evaluate_csr(idx, off)
        best = 1e9
        for _ in range(3):
            ev.evaluate_csr(idx, off)
            best = min(best, ev.timing()["fit_kernel_ms"])
        print("cfg4 shape k=%3d pop=%d: K1 %.3f ms = %.1f us per chromosome-SM" % (k, pop, best, best * 1e3 * 148 / pop), flush=True)
