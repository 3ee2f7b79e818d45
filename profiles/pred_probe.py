"""Per-fit path policy probe: fit-only kernel + predictBlocksKernel against the per-fit kernel predicting in place
(GSB_NO_PREDICT=1), by subset size.   usage: python profiles/pred_probe.py [cfg3|cfg4] POP"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gaselect_b200 import synth, evaluators as E, _capi

which, pop = sys.argv[1], int(sys.argv[2])
n, p = (150, 1000) if which == "cfg3" else (500, 5000)
X, y = synth.nir_spectra(n, p, seed=777)
rng = np.random.default_rng(1)
for k in ((8, 16, 24, 32, 40, 64) if which == "cfg3" else (24, 40, 64, 110, 160, 200)):
    subsets = [np.sort(rng.choice(p, size=k, replace=False)).astype(np.uint32) for _ in range(pop)]
    idx, off = _capi.to_csr(subsets)
    line = "%s k=%3d pop=%d:" % (which, k, pop)
    for env in ({}, {"GSB_NO_PREDICT": "1"}):
        os.environ.pop("GSB_NO_PREDICT", None)
        os.environ.update(env)
        os.environ["GSB_NO_KSPACE"] = "1"
        os.environ["GSB_NO_SCORES"] = "1"
        with E.PLSEvaluator(X, y, 123, numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5) as ev:
            ev.evaluate_csr(idx, off)
            best = 1e9
            for _ in range(4):
                ev.evaluate_csr(idx, off)
                t = ev.timing()
                best = min(best, t["fit_kernel_ms"])
        line += "  %s %.3f ms" % ("in-kernel prediction" if env else "fit-only + predictBlocks", best)
    print(line, flush=True)
