import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from gaselect_b200 import synth, evaluators as E
from oracle import bindings
orc = bindings.OracleLib()
rng = np.random.default_rng(7)
for case, (n, p, R, O, I, A, kmax) in enumerate([(600, 300, 2, 5, 4, 10, 200), (1500, 100, 2, 3, 3, 6, 100), (4000, 60, 1, 1, 5, 8, 60), (1000, 2000, 1, 4, 3, 10, 120),
                                                  (300, 800, 3, 10, 9, 10, 250), (161, 500, 2, 5, 4, 10, 300), (192, 500, 2, 5, 4, 10, 230), (160, 500, 2, 5, 4, 12, 200),
                                                  (150, 1000, 10, 5, 4, 10, 1000), (2000, 20, 2, 1, 7, 0, 20)]):
    gauss = A == 0 or case % 2 == 0
    X, y = (synth.gaussian(n, p, seed=case) if gauss else synth.nir_spectra(n, p, seed=case))
    sv = orc.make_seedvec(100 + case)
    sizes = [int(rng.integers(1, kmax + 1)) for _ in range(12)] + [kmax]
    subsets = [np.sort(rng.choice(p, size=k, replace=False)).astype(np.uint32) for k in sizes]
    kw = dict(numReplications=R, innerSegments=I, outerSegments=O, maxNComp=A)
    try:
        want, wst = orc.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=16)
        with E.PLSEvaluator(X, y, sv, **kw) as ev:
            got, st = ev.evaluate_population(subsets)
            t = ev.timing()
        ok = (wst == 0) & np.isfinite(want)
        err = np.abs(got[ok] - want[ok]) / np.abs(want[ok])
        print("n=%4d p=%4d R=%d O=%d I=%d A=%d kmax=%d %s: status %s max rel err %.1e  (K1 %.2f ms, %d launches, retries %d)" % (
            n, p, R, O, I, A, kmax, "gauss" if gauss else "nir", "same" if (np.where(st == 4, 0, st) == wst).all() else "DIFFER %s %s" % (st.tolist(), wst.tolist()), err.max() if ok.any() else 0.0,
            t["fit_kernel_ms"], t["kernel_launches"], t["retries"]), flush=True)
    except Exception as exc:
        print("n=%d p=%d %r: EXCEPTION %r" % (n, p, kw, exc), flush=True)
