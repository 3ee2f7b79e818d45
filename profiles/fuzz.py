"""Randomised configurations of all three evaluators through the engine against the plain-C oracle: rows, columns, CV shape,
component limits (incl. 0 = unconstrained), statistics, subset sizes from 1 column to all of them.  Prints every case that
differs in status or by more than 1e-8.   usage: python profiles/fuzz.py [CASES] [SEED]"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gaselect_b200 import synth, evaluators as E, _capi
from oracle import bindings

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
orc = bindings.OracleLib()
bad = 0
for case in range(cases):
    kind = rng.choice(["pls", "pls", "pls", "fit", "lm"])
    n = int(rng.choice([8, 12, 30, 60, 100, 150, 161, 190, 250, 400]))
    p = int(rng.choice([3, 10, 40, 120, 300, 700]))
    gauss = rng.random() < 0.5
    X, y = (synth.gaussian(n, p, seed=case) if gauss else synth.nir_spectra(n, p, seed=case))
    sv = orc.make_seedvec(int(rng.integers(1, 1 << 30)))
    npop = int(rng.integers(1, 40))
    kmax = p if kind != "lm" else min(p, n - 2)
    sizes = [int(rng.integers(1, kmax + 1)) for _ in range(npop)]
    if rng.random() < 0.3:
        sizes[0] = kmax
    subsets = [np.sort(rng.choice(p, size=k, replace=False)).astype(np.uint32) for k in sizes]
    try:
        if kind == "pls":
            outer = int(rng.choice([1, 1, 2, 3, 5]))
            kw = dict(numReplications=int(rng.integers(1, 4)), innerSegments=int(rng.integers(2, 8)), outerSegments=outer,
                      maxNComp=int(rng.choice([0, 1, 3, 10, 14])) if gauss else int(rng.choice([1, 3, 6, 10])),
                      testSetSize=float(rng.choice([0.0, 0.2, 0.35])) if outer == 1 else 0.0, sdfact=float(rng.choice([1.0, 0.5, 2.0])),
                      complexityPenalty=float(rng.choice([0.0, 0.01])))
            want, wst = orc.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=8)
            with E.PLSEvaluator(X, y, sv, **kw) as ev:
                got, st = ev.evaluate_population(subsets)
        elif kind == "fit":
            kw = dict(numSegments=int(rng.integers(2, 9)), statistic=str(rng.choice(["BIC", "AIC", "adjusted.r.squared", "r.squared"])),
                      maxNComp=int(rng.choice([0, 2, 8, 12])) if gauss else int(rng.choice([2, 8])), sdfact=float(rng.choice([1.0, 0.5])))
            want, wst = orc.fit_evaluator(X, y, sv, **kw).evaluate(subsets, threads=8)
            with E.BICEvaluator(X, y, sv, **kw) as ev:
                got, st = ev.evaluate_population(subsets)
        else:
            kw = dict(statistic=str(rng.choice(["BIC", "AIC", "adjusted.r.squared", "r.squared"])))
            want, wst = orc.lm_evaluator(X, y, **kw).evaluate(subsets, threads=8)
            with E.LMEvaluator(X, y, **kw) as ev:
                got, st = ev.evaluate_population(subsets)
    except Exception as exc:
        bad += 1
        print("case %d %s n=%d p=%d %s %r: EXCEPTION %r" % (case, kind, n, p, "gauss" if gauss else "nir", kw, exc), flush=True)
        continue
    st = np.where(st == 4, 0, st)
    ok = wst == 0
    fin = ok & np.isfinite(want)
    err = np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1e-300) if fin.any() else np.zeros(1)
    nonfinite_same = bool((np.isnan(got[ok & ~fin]) == np.isnan(want[ok & ~fin])).all() and (got[ok & ~fin][~np.isnan(got[ok & ~fin])] == want[ok & ~fin][~np.isnan(want[ok & ~fin])]).all())
    flag = (st != wst).any() or err.max() > 1e-8 or not nonfinite_same
    bad += bool(flag)
    print("case %3d %-3s n=%3d p=%3d %-5s pop=%2d kmax=%3d %s  status %s  max rel err %.1e%s" % (
        case, kind, n, p, "gauss" if gauss else "nir", npop, max(sizes), kw, "same" if (st == wst).all() else "DIFFER %s vs %s" % (st.tolist(), wst.tolist()),
        err.max(), "   <-- CHECK" if flag else ""), flush=True)
print("%d of %d cases need a look" % (bad, cases))
