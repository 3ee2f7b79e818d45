"""Regenerates tests/golden/reference_vectors.json by EXECUTING THE REFERENCE'S OWN CODE
(oracle/_ref/libgaselect_ref.so, built by `make -C oracle ref` from /root/reference/src).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Inputs are regenerated from gaselect_b200.synth on the consuming side; only parameters and
the reference's outputs are stored.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gaselect_b200 import synth  # noqa: E402
from oracle.bindings import RefLib  # noqa: E402

CASES = [
    # name, data spec, evaluator spec, subset spec
    dict(name="closed_rdcv", data=dict(kind="closed_form", n=40, p=30),
         evaluator=dict(kind="pls", numReplications=2, maxNComp=5, innerSegments=4, outerSegments=3),
         subsets=dict(kind="list", lists=[[2, 5, 11, 29], list(range(12)), [7], [1, 2, 5, 8, 11, 13, 17, 19, 23, 29], []])),
    dict(name="closed_srcv_penalty", data=dict(kind="closed_form", n=40, p=30),
         evaluator=dict(kind="pls", numReplications=3, maxNComp=0, innerSegments=7, outerSegments=1, complexityPenalty=0.01),
         subsets=dict(kind="list", lists=[[2, 5, 11, 29], list(range(12)), [0], list(range(0, 30, 2))])),
    dict(name="closed_srcv_testsize", data=dict(kind="closed_form", n=40, p=30),
         evaluator=dict(kind="pls", numReplications=2, maxNComp=0, innerSegments=7, outerSegments=1, testSetSize=0.4),
         subsets=dict(kind="list", lists=[[2, 5, 11, 29], [3, 4, 9]])),
    dict(name="closed_rdcv_ragged", data=dict(kind="closed_form", n=43, p=30),
         evaluator=dict(kind="pls", numReplications=3, maxNComp=4, innerSegments=3, outerSegments=4, sdfact=0.5),
         subsets=dict(kind="random", count=12, kmin=1, kmax=20, seed=11)),
    dict(name="cfg1_gauss_srcv", data=dict(kind="gaussian", n=100, p=200, seed=777),
         evaluator=dict(kind="pls"),
         subsets=dict(kind="random", count=16, kmin=5, kmax=30, seed=7)),
    dict(name="cfg3_nir_rdcv", data=dict(kind="nir", n=150, p=1000, seed=777, noise=1e-2),
         evaluator=dict(kind="pls", numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5),
         subsets=dict(kind="random", count=40, kmin=5, kmax=200, seed=5)),
    dict(name="nir_small_rdcv", data=dict(kind="nir", n=60, p=120, seed=3, noise=1e-2),
         evaluator=dict(kind="pls", numReplications=4, maxNComp=6, innerSegments=3, outerSegments=4),
         subsets=dict(kind="random", count=24, kmin=1, kmax=40, seed=9)),
    dict(name="closed_fit_bic", data=dict(kind="closed_form", n=40, p=30),
         evaluator=dict(kind="fit", numSegments=7, statistic="BIC"),
         subsets=dict(kind="list", lists=[[2, 5, 11, 29], list(range(12)), [7]])),
    dict(name="closed_fit_aic", data=dict(kind="closed_form", n=40, p=30),
         evaluator=dict(kind="fit", numSegments=7, statistic="AIC"),
         subsets=dict(kind="list", lists=[[2, 5, 11, 29]])),
    dict(name="closed_fit_adjr2", data=dict(kind="closed_form", n=40, p=30),
         evaluator=dict(kind="fit", numSegments=7, statistic="adjusted.r.squared"),
         subsets=dict(kind="list", lists=[[2, 5, 11, 29]])),
    dict(name="closed_fit_r2", data=dict(kind="closed_form", n=40, p=30),
         evaluator=dict(kind="fit", numSegments=5, statistic="r.squared", maxNComp=3),
         subsets=dict(kind="list", lists=[[2, 5, 11, 29], [0, 1, 2, 3, 4, 5]])),
    dict(name="nir_fit_bic", data=dict(kind="nir", n=150, p=400, seed=21, noise=1e-2),
         evaluator=dict(kind="fit", numSegments=7, statistic="BIC", maxNComp=8),
         subsets=dict(kind="random", count=16, kmin=3, kmax=60, seed=13)),
    dict(name="closed_lm_bic", data=dict(kind="closed_form", n=40, p=30),
         evaluator=dict(kind="lm", statistic="BIC"),
         subsets=dict(kind="list", lists=[[2, 5, 11, 29], [1, 2, 5, 8, 11, 13, 17, 19, 23, 29]])),
    dict(name="closed_lm_aic", data=dict(kind="closed_form", n=40, p=30),
         evaluator=dict(kind="lm", statistic="AIC"), subsets=dict(kind="list", lists=[[2, 5, 11, 29]])),
    dict(name="closed_lm_adjr2", data=dict(kind="closed_form", n=40, p=30),
         evaluator=dict(kind="lm", statistic="adjusted.r.squared"), subsets=dict(kind="list", lists=[[2, 5, 11, 29]])),
    dict(name="closed_lm_r2", data=dict(kind="closed_form", n=40, p=30),
         evaluator=dict(kind="lm", statistic="r.squared"), subsets=dict(kind="list", lists=[[2, 5, 11, 29]])),
    dict(name="cfg2_gauss_lm", data=dict(kind="gaussian", n=200, p=500, seed=777),
         evaluator=dict(kind="lm", statistic="BIC"),
         subsets=dict(kind="random", count=32, kmin=5, kmax=50, seed=6)),
]


def make_data(spec):
    spec = dict(spec)
    kind = spec.pop("kind")
    return {"closed_form": synth.closed_form, "gaussian": synth.gaussian, "nir": synth.nir_spectra}[kind](**spec)


def make_subsets(spec, p):
    if spec["kind"] == "list":
        return [list(map(int, s)) for s in spec["lists"]]
    return [list(map(int, s)) for s in synth.random_subsets(p, spec["count"], spec["kmin"], spec["kmax"], spec["seed"])]


def make_evaluator(lib, spec, X, y, seedvec):
    spec = dict(spec)
    kind = spec.pop("kind")
    if kind == "pls":
        return lib.pls_evaluator(X, y, seedvec, **spec)
    if kind == "fit":
        return lib.fit_evaluator(X, y, seedvec, **spec)
    return lib.lm_evaluator(X, y, **spec)


def main():
    lib = RefLib()
    out = {"generator": "tests/golden/make_golden.py via oracle/_ref (reference sources, stand-in linear algebra)",
           "ga_seed": 123}
    sv = lib.make_seedvec(123)
    out["rng"] = {
        "u32_first8": {str(s): lib.rng_u32(s, 8).tolist() for s in (123, 1, 0, 4294967295)},
        "xor_checksum_2000": {str(s): int(np.bitwise_xor.reduce((lib.rng_u32(s, 2000).astype(np.uint64) + np.arange(2000, dtype=np.uint64)) & np.uint64(0xFFFFFFFF))) for s in (123, 1, 0, 4294967295)},
        "seedvec123_head": sv[:4].tolist(), "seedvec123_tail": int(sv[623]),
        "seedvec123_xor": int(np.bitwise_xor.reduce(sv)),
        "vec_first8": lib.rng_vec(sv, 8).tolist(),
        "vec_uniform_0_150": [float(v) for v in lib.rng_uniform(sv, 0, 150, 6)],
    }
    out["shuffle40x3"] = lib.shuffle(sv, 40, 3).tolist()
    out["shuffle150x2_xor"] = [int(np.bitwise_xor.reduce(r * np.arange(1, 151, dtype=np.uint32))) for r in lib.shuffle(sv, 150, 2)]
    cases = []
    for case in CASES:
        X, y = make_data(case["data"])
        ev = make_evaluator(lib, case["evaluator"], X, y, sv)
        subsets = make_subsets(case["subsets"], X.shape[1])
        fit, status = ev.evaluate(subsets)
        seg = ev.segmentation()
        rec = dict(case)
        rec["fitness"] = [float(v) for v in fit]
        rec["status"] = [int(v) for v in status]
        rec["segmentation_count"] = len(seg)
        # full vectors for small problems, a digest for large ones
        if X.shape[0] <= 64:
            rec["segmentation"] = [s.tolist() for s in seg]
        rec["segmentation_digest"] = [int((np.asarray(s, dtype=np.uint64) * np.arange(1, len(s) + 1, dtype=np.uint64)).sum() % (1 << 61)) for s in seg]
        rec["data_probe"] = [float(X[0, 0]), float(X[-1, -1]), float(y[0]), float(y[-1])]
        cases.append(rec)
        print(case["name"], len(subsets), "subsets", len(seg), "segments")
    out["cases"] = cases

    # raw SIMPLS (the hidden C_simpls entry)
    X, y = synth.closed_form()
    B, b0 = lib.simpls_fit(X, y, [2, 5, 11, 29], 3)
    out["simpls_closed"] = {"cols": [2, 5, 11, 29], "ncomp": 3, "coef": B.T.tolist(), "intercepts": b0.tolist()}

    # host GA pieces and full GA runs
    kats = []
    for p, mp, co in ((30, 0.01, 0), (100, 0.01, 0), (30, 0.3, 1), (100, 0.3, 1)):
        lists, mutated, nxt = lib.chromosome_kat(p, sv, mutationProb=mp, crossover=co)
        kats.append(dict(p=p, mutationProb=mp, crossover=co, lists=lists, mutated=mutated, next_draw=nxt))
    out["chromosome_kat"] = kats
    ev = lib.pls_evaluator(X, y, sv, numReplications=2, maxNComp=5, innerSegments=4, outerSegments=3)
    runs = []
    for T in (1, 4):
        r = lib.ga_run(ev, sv, 30, 40, 8, elitism=5, minVariables=3, maxVariables=10, numThreads=T)
        runs.append(dict(numThreads=T, best=[float(v) for v in r["fitnessEvolution"][:, 0]],
                         evolution=r["fitnessEvolution"].tolist(),
                         subsets=[s.tolist() for s in r["subsets"]], fitness=[float(v) for v in r["fitness"]],
                         evaluations=r["evaluations"]))
    out["ga_runs"] = runs

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
