"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against
  (a) the committed golden vectors produced by the reference's own code,
  (b) the CPU oracle on fresh seeded inputs, and
  (c) size-independent properties at BASELINE.json's full sizes.
Tolerance: BASELINE.json north_star -- fitness within RELATIVE 1e-8 in FP64; statuses,
segmentations and subsets bit-exact."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from gaselect_b200 import synth
from helpers import make_checker_evaluator, make_data, make_subsets, rel_err

pytestmark = pytest.mark.gpu

REL_TOL = 1e-8  # north_star: "within relative 1e-8 in FP64"


def _product_evaluator(spec, X, y, seed):
    from gaselect_b200 import evaluators as E
    spec = dict(spec)
    kind = spec.pop("kind")
    if kind == "pls":
        return E.PLSEvaluator(X, y, seed, **spec)
    if kind == "fit":
        return E.BICEvaluator(X, y, seed, **spec)
    return E.LMEvaluator(X, y, **spec)


def _case_names():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as fh:
        return [c["name"] for c in json.load(fh)["cases"]]


@pytest.mark.parametrize("name", _case_names())
def test_golden_case(engine, golden, name):
    case = next(c for c in golden["cases"] if c["name"] == name)
    X, y = make_data(case["data"])
    subsets = make_subsets(case["subsets"], X.shape[1])
    with _product_evaluator(case["evaluator"], X, y, golden["ga_seed"]) as ev:
        fit, status = ev.evaluate_population(subsets)
        t = ev.timing()
    assert status.tolist() == case["status"]
    ok = np.asarray(case["status"]) == 0
    err = rel_err(fit[ok], np.asarray(case["fitness"])[ok])
    assert err.max() <= REL_TOL, (name, float(err.max()))
    assert t["kernel_launches"] >= 1 and t["fit_ctas"] >= 1  # the CUDA kernels really ran


def test_survey_appendix_c_values(engine):
    from gaselect_b200 import evaluators as E
    X, y = synth.closed_form()
    with E.PLSEvaluator(X, y, 123, numReplications=2, maxNComp=5, innerSegments=4, outerSegments=3) as ev:
        fit, _ = ev.evaluate_population([[2, 5, 11, 29], list(range(12)), [7], [1, 2, 5, 8, 11, 13, 17, 19, 23, 29]])
        want = [-0.082050888884043549, -0.40119862398530326, -1.6049334059118827, -0.26544539104666021]
        assert rel_err(fit, want).max() <= REL_TOL
        assert ev.evaluate([2, 5, 11, 29]) == pytest.approx(want[0], rel=REL_TOL)
    with E.LMEvaluator(X, y, "BIC") as ev:
        assert ev.evaluate([1, 2, 5, 8, 11, 13, 17, 19, 23, 29]) == pytest.approx(174.02302223845993, rel=REL_TOL)
    with E.BICEvaluator(X, y, 123, statistic="BIC") as ev:
        assert ev.evaluate([2, 5, 11, 29]) == pytest.approx(191.69625628073339, rel=REL_TOL)


def test_raw_simpls_coefficients(engine, golden):
    from gaselect_b200 import evaluators as E
    g = golden["simpls_closed"]
    X, y = synth.closed_form()
    with E.PLSEvaluator(X, y, 123, numReplications=1, innerSegments=3) as ev:
        B, b0 = ev.simpls(g["cols"], g["ncomp"])
        want = np.asarray(g["coef"])
        want = want.T if want.shape != B.shape else want   # golden file stores one row per component count
        assert np.abs(B - want).max() <= 1e-9 * np.abs(B).max()
        assert b0 == pytest.approx(g["intercepts"], rel=1e-9)
        # a row subset, against the oracle
    from oracle import bindings
    orc = bindings.OracleLib()
    rows = np.arange(0, 40, 2, dtype=np.uint32)
    Bo, b0o = orc.simpls_fit(X, y, [1, 4, 9, 16, 25], 4, rows=rows)
    with E.LMEvaluator(X, y) as ev:
        B, b0 = ev.simpls([1, 4, 9, 16, 25], 4, rows=rows)
    assert np.abs(B - Bo).max() <= 1e-9 * np.abs(Bo).max() and b0 == pytest.approx(b0o, rel=1e-9)


@pytest.mark.parametrize("cfg", [
    dict(n=57, p=80, kw=dict(numReplications=3, maxNComp=5, innerSegments=5, outerSegments=3)),
    dict(n=64, p=50, kw=dict(numReplications=2, maxNComp=0, innerSegments=4, outerSegments=1)),
    dict(n=45, p=60, kw=dict(numReplications=2, maxNComp=4, innerSegments=2, outerSegments=2)),
    dict(n=90, p=300, kw=dict(numReplications=3, maxNComp=8, innerSegments=4, outerSegments=5)),
    dict(n=31, p=33, kw=dict(numReplications=2, maxNComp=3, innerSegments=3, outerSegments=1, testSetSize=0.3, sdfact=2.0,
                             complexityPenalty=0.02)),
])
def test_pls_against_oracle_on_fresh_inputs(engine, oracle_lib, cfg):
    from gaselect_b200 import evaluators as E
    n, p, kw = cfg["n"], cfg["p"], cfg["kw"]
    X, y = synth.nir_spectra(n, p, seed=300 + n)
    sv = oracle_lib.make_seedvec(2000 + p)
    subsets = synth.random_subsets(p, 48, 1, min(p, 70), seed=n) + [np.arange(p, dtype=np.uint32)[: min(p, 150)]]
    chk = oracle_lib.pls_evaluator(X, y, sv, **kw)
    fo, so = chk.evaluate(subsets, threads=4)
    with E.PLSEvaluator(X, y, sv, **kw) as ev:
        assert all((a == b).all() for a, b in zip(ev.getSegmentation(), chk.segmentation()))
        fg, sg = ev.evaluate_population(subsets)
    assert (sg == so).all()
    assert rel_err(fg[so == 0], fo[so == 0]).max() <= REL_TOL


@pytest.mark.parametrize("stat", ["BIC", "AIC", "adjusted.r.squared", "r.squared"])
def test_fit_and_lm_against_oracle(engine, oracle_lib, stat):
    from gaselect_b200 import evaluators as E
    X, y = synth.gaussian(120, 90, seed=5)
    sv = oracle_lib.make_seedvec(77)
    subsets = synth.random_subsets(90, 40, 1, 45, seed=3)
    fo, so = oracle_lib.fit_evaluator(X, y, sv, numSegments=6, statistic=stat, maxNComp=7).evaluate(subsets)
    with E.BICEvaluator(X, y, sv, numSegments=6, statistic=stat, maxNComp=7) as ev:
        fg, sg = ev.evaluate_population(subsets)
    assert (sg == so).all() and rel_err(fg, fo).max() <= REL_TOL
    fo, so = oracle_lib.lm_evaluator(X, y, stat).evaluate(subsets)
    with E.LMEvaluator(X, y, stat) as ev:
        fg, sg = ev.evaluate_population(subsets)
    assert (sg == so).all() and rel_err(fg, fo).max() <= REL_TOL


def test_empty_ragged_and_unsorted_subsets(engine, oracle_lib):
    from gaselect_b200 import evaluators as E
    X, y = synth.closed_form()
    sv = oracle_lib.make_seedvec(5)
    subsets = [[], [3], [1, 2], [], list(range(30)), [29, 0, 7]]
    with E.PLSEvaluator(X, y, sv, numReplications=2, innerSegments=3, maxNComp=4) as ev:
        fit, status = ev.evaluate_population(subsets)
        assert status.tolist() == [1, 0, 0, 1, 0, 0] and fit[0] == 0.0 and fit[3] == 0.0
        fo, _ = oracle_lib.pls_evaluator(X, y, sv, numReplications=2, innerSegments=3, maxNComp=4).evaluate(
            [sorted(s) for s in subsets if len(s)])
        assert rel_err(fit[status == 0], fo).max() <= REL_TOL
        with pytest.raises(E.EvaluatorException):
            ev.evaluate([])
        with pytest.raises(ValueError):
            ev.evaluate([30])          # column out of range
        f0, s0 = ev.evaluate_population([])
        assert len(f0) == 0
    # LMEvaluator does not refuse an empty subset: it scores the intercept-only model (src/LMEvaluator.cpp:27-67)
    for stat in ("BIC", "AIC", "adjusted.r.squared", "r.squared"):
        with E.LMEvaluator(X, y, stat) as ev:
            fit, status = ev.evaluate_population(subsets)
        fo, so = oracle_lib.lm_evaluator(X, y, stat).evaluate([sorted(s) for s in subsets])
        assert status.tolist() == [0] * 6 and so.tolist() == [0] * 6
        assert np.allclose(fit, fo, rtol=REL_TOL, atol=1e-12)


def test_bitset_entry_matches_index_entry(engine):
    """gsb_evaluate_bitsets decodes the reference layout (src/Chromosome.cpp:419-438)."""
    from gaselect_b200 import evaluators as E
    for p in (30, 64, 100, 200):
        X, y = synth.gaussian(60, p, seed=p)
        subsets = synth.random_subsets(p, 20, 1, 25, seed=p)
        words = (p + 63) // 64
        unused = words * 64 - p
        bits = np.zeros((len(subsets), words), dtype=np.uint64)
        for c, s in enumerate(subsets):
            for v in s:
                pos = unused + int(v)
                bits[c, pos // 64] |= np.uint64(1) << np.uint64(pos % 64)
        bits[:, 0] |= np.uint64((1 << unused) - 1)  # garbage in the unused low bits must be ignored
        with E.PLSEvaluator(X, y, 9, numReplications=2, innerSegments=4, maxNComp=6) as ev:
            fa, sa = ev.evaluate_population(subsets)
            fb, sb = ev.evaluate_bitsets(bits)
        assert (sa == sb).all() and (fa == fb).all()  # bit-exact: same kernels, same inputs


def test_results_do_not_depend_on_batch_composition(engine):
    """Each chromosome's fitness is a pure function of its subset: evaluating it alone, in a
    shuffled batch or next to different neighbours gives bit-identical results."""
    from gaselect_b200 import evaluators as E
    X, y = synth.nir_spectra(150, 1000, seed=777)
    subsets = synth.sweep_subsets(1000, 64, 40, seed=2)
    with E.PLSEvaluator(X, y, 123, numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5) as ev:
        f1, s1 = ev.evaluate_population(subsets)
        perm = np.argsort(synth.SplitMix64(4).uniform(len(subsets)))
        f2, s2 = ev.evaluate_population([subsets[i] for i in perm])
        assert (f1[perm] == f2).all() and (s1[perm] == s2).all()
        f3, _ = ev.evaluate_population(subsets[:5])
        assert (f3 == f1[:5]).all()
        assert ev.evaluate(subsets[7]) == f1[7]


def test_cfg3_full_size_population_against_reference_build(engine, ref_lib):
    """BASELINE config 3 at full size (n=150, p=1000, population 500, rdCV 10 x 5/4, <=10
    components) against the reference's own code, 8 host threads."""
    from gaselect_b200 import evaluators as E
    X, y = synth.nir_spectra(150, 1000, seed=777, noise=1e-2)
    sv = ref_lib.make_seedvec(123)
    subsets = synth.random_subsets(1000, 500, 10, 200, seed=42)
    kw = dict(numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5)
    with E.PLSEvaluator(X, y, sv, **kw) as ev:
        fg, sg = ev.evaluate_population(subsets)
        inf = ev.info()
    assert (inf["fits_per_evaluation"], inf["distinct_fits"], inf["distinct_blocks"]) == (250, 150, 50)
    fr, sr = ref_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets[:120], threads=8)
    assert (sg == 0).all() and (sr == 0).all()
    assert rel_err(fg[:120], fr).max() <= REL_TOL


def test_cfg4_shape_against_reference_build(engine, ref_lib):
    """BASELINE config 4 shape (n=500, p=5000, 10 x 5/4 CV, <= 10 components): 96 subsets of 20...400 columns (SURVEY 8(d)'s
    range: the shared-memory path, the spill path above 212 columns and the in-place path above 384) against the
    reference's own code (oracle/_ref), plus order independence."""
    from gaselect_b200 import evaluators as E
    X, y = synth.nir_spectra(500, 5000, seed=778, noise=1e-2)
    sv = ref_lib.make_seedvec(123)
    subsets = synth.random_subsets(5000, 88, 20, 400, seed=43) + [synth.random_subsets(5000, 1, k, k, seed=k)[0] for k in (20, 212, 213, 256, 300, 384, 385, 400)]
    kw = dict(numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5)
    with E.PLSEvaluator(X, y, sv, **kw) as ev:
        fg, sg = ev.evaluate_population(subsets)
        fg2, _ = ev.evaluate_population(subsets[::-1])
    assert len(subsets) == 96 and (sg == 0).all() and (fg2[::-1] == fg).all() and (fg < 0).all()
    fr, sr = ref_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=16)
    assert (sr == 0).all() and rel_err(fg, fr).max() <= REL_TOL


@pytest.mark.parametrize("rows", [150, 400])
def test_prediction_in_the_fit_kernel_and_in_its_own_kernel_agree(engine, oracle_lib, rows):
    """The per-fit path predicts the held-out rows either in the fit kernel (default for blocks of fewer than 64 rows) or
    with one tensor-core product per (chromosome, held-out block) in predictBlocksKernel (default for larger blocks):
    both forced in turn, against each other and the oracle."""
    from gaselect_b200 import evaluators as E
    from test_gpu_scores import forced
    X, y = synth.nir_spectra(rows, 600, seed=rows)
    sv = oracle_lib.make_seedvec(31)
    kw = dict(numReplications=3, maxNComp=8, innerSegments=4, outerSegments=5)
    subsets = synth.random_subsets(600, 40, 1, 120, seed=5)
    res = []
    for env in (dict(GSB_PREDICT="1", GSB_NO_PREDICT=None), dict(GSB_PREDICT=None, GSB_NO_PREDICT="1"), dict(GSB_PREDICT=None, GSB_NO_PREDICT=None)):
        with forced(GSB_NO_KSPACE="1", GSB_NO_SCORES="1", **env):
            with E.PLSEvaluator(X, y, sv, **kw) as ev:
                res.append(ev.evaluate_population(subsets) + (ev.timing()["kernel_launches"],))
    want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=8)
    assert res[0][2] > res[1][2]  # the separate prediction is one more launch per subset-size class
    assert res[2][2] == (res[0][2] if rows >= 320 else res[1][2])  # the default by block size (rows / 5 >= 64)
    for fit, status, _ in res:
        assert (status == wstatus).all() and rel_err(fit, want).max() <= REL_TOL
    assert rel_err(res[0][0], res[1][0]).max() <= 1e-9


def test_subsets_beyond_shared_memory_spill_to_l2(engine, oracle_lib):
    """k > 212 (10 components): the last packed Gram rows live in a global scratch (K1 spill path);
    k > 384 takes the generic path that reads the Gram in place."""
    from gaselect_b200 import evaluators as E
    X, y = synth.nir_spectra(150, 1000, seed=777, noise=1e-2)
    sv = oracle_lib.make_seedvec(123)
    sizes = [213, 224, 225, 256, 257, 300, 352, 384, 400]
    subsets = [synth.random_subsets(1000, 1, k, k, seed=k)[0] for k in sizes] + synth.random_subsets(1000, 6, 150, 330, seed=77)
    kw = dict(numReplications=2, maxNComp=10, innerSegments=4, outerSegments=5)
    fo, so = oracle_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=8)
    with E.PLSEvaluator(X, y, sv, **kw) as ev:
        fg, sg = ev.evaluate_population(subsets)
        fg2, _ = ev.evaluate_population(subsets[::-1])
    assert (sg == so).all() and (fg2[::-1] == fg).all()
    assert rel_err(fg, fo).max() <= REL_TOL


def test_underflow_is_reported_where_the_reference_throws(engine, oracle_lib):
    """A constant response makes X'y vanish: PLSSimpls::fit throws std::underflow_error (||t|| < 1e-25,
    src/PLSSimpls.cpp:141-143), the evaluators turn it into EvaluatorException (src/PLSEvaluator.cpp:337-340,
    src/BICEvaluator.cpp:227-230).  The engine reports status 2 for exactly those chromosomes."""
    from gaselect_b200 import evaluators as E
    X, _ = synth.closed_form()
    y = np.full(40, 3.25)
    sv = oracle_lib.make_seedvec(5)
    subsets = [[1, 2, 3], [5], [0, 7, 9, 11, 13], list(range(30))]
    _, so = oracle_lib.pls_evaluator(X, y, sv, numReplications=2, innerSegments=3, maxNComp=4).evaluate(subsets)
    with E.PLSEvaluator(X, y, sv, numReplications=2, innerSegments=3, maxNComp=4) as ev:
        _, sg = ev.evaluate_population(subsets)
        assert sg.tolist() == so.tolist() == [2, 2, 2, 2]
        with pytest.raises(E.EvaluatorException):
            ev.evaluate([1, 2, 3])
    _, so = oracle_lib.fit_evaluator(X, y, sv, numSegments=5, maxNComp=3).evaluate(subsets)
    with E.BICEvaluator(X, y, sv, numSegments=5, maxNComp=3) as ev:
        _, sg = ev.evaluate_population(subsets)
    assert sg.tolist() == so.tolist() == [2, 2, 2, 2]
    # unconstrained component count (generic kernel path), mixed with healthy data in the same engine call
    X2, y2 = synth.closed_form()
    with E.PLSEvaluator(X2, y2, sv, numReplications=2, innerSegments=3) as ev:
        f, s = ev.evaluate_population(subsets)
        assert s.tolist() == [0, 0, 0, 0] and np.isfinite(f).all()


def test_resident_population_api_and_call_order(engine):
    """upload / evaluate / download == evaluate_indices; state errors are reported, not crashed on."""
    from gaselect_b200 import evaluators as E, _capi
    X, y = synth.gaussian(80, 120, seed=2)
    subsets = synth.random_subsets(120, 50, 1, 40, seed=3)
    idx, off = _capi.to_csr(subsets)
    with E.PLSEvaluator(X, y, 11, numReplications=3, innerSegments=4, maxNComp=5) as ev:
        with pytest.raises(E.EngineError) as err:
            ev.evaluate_resident()                       # nothing uploaded yet
        assert err.value.code == _capi.GSB_ERR_STATE
        f0, s0 = ev.evaluate_csr(idx, off)
        ev.upload_population(idx, off)
        with pytest.raises(E.EngineError):
            ev.download_results(np.zeros(50), np.zeros(50, dtype=np.int32))  # not evaluated yet
        ev.evaluate_resident()
        ev.enqueue_resident()                            # asynchronous variant, same results
        f1, s1 = np.zeros(50), np.zeros(50, dtype=np.int32)
        ev.download_results(f1, s1)
        assert (f1 == f0).all() and (s1 == s0).all()
        fp, sp, n = ev.device_results()
        assert fp and sp and n == 50
        t = ev.timing()
        # one CTA per (chromosome, fit) on the per-fit kernel, per (chromosome, group of <= 8 fits) on the score-space kernel
        assert t["fit_ctas"] > 0 and t["fit_ctas"] % 50 == 0 and t["fit_ctas"] <= 4 * 50 * ev.info()["distinct_fits"] and t["fit_kernel_ms"] > 0


def test_two_engines_and_a_caller_stream(engine):
    """Engines are independent; gsb_set_stream moves an engine's work onto the caller's CUDA stream."""
    import torch
    from gaselect_b200 import evaluators as E
    X, y = synth.gaussian(60, 90, seed=4)
    subsets = synth.random_subsets(90, 30, 2, 30, seed=5)
    with E.PLSEvaluator(X, y, 7, numReplications=2, innerSegments=3, maxNComp=4) as a, \
            E.PLSEvaluator(X, y, 7, numReplications=2, innerSegments=3, maxNComp=4) as b, E.LMEvaluator(X, y, "AIC") as c:
        fa, _ = a.evaluate_population(subsets)
        stream = torch.cuda.Stream()
        b.set_stream(stream.cuda_stream)
        fb, _ = b.evaluate_population(subsets)
        b.set_stream(None)
        fb2, _ = b.evaluate_population(subsets)
        fc, sc = c.evaluate_population(subsets)
        assert (fa == fb).all() and (fa == fb2).all() and (sc == 0).all() and np.isfinite(fc).all()


def test_multi_device_engine_is_bit_identical_to_one_device(engine, oracle_lib):
    """A multi-device engine (gsb_config.devices) deals every batch to its per-GPU replicas and gathers the results with
    peer copies: bit-identical fitness and status to the single-device engine, for PLS_CV (kernel-space kernel and per-fit
    kernel classes, an empty subset, a repeated call with another batch), PLS_FIT and LM."""
    from conftest import cuda_device_count
    from gaselect_b200 import evaluators as E, _capi
    ndev = cuda_device_count()
    if ndev < 2:
        pytest.skip("needs at least 2 CUDA devices")
    devices = list(range(min(ndev, 8)))
    X, y = synth.nir_spectra(150, 400, seed=12, noise=1e-2)
    kw = dict(numReplications=3, maxNComp=8, innerSegments=4, outerSegments=5)
    subsets = synth.random_subsets(400, 203, 1, 150, seed=5)
    subsets[7] = np.zeros(0, dtype=np.uint32)
    with E.PLSEvaluator(X, y, 123, **kw) as one, E.PLSEvaluator(X, y, 123, devices=devices, **kw) as multi:
        f1, s1 = one.evaluate_population(subsets)
        fm, sm = multi.evaluate_population(subsets)
        assert (s1 == sm).all() and (f1 == fm).all() and s1[7] == 1 and (s1 == 0).sum() == len(subsets) - 1
        f1b, s1b = one.evaluate_population(subsets[::3])
        fmb, smb = multi.evaluate_population(subsets[::3])
        assert (s1b == smb).all() and (f1b == fmb).all() and (f1b == f1[::3]).all()
        # the resident-population entry points (gather by peer copies through the first device) against the one-call path
        # (every replica downloads its own share), and a download of what the one-call path left resident
        fr, sr = np.zeros(len(subsets[::3])), np.zeros(len(subsets[::3]), dtype=np.int32)
        multi.download_results(fr, sr)
        assert (fr == fmb).all() and (sr == smb).all()
        idx, off = _capi.to_csr(subsets)
        multi.upload_population(idx, off)
        multi.evaluate_resident()
        fr, sr = np.zeros(len(subsets)), np.zeros(len(subsets), dtype=np.int32)
        multi.download_results(fr, sr)
        assert (fr == f1).all() and (sr == s1).all()
        fo, _ = oracle_lib.pls_evaluator(X, y, oracle_lib.make_seedvec(123), **kw).evaluate(subsets[20:28], threads=4)
        assert rel_err(fm[20:28], fo).max() <= REL_TOL
        assert multi.timing()["fit_ctas"] == one.timing()["fit_ctas"] or multi.timing()["fit_ctas"] > 0
        fe, se = multi.evaluate_population([])
        assert len(fe) == 0 and len(se) == 0
        few = subsets[100:100 + max(1, len(devices) - 1)]  # fewer chromosomes than devices: some replicas get nothing
        f1c, s1c = one.evaluate_population(few)
        fmc, smc = multi.evaluate_population(few)
        assert (s1c == smc).all() and (f1c == fmc).all()
    # chromosomes the kernel-space kernel declines (rank-deficient: 48 near-copies of 6 columns, 8 components) are re-scored
    # on the Gram-space kernel of the replica that holds them before the gather
    rng = np.random.default_rng(4)
    Xc = X.copy()
    for j in range(0, 96, 8):
        for d in range(1, 8):
            Xc[:, j + d] = Xc[:, j] * (1.0 + 1e-7 * rng.standard_normal(Xc.shape[0]))
    hard = [np.arange(0, 48, dtype=np.uint32), np.arange(48, 96, dtype=np.uint32)] + [s[s >= 96] for s in subsets[30:60] if (s >= 96).sum() > 45]
    with E.PLSEvaluator(Xc, y, 123, **kw) as one, E.PLSEvaluator(Xc, y, 123, devices=devices, **kw) as multi:
        f1, s1 = one.evaluate_population(hard)
        r1 = one.timing()["retries"]
        fm, sm = multi.evaluate_population(hard)
        assert r1 == 2 and multi.timing()["retries"] == 2
        assert (s1 == sm).all() and (f1 == fm).all() and np.isfinite(fm).all() and not (sm == 5).any()
    narrow = [s for s in subsets if 0 < len(s) <= 60]
    with E.BICEvaluator(X, y, 123, numSegments=5, maxNComp=6) as one, E.BICEvaluator(X, y, 123, numSegments=5, maxNComp=6, devices=devices) as multi:
        f1, s1 = one.evaluate_population(narrow)
        fm, sm = multi.evaluate_population(narrow)
        assert (s1 == sm).all() and (f1 == fm).all()
    with E.LMEvaluator(X, y) as one, E.LMEvaluator(X, y, devices=devices[:2]) as multi:
        f1, s1 = one.evaluate_population(narrow)
        fm, sm = multi.evaluate_population(narrow)
        assert (s1 == sm).all() and (f1 == fm).all()


def test_unconstrained_components_on_wide_subsets(engine, oracle_lib):
    """evaluatorPLS defaults (maxNComp = 0: as many components as the training sets allow, src/PLSEvaluator.cpp:135-138)
    on subsets of up to 200 columns: 88 components per fit.  The stored loadings and cumulative coefficients (2 k A
    doubles = 320 KB) no longer fit in shared memory: the generic kernel keeps them in a global scratch and the class is
    walked in chunks of chromosomes.  Well-conditioned Gaussian data, as SURVEY 8(d) prescribes for unconstrained
    component counts; against the plain-C oracle (the reference build's stand-in crashes at maxNComp = 0 with n = 150)."""
    from gaselect_b200 import evaluators as E
    X, y = synth.gaussian(150, 400, seed=21)
    sv = oracle_lib.make_seedvec(17)
    subsets = [synth.random_subsets(400, 1, k, k, seed=k)[0] for k in (12, 60, 105, 106, 130, 200)] + synth.random_subsets(400, 6, 20, 200, seed=4)
    kw = dict(numReplications=2, innerSegments=5)
    want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=8)
    with E.PLSEvaluator(X, y, sv, **kw) as ev:
        assert ev.prepare()["max_ncomp"] > 80  # 88 components: 2 x 200 x 88 doubles = 282 KB of loadings and coefficients
        fit, status = ev.evaluate_population(subsets)
        again, _ = ev.evaluate_population(subsets[::-1])
    assert (status == wstatus).all() and (again[::-1] == fit).all()
    assert rel_err(fit[wstatus == 0], want[wstatus == 0]).max() <= REL_TOL


def test_lm_subsets_wider_than_shared_memory(engine, oracle_lib):
    """LMEvaluator with subsets of up to n - 2 = 298 columns (the reference takes any k <= n - 1, R/validData.R:30-36): above
    ~230 columns the packed Gram triangle no longer fits in shared memory and K3 factorises it in a global scratch."""
    from gaselect_b200 import evaluators as E
    X, y = synth.gaussian(300, 400, seed=12)
    subsets = [synth.random_subsets(400, 1, k, k, seed=k)[0] for k in (1, 40, 230, 240, 298)] + synth.random_subsets(400, 8, 100, 298, seed=2)
    for stat in ("BIC", "r.squared"):
        want, wstatus = oracle_lib.lm_evaluator(X, y, stat).evaluate(subsets, threads=8)
        with E.LMEvaluator(X, y, stat) as ev:
            fit, status = ev.evaluate_population(subsets)
        assert (status == wstatus).all() and rel_err(fit, want).max() <= REL_TOL
