"""CPU tests: the oracle restatement against (a) the committed golden vectors produced by the
reference's own code (tests/golden/make_golden.py) and SURVEY.md Appendix C, and (b) live
cross-execution with oracle/_ref where that library is present."""
import numpy as np
import pytest

from gaselect_b200 import synth
from helpers import make_checker_evaluator, make_data, make_subsets, rel_err, seg_digest

FIT_TOL = 1e-10  # oracle vs reference build: both X-space FP64, only summation order differs


def test_rng_known_answers(oracle_lib, golden):
    g = golden["rng"]
    for seed, first in g["u32_first8"].items():
        assert oracle_lib.rng_u32(int(seed), 8).tolist() == first
    for seed, want in g["xor_checksum_2000"].items():
        d = oracle_lib.rng_u32(int(seed), 2000).astype(np.uint64)
        got = int(np.bitwise_xor.reduce((d + np.arange(2000, dtype=np.uint64)) & np.uint64(0xFFFFFFFF)))
        assert got == want  # 2000 draws exercise all six index cases of src/RNG.cpp:92-167
    sv = oracle_lib.make_seedvec(golden["ga_seed"])
    assert sv[:4].tolist() == g["seedvec123_head"] and int(sv[623]) == g["seedvec123_tail"]
    assert int(np.bitwise_xor.reduce(sv)) == g["seedvec123_xor"]
    assert oracle_lib.rng_vec(sv, 8).tolist() == g["vec_first8"]
    assert [float(v) for v in oracle_lib.rng_uniform(sv, 0, 150, 6)] == g["vec_uniform_0_150"]


def test_rng_survey_appendix_c1(oracle_lib):
    # literal values of SURVEY.md Appendix C.1 (produced by the unmodified src/RNG.cpp)
    assert oracle_lib.rng_u32(123, 6).tolist() == [2866969740, 3565251014, 4277846663, 460894080, 3325959662, 3310279934]
    assert oracle_lib.rng_u32(1, 6).tolist() == [4223579984, 48694833, 3264091594, 3536251047, 684290322, 2888275087]
    assert oracle_lib.rng_u32(0, 6).tolist() == [3664475182, 1448425114, 855785591, 2204380445, 3148018086, 3849643884]
    assert oracle_lib.rng_u32(4294967295, 6).tolist() == [2231073311, 2374912395, 1193726977, 3718570002, 3815165155, 1302313332]
    sv = oracle_lib.make_seedvec(123)
    assert oracle_lib.rng_vec(sv, 6).tolist() == [2738463805, 2572576382, 2132272039, 2119152319, 2475546466, 737790329]
    u = oracle_lib.rng_uniform(sv, 0, 150, 4)
    assert u.tolist() == [95.639743551146239, 89.846192230470479, 74.468740692827851, 74.010539765004069]


def test_shuffled_set(oracle_lib, golden):
    sv = oracle_lib.make_seedvec(golden["ga_seed"])
    assert oracle_lib.shuffle(sv, 40, 3).tolist() == golden["shuffle40x3"]
    # SURVEY.md Appendix C.2, first call
    assert oracle_lib.shuffle(sv, 40, 1)[0].tolist() == [25, 24, 20, 21, 1, 11, 36, 32, 34, 9, 12, 10, 0, 7, 35, 27, 33, 6, 15, 4,
                                                         22, 37, 30, 17, 38, 28, 39, 13, 29, 14, 16, 5, 18, 19, 3, 8, 2, 31, 23, 26]
    got = [int(np.bitwise_xor.reduce(r * np.arange(1, 151, dtype=np.uint32))) for r in oracle_lib.shuffle(sv, 150, 2)]
    assert got == golden["shuffle150x2_xor"]


def test_closed_form_data_probe():
    X, y = synth.closed_form()
    assert X[0, 0] == pytest.approx(0.85859348094331045, rel=1e-15)
    assert X[39, 29] == pytest.approx(-0.0062681448296003084, rel=1e-12)
    assert y[0] == pytest.approx(-0.12391978207655883, rel=1e-13)
    assert y[39] == pytest.approx(1.1429555738176309, rel=1e-13)


def _case_names(golden_path="tests/golden/reference_vectors.json"):
    import json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, golden_path)) as fh:
        return [c["name"] for c in json.load(fh)["cases"]]


@pytest.mark.parametrize("name", _case_names())
def test_oracle_matches_golden_case(oracle_lib, golden, name):
    case = next(c for c in golden["cases"] if c["name"] == name)
    X, y = make_data(case["data"])
    assert [float(X[0, 0]), float(X[-1, -1]), float(y[0]), float(y[-1])] == pytest.approx(case["data_probe"], rel=1e-12)
    sv = oracle_lib.make_seedvec(golden["ga_seed"])
    ev = make_checker_evaluator(oracle_lib, case["evaluator"], X, y, sv)
    seg = ev.segmentation()
    assert len(seg) == case["segmentation_count"]
    assert seg_digest(seg) == case["segmentation_digest"]            # bit-exact CV row partitions
    if "segmentation" in case:
        assert [s.tolist() for s in seg] == case["segmentation"]
    subsets = make_subsets(case["subsets"], X.shape[1])
    fit, status = ev.evaluate(subsets)
    assert status.tolist() == case["status"]
    ok = np.asarray(case["status"]) == 0
    assert rel_err(fit[ok], np.asarray(case["fitness"])[ok]).max() <= FIT_TOL


def test_survey_appendix_c3_to_c6(oracle_lib):
    X, y = synth.closed_form()
    sv = oracle_lib.make_seedvec(123)
    ev = oracle_lib.pls_evaluator(X, y, sv, numReplications=2, maxNComp=5, innerSegments=4, outerSegments=3)
    seg = ev.segmentation()
    assert len(seg) == 60
    assert seg[0].tolist() == [0, 4, 6, 7, 9, 10, 12, 15, 17, 22, 27, 28, 30, 32, 33, 34, 35, 37, 38]
    assert seg[1].tolist() == [1, 11, 20, 21, 24, 25, 36]
    assert seg[9].tolist() == [2, 3, 5, 8, 13, 14, 16, 18, 19, 23, 26, 29, 31, 39]
    assert seg[11].tolist() == [5, 13, 14, 16, 18, 29, 39]
    fit, _ = ev.evaluate([[2, 5, 11, 29], list(range(12)), [7], [1, 2, 5, 8, 11, 13, 17, 19, 23, 29]])
    want = [-0.082050888884043549, -0.40119862398530326, -1.6049334059118827, -0.26544539104666021]
    assert rel_err(fit, want).max() < 1e-10
    ev = oracle_lib.pls_evaluator(X, y, sv, numReplications=3, maxNComp=0, innerSegments=7, outerSegments=1, complexityPenalty=0.01)
    assert len(ev.segmentation()) == 42
    fit, _ = ev.evaluate([[2, 5, 11, 29], list(range(12))])
    assert rel_err(fit, [-0.083226438687247023, -0.22331462446936037]).max() < 1e-10
    ev = oracle_lib.pls_evaluator(X, y, sv, numReplications=2, maxNComp=0, innerSegments=7, outerSegments=1, testSetSize=0.4)
    seg = ev.segmentation()
    assert len(seg) == 32 and len(seg[15]) == 17
    assert rel_err(ev.evaluate([[2, 5, 11, 29]])[0], [-0.069135724855778374]).max() < 1e-10
    # C.4 / C.5
    want_fit = {"BIC": 191.69625628073339, "AIC": 201.82953300541701, "adjusted.r.squared": 0.99758366715517111, "r.squared": 0.99795541066976012}
    want_lm = {"BIC": 195.40120575418379, "AIC": 203.84560302475347, "adjusted.r.squared": 0.9976556777900516, "r.squared": 0.9979562319195322}
    for stat in want_fit:
        assert rel_err(oracle_lib.fit_evaluator(X, y, sv, statistic=stat).evaluate([[2, 5, 11, 29]])[0], [want_fit[stat]]).max() < 1e-10
        assert rel_err(oracle_lib.lm_evaluator(X, y, stat).evaluate([[2, 5, 11, 29]])[0], [want_lm[stat]]).max() < 1e-10
    assert rel_err(oracle_lib.lm_evaluator(X, y, "BIC").evaluate([[1, 2, 5, 8, 11, 13, 17, 19, 23, 29]])[0], [174.02302223845993]).max() < 1e-10
    # C.6 raw SIMPLS
    B, b0 = oracle_lib.simpls_fit(X, y, [2, 5, 11, 29], 3)
    assert b0 == pytest.approx([1.5165956523380317, 1.5015457624543409, 1.5020288073313739], rel=1e-10)
    assert B[:, 2] == pytest.approx([0.98911261755332736, -1.9973398340715327, 0.4913494131847313, 0.24901465612526905], rel=1e-9)


def test_oracle_vs_reference_build_live(oracle_lib, ref_lib):
    """Live cross-execution (container / wherever oracle/_ref travelled): fresh random cases."""
    for seed, (n, p, kw) in enumerate([
        (57, 80, dict(numReplications=3, maxNComp=5, innerSegments=5, outerSegments=3)),
        (64, 50, dict(numReplications=2, maxNComp=0, innerSegments=4, outerSegments=1)),
        (45, 60, dict(numReplications=2, maxNComp=4, innerSegments=2, outerSegments=2, testSetSize=0.3)),
    ]):
        X, y = synth.nir_spectra(n, p, seed=100 + seed)
        sv = ref_lib.make_seedvec(1000 + seed)
        assert (sv == oracle_lib.make_seedvec(1000 + seed)).all()
        er, eo = ref_lib.pls_evaluator(X, y, sv, **kw), oracle_lib.pls_evaluator(X, y, sv, **kw)
        sr, so = er.segmentation(), eo.segmentation()
        assert len(sr) == len(so) and all((a == b).all() for a, b in zip(sr, so))
        subsets = synth.random_subsets(p, 10, 1, 25, seed=seed)
        fr, st_r = er.evaluate(subsets)
        fo, st_o = eo.evaluate(subsets, threads=2)
        assert (st_r == st_o).all() and rel_err(fo, fr).max() <= FIT_TOL
        for stat in ("BIC", "r.squared"):
            fr, _ = ref_lib.fit_evaluator(X, y, sv, statistic=stat, maxNComp=6).evaluate(subsets)
            fo, _ = oracle_lib.fit_evaluator(X, y, sv, statistic=stat, maxNComp=6).evaluate(subsets)
            assert rel_err(fo, fr).max() <= 1e-9
            fr, _ = ref_lib.lm_evaluator(X, y, stat).evaluate(subsets)
            fo, _ = oracle_lib.lm_evaluator(X, y, stat).evaluate(subsets)
            assert rel_err(fo, fr).max() <= 1e-9


def test_empty_subset_status(oracle_lib):
    X, y = synth.closed_form()
    sv = oracle_lib.make_seedvec(5)
    fit, status = oracle_lib.pls_evaluator(X, y, sv, numReplications=1, innerSegments=3).evaluate([[], [1, 2]])
    assert status.tolist() == [1, 0] and fit[0] == 0.0
