"""K1k, the kernel-space tensor-core kernel (gaselect_b200/csrc/simpls_kspace.cu): SIMPLS carried by vectors over the
rows with one K = Z Z' per chromosome.  Checked against the CPU oracle for every subset size (1 column to several
chunks of columns), every split of a chromosome's fits over CTAs, nested and single CV, segments that do not tile,
underflow; and against the per-fit kernel.  Reference semantics: PLSEvaluator::evaluate (src/PLSEvaluator.cpp:71-91,
228-343) over PLSSimpls::fit (src/PLSSimpls.cpp:106-166)."""
import numpy as np
import pytest

from gaselect_b200 import synth
from helpers import rel_err
from test_gpu_scores import forced

REL_TOL = 1e-8  # north_star: "within relative 1e-8 in FP64"

pytestmark = pytest.mark.gpu

KW = dict(numReplications=2, maxNComp=6, innerSegments=3, outerSegments=4)  # 60 rows, 20 distinct fits: 3 passes of 8


@pytest.fixture(scope="module")
def case(oracle_lib):
    X, y = synth.nir_spectra(60, 260, seed=41)
    sv = oracle_lib.make_seedvec(77)
    subsets = synth.random_subsets(260, 40, 1, 230, seed=3) + [np.arange(200, dtype=np.uint32), np.array([7], dtype=np.uint32)]
    want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **KW).evaluate(subsets, threads=4)
    return X, y, sv, subsets, want, wstatus


@pytest.mark.parametrize("nf", [8, 16])
@pytest.mark.parametrize("split", [1, 2, 3])
def test_every_subset_size_and_split_against_the_oracle(engine, case, split, nf):
    """GSB_KSPACE_MINK=0 sends every subset (1..230 columns: up to four column chunks of K) through K1k, with 16 fits
    per pass (the fits' stored vectors in tensor memory: GSB_KSPACE_NF=16) and with the default 8."""
    from gaselect_b200 import evaluators as E
    X, y, sv, subsets, want, wstatus = case
    with forced(GSB_KSPACE_MINK="0", GSB_KSPACE_SPLIT=str(split), GSB_NO_KSPACE=None, GSB_KSPACE_NF=str(nf)):
        with E.PLSEvaluator(X, y, sv, **KW) as ev:
            fit, status = ev.evaluate_population(subsets)
            passes = -(-ev.info()["distinct_fits"] // nf)
            assert ev.timing()["fit_ctas"] == len(subsets) * min(split, passes)  # one launch, <= `split` CTAs per chromosome
            assert (status == wstatus).all()
            assert rel_err(fit[wstatus == 0], want[wstatus == 0]).max() <= REL_TOL
            again, _ = ev.evaluate_population(subsets)
            assert (again == fit).all()  # fixed summation order: bitwise reproducible


def test_default_policy_matches_the_per_fit_kernel(engine, case):
    from gaselect_b200 import evaluators as E
    X, y, sv, subsets, want, wstatus = case
    with forced(GSB_KSPACE_MINK=None, GSB_KSPACE_SPLIT=None, GSB_NO_KSPACE=None, GSB_SCORE_CFG=None, GSB_NO_SCORES=None):
        with E.PLSEvaluator(X, y, sv, **KW) as ev:
            f1, s1 = ev.evaluate_population(subsets)
            ctas1 = ev.timing()["fit_ctas"]
    with forced(GSB_NO_KSPACE="1", GSB_NO_SCORES="1"):
        with E.PLSEvaluator(X, y, sv, **KW) as ev:
            f0, s0 = ev.evaluate_population(subsets)
            ctas0 = ev.timing()["fit_ctas"]
            nfits = ev.info()["distinct_fits"]
    assert ctas0 == len(subsets) * nfits and ctas1 < ctas0      # different kernels ran
    assert (s0 == s1).all() and rel_err(f1, f0).max() <= 1e-9  # same arithmetic up to rounding


def test_any_segmentation(engine, oracle_lib):
    """K1k needs no structure in the segments: 62 rows in 4 ragged outer segments, 150 rows in 3 (50-row blocks),
    single CV (outerSegments=1: a test set drawn once per replication), 160 rows (the largest K).  190 rows do not
    fit: the engine stays on the other kernels."""
    from gaselect_b200 import evaluators as E
    for n, kw, kspace in [(62, dict(numReplications=2, maxNComp=4, innerSegments=3, outerSegments=4), True),
                          (150, dict(numReplications=2, maxNComp=5, innerSegments=2, outerSegments=3), True),
                          (100, dict(numReplications=3, maxNComp=10, innerSegments=5, outerSegments=1, testSetSize=0.25), True),
                          (160, dict(numReplications=1, maxNComp=10, innerSegments=4, outerSegments=5), True),
                          (157, dict(numReplications=1, maxNComp=7, innerSegments=3, outerSegments=2), True),
                          (190, dict(numReplications=1, maxNComp=4, innerSegments=4, outerSegments=5), False)]:
        X, y = synth.nir_spectra(n, 90, seed=n)
        sv = oracle_lib.make_seedvec(n)
        subsets = synth.random_subsets(90, 24, 1, 80, seed=n + 1)
        want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=4)
        with forced(GSB_KSPACE_MINK="0", GSB_KSPACE_SPLIT="1", GSB_NO_KSPACE=None):
            with E.PLSEvaluator(X, y, sv, **kw) as ev:
                fit, status = ev.evaluate_population(subsets)
                ran = ev.timing()["fit_ctas"] == len(subsets)
        assert (status == wstatus).all() and rel_err(fit[wstatus == 0], want[wstatus == 0]).max() <= REL_TOL, (n, kw)
        assert ran == kspace, (n, kw)


def test_more_components_than_the_kernel_holds_fall_back(engine, oracle_lib):
    """maxNComp > 10: the nine stored vectors per fit live in registers, so K1k declines."""
    from gaselect_b200 import evaluators as E
    X, y = synth.nir_spectra(60, 90, seed=5)
    sv = oracle_lib.make_seedvec(5)
    kw = dict(numReplications=1, maxNComp=14, innerSegments=3, outerSegments=4)
    subsets = synth.random_subsets(90, 12, 20, 80, seed=6)
    want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=4)
    with forced(GSB_KSPACE_MINK="0", GSB_KSPACE_SPLIT="1", GSB_NO_KSPACE=None):
        with E.PLSEvaluator(X, y, sv, **kw) as ev:
            fit, status = ev.evaluate_population(subsets)
            assert ev.timing()["fit_ctas"] > len(subsets)
    assert (status == wstatus).all() and rel_err(fit, want).max() <= REL_TOL


def test_underflow_and_tiny_subsets(engine, oracle_lib):
    """Constant response: X'y vanishes and the reference throws (src/PLSSimpls.cpp:141-143) -> status 2 for every
    chromosome; single-column subsets run one component."""
    from gaselect_b200 import evaluators as E
    X, y = synth.nir_spectra(60, 50, seed=8)
    sv = oracle_lib.make_seedvec(9)
    subsets = [[1, 2, 3], [5], list(range(0, 50, 2)), list(range(50))]
    with forced(GSB_KSPACE_MINK="0", GSB_KSPACE_SPLIT="1", GSB_NO_KSPACE=None):
        with E.PLSEvaluator(X, np.full(60, 1.5), sv, **KW) as ev:
            _, status = ev.evaluate_population(subsets)
            assert ev.timing()["fit_ctas"] == len(subsets)
        assert status.tolist() == [2, 2, 2, 2]
        want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **KW).evaluate(subsets)
        with E.PLSEvaluator(X, y, sv, **KW) as ev:
            fit, status = ev.evaluate_population(subsets)
        assert (status == wstatus).all() and rel_err(fit, want).max() <= REL_TOL


def test_config3_shape_with_wide_subsets(engine, oracle_lib):
    """BASELINE config 3 (n=150, p=1000, 10 x 5/4 CV, <= 10 components): 150 distinct fits = 19 passes; subsets up to
    400 columns (three chunks of 152 columns in the K build)."""
    from gaselect_b200 import evaluators as E
    X, y = synth.nir_spectra(150, 1000, seed=777)
    sv = oracle_lib.make_seedvec(123)
    kw = dict(numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5)
    subsets = synth.random_subsets(1000, 10, 57, 400, seed=12) + synth.random_subsets(1000, 4, 2, 12, seed=13)
    want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=8)
    for env in (dict(GSB_KSPACE_MINK="0", GSB_KSPACE_SPLIT=None), dict(GSB_KSPACE_MINK=None, GSB_KSPACE_SPLIT="5")):
        with forced(GSB_NO_KSPACE=None, **env):
            with E.PLSEvaluator(X, y, sv, **kw) as ev:
                fit, status = ev.evaluate_population(subsets)
        assert (status == wstatus).all() and rel_err(fit, want).max() <= REL_TOL


@pytest.mark.parametrize("stat", ["BIC", "adjusted.r.squared"])
def test_bic_evaluator_on_the_kernel_space_kernel(engine, oracle_lib, stat):
    """BICEvaluator (src/BICEvaluator.cpp:54-91, 144-233): K-fold one-SE rule + a refit on ALL rows whose residuals
    are taken on those same rows -- for K1k just one more fit whose prediction task covers every row."""
    from gaselect_b200 import evaluators as E
    X, y = synth.nir_spectra(120, 150, seed=15)
    sv = oracle_lib.make_seedvec(31)
    subsets = synth.random_subsets(150, 30, 1, 110, seed=4)
    fo, so = oracle_lib.fit_evaluator(X, y, sv, numSegments=6, statistic=stat, maxNComp=7).evaluate(subsets, threads=4)
    with forced(GSB_KSPACE_MINK="0", GSB_KSPACE_SPLIT="1", GSB_NO_KSPACE=None):
        with E.BICEvaluator(X, y, sv, numSegments=6, statistic=stat, maxNComp=7) as ev:
            fg, sg = ev.evaluate_population(subsets)
            assert ev.timing()["fit_ctas"] == len(subsets)
    assert (sg == so).all() and rel_err(fg[so == 0], fo[so == 0]).max() <= REL_TOL
    with forced(GSB_KSPACE_MINK=None, GSB_KSPACE_SPLIT=None, GSB_NO_KSPACE=None):  # default policy: mixed kernels
        with E.BICEvaluator(X, y, sv, numSegments=6, statistic=stat, maxNComp=7) as ev:
            fd, sd = ev.evaluate_population(subsets)
    assert (sd == so).all() and rel_err(fd[so == 0], fo[so == 0]).max() <= REL_TOL


@pytest.mark.parametrize("split", [1, 2, 3, 4])
def test_outer_fits_stop_at_optncomp(engine, oracle_lib, split):
    """The reference fits its outer models with optNComp components only (src/PLSEvaluator.cpp:316-317).  K1k runs the
    one-SE rule in the kernel once a CTA's MSEP fits are done and stops its outer-only fits there: bit-identical fitness
    (a fit's first components do not depend on later ones), fewer tensor-core products."""
    from gaselect_b200 import evaluators as E
    X, y = synth.nir_spectra(150, 400, seed=5)
    sv = oracle_lib.make_seedvec(9)
    kw = dict(numReplications=6, maxNComp=10, innerSegments=4, outerSegments=5)
    subsets = synth.random_subsets(400, 24, 41, 160, seed=2)
    res = {}
    for early in (True, False):
        with forced(GSB_NO_KSPACE=None, GSB_KSPACE_SPLIT=str(split), GSB_NO_EARLYSTOP=None if early else "1"):
            with E.PLSEvaluator(X, y, sv, **kw) as ev:
                fit, status = ev.evaluate_population(subsets)
                res[early] = (fit, status, ev.timing()["fit_tensor_gflop"], ev.timing()["fit_ctas"])
    assert (res[True][0] == res[False][0]).all() and (res[True][1] == res[False][1]).all()
    assert res[True][3] == res[False][3] == len(subsets) * split
    if 6 % split == 0:  # whole replications per CTA
        assert res[True][2] < res[False][2], (res[True][2], res[False][2])
    else:               # an uneven deal of replications: the engine keeps plan order and runs every component
        assert res[True][2] == res[False][2]
    want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=8)
    assert (res[True][1] == wstatus).all() and rel_err(res[True][0], want).max() <= REL_TOL


def test_near_tie_flag(engine, oracle_lib):
    """GSB_STATUS_FLAGGED (4): a comparison of the one-SE rule (src/PLSEvaluator.cpp:284-308) was closer than the
    relative band; the fitness is the same valid number.  The default band (1e-9) flags nothing on this data; a band of
    0.5 flags nearly everything; a negative band switches the flag off.  K1k and the per-fit kernel alike."""
    from gaselect_b200 import evaluators as E, _capi
    X, y = synth.nir_spectra(60, 120, seed=3)
    sv = oracle_lib.make_seedvec(123)
    kw = dict(numReplications=3, maxNComp=6, innerSegments=3, outerSegments=4)
    subsets = synth.random_subsets(120, 40, 2, 100, seed=9)
    out = {}
    for band in (0.0, 0.5, -1.0):
        with E.PLSEvaluator(X, y, sv, tieBand=band, **kw) as ev:
            out[band] = ev.evaluate_population(subsets)
            if band == 0.5:
                assert ev.evaluate(subsets[0]) == out[band][0][0]  # the wrapper returns the fitness, it does not raise
    assert (out[0.0][1] == 0).all() and (out[-1.0][1] == 0).all()
    assert (out[0.5][1] == _capi.STATUS_FLAGGED).sum() > len(subsets) // 2 and set(out[0.5][1].tolist()) <= {0, 4}
    assert (out[0.5][0] == out[0.0][0]).all() and (out[-1.0][0] == out[0.0][0]).all()
    with E.BICEvaluator(X, y, sv, numSegments=5, maxNComp=6, tieBand=0.5) as ev:
        fb, sb = ev.evaluate_population(subsets)
    with E.BICEvaluator(X, y, sv, numSegments=5, maxNComp=6) as ev:
        f0, s0 = ev.evaluate_population(subsets)
    assert (fb == f0).all() and (s0 == 0).all() and (sb == 4).any()


def test_ill_conditioned_subsets_are_rescored_on_the_gram_space_kernel(engine, oracle_lib):
    """K1k takes |v|^2 of an orthogonalised loading by subtraction on K = Z Z'; the reference takes the norm of the
    explicit vector (src/PLSSimpls.cpp:57-63).  For nearly collinear columns the subtraction leaves nothing: the kernel
    marks the chromosome (GSB_STATUS_RETRY), and the download has it scored again by the Gram-space kernel -- no
    chromosome is reported as an underflow that the reference scores; the result is the per-fit path's to the bit."""
    from gaselect_b200 import evaluators as E
    rng = np.random.default_rng(4)
    X, y = synth.nir_spectra(90, 300, seed=8)
    X = X.copy()
    for j in range(0, 96, 8):  # groups of eight columns that agree to 1e-7: a subset of 48 of them has rank 6 (+ 1e-7), ten components
        for d in range(1, 8):
            X[:, j + d] = X[:, j] * (1.0 + 1e-7 * rng.standard_normal(90))
    sv = oracle_lib.make_seedvec(21)
    kw = dict(numReplications=2, maxNComp=10, innerSegments=3, outerSegments=3)
    subsets = [np.arange(0, 48, dtype=np.uint32), np.arange(48, 96, dtype=np.uint32)] + synth.random_subsets(300, 10, 45, 120, seed=6)
    subsets = [s if i < 2 else s[s >= 96] for i, s in enumerate(subsets)]
    with forced(GSB_NO_KSPACE=None, GSB_KSPACE_MINK=None):
        with E.PLSEvaluator(X, y, sv, **kw) as ev:
            f1, s1 = ev.evaluate_population(subsets)
            retries = ev.timing()["retries"]
            bits, _ = ev.evaluate_population(subsets)
    with forced(GSB_NO_KSPACE="1", GSB_NO_SCORES="1"):
        with E.PLSEvaluator(X, y, sv, **kw) as ev:
            f0, s0 = ev.evaluate_population(subsets)
    want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=4)
    assert retries == 2, retries                        # the collinear subsets were declined, the ordinary ones were not
    assert (s1 == s0).all() and (s1[2:] == wstatus[2:]).all() and np.isfinite(f1).all() and (bits == f1).all()
    assert (f1[:2] == f0[:2]).all()                      # re-scored by the same kernel that GSB_NO_KSPACE selects
    assert rel_err(f1[2:], want[2:]).max() <= REL_TOL   # ordinary subsets: parity as everywhere


def test_tables_that_do_not_fit_leave_everything_to_the_kernel_space_kernel(engine, oracle_lib):
    """p = 60 000 columns, 5 x 3-fold CV: the unit moment tables of the Gram-space path (p (p + 1) / 2 entries x 16 units
    = 230 GB) do not fit in device memory.  The reference handles any p up to 65535 (src/Control.h:67-68);
    K1k needs no tables, so it takes the subsets of <= 40 columns as well instead of the call failing with GSB_ERR_MEMORY."""
    from gaselect_b200 import evaluators as E
    X, y = synth.gaussian(48, 60000, seed=5)
    sv = oracle_lib.make_seedvec(3)
    kw = dict(numReplications=5, maxNComp=6, innerSegments=3, outerSegments=3)
    subsets = synth.random_subsets(60000, 12, 2, 40, seed=8) + synth.random_subsets(60000, 6, 41, 150, seed=9)
    want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=8)
    with forced(GSB_NO_KSPACE=None, GSB_KSPACE_MINK=None, GSB_KSPACE_SPLIT=None):
        with E.PLSEvaluator(X, y, sv, **kw) as ev:
            fit, status = ev.evaluate_population(subsets)
            assert ev.info()["gram_bytes"] == 0               # no tables were built
            assert ev.timing()["kernel_launches"] == 2        # one K1k launch + K2
    assert (status == wstatus).all() and rel_err(fit, want).max() <= REL_TOL


def test_segmentations_beyond_the_pack_stage(engine, oracle_lib):
    """120 replications of 4 x 3-fold CV: 1 + 120 x (4 + 12) unit blocks, more than the pack stage of the per-fit kernel
    holds in shared memory.  With n <= 160 the kernel-space kernel takes every subset size (it needs no tables); where no
    other kernel can serve the plan (n = 200) the engine says so instead of failing inside a launch."""
    from gaselect_b200 import evaluators as E, _capi
    sv = oracle_lib.make_seedvec(11)
    kw = dict(numReplications=120, maxNComp=5, innerSegments=3, outerSegments=4)
    X, y = synth.nir_spectra(60, 300, seed=21)
    subsets = synth.random_subsets(300, 10, 2, 40, seed=3) + synth.random_subsets(300, 6, 41, 120, seed=4)
    want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=8)
    with forced(GSB_NO_KSPACE=None, GSB_KSPACE_MINK=None, GSB_KSPACE_SPLIT=None):
        with E.PLSEvaluator(X, y, sv, **kw) as ev:
            fit, status = ev.evaluate_population(subsets)
            assert ev.info()["gram_bytes"] == 0
        assert (status == wstatus).all() and rel_err(fit, want).max() <= REL_TOL
        X2, y2 = synth.nir_spectra(200, 300, seed=22)
        with E.PLSEvaluator(X2, y2, sv, **kw) as ev:
            with pytest.raises(_capi.EngineError, match="pack stage"):
                ev.evaluate_population(subsets[:4])
