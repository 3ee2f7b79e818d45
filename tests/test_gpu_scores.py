"""K1x, the score-space tensor-core kernel (gaselect_b200/csrc/simpls_scores.cu): every configuration the engine can
pick (fits per CTA, cluster size, loadings in shared memory or in the L2 scratch) against the CPU oracle, the plans
it must refuse (segments that do not tile into disjoint row blocks), and its agreement with the per-fit kernel.
The reference semantics checked are those of PLSEvaluator::evaluate (src/PLSEvaluator.cpp:71-91, 228-343)."""
import os

import numpy as np
import pytest

from gaselect_b200 import synth
from helpers import rel_err

REL_TOL = 1e-8  # north_star: "within relative 1e-8 in FP64"

pytestmark = pytest.mark.gpu


class forced:
    """GSB_SCORE_CFG=nf,cs,vg / GSB_NO_SCORES=1 for the duration of a block (read by the engine at every call)."""

    def __init__(self, **env):
        self.env = env

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.env}
        for k, v in self.env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.fixture(autouse=True)
def _no_kspace():
    """These tests are about K1x: keep K1k (tests/test_gpu_kspace.py), which would otherwise take the larger subsets."""
    with forced(GSB_NO_KSPACE="1"):
        yield


KW = dict(numReplications=2, maxNComp=6, innerSegments=3, outerSegments=4)  # 60 rows: 4 atoms of 15, 10 fits / replication


@pytest.fixture(scope="module")
def case(oracle_lib):
    X, y = synth.nir_spectra(60, 260, seed=41)
    sv = oracle_lib.make_seedvec(77)
    subsets = synth.random_subsets(260, 40, 1, 230, seed=3) + [np.arange(200, dtype=np.uint32), np.array([7], dtype=np.uint32)]
    want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **KW).evaluate(subsets, threads=4)
    return X, y, sv, subsets, want, wstatus


@pytest.mark.parametrize("cfg,groups", [
    ("16,1,0", 2), ("16,1,1", 2), ("8,1,0", 4), ("8,1,1", 4), ("16,2,0", 2), ("16,2,1", 2), ("8,2,0", 4), ("8,2,1", 4),
    ("16,3,0", 2), ("16,3,1", 2), ("8,3,0", 4), ("8,3,1", 4),
])
def test_every_configuration_against_the_oracle(engine, case, cfg, groups):
    """Subsets beyond a configuration's capacity take the next kernel in the engine's list; the rest must run where
    they were forced to (CTA count) and agree with the oracle at 1e-8."""
    from gaselect_b200 import evaluators as E
    X, y, sv, subsets, want, wstatus = case
    with forced(GSB_SCORE_CFG=cfg, GSB_NO_SCORES=None):
        with E.PLSEvaluator(X, y, sv, **KW) as ev:
            fit, status = ev.evaluate_population(subsets)
            assert (status == wstatus).all()
            assert rel_err(fit[wstatus == 0], want[wstatus == 0]).max() <= REL_TOL
            # a small population that certainly fits the configuration: CTAs = chromosomes x groups x cluster size
            small = [s for s in subsets if len(s) <= 40][:8]
            ev.evaluate_population(small)
            cs = int(cfg.split(",")[1])
            assert ev.timing()["fit_ctas"] == len(small) * groups * cs


def test_default_policy_matches_the_per_fit_kernel(engine, case):
    from gaselect_b200 import evaluators as E
    X, y, sv, subsets, want, wstatus = case
    with forced(GSB_SCORE_CFG=None, GSB_NO_SCORES=None):
        with E.PLSEvaluator(X, y, sv, **KW) as ev:
            f1, s1 = ev.evaluate_population(subsets)
            ctas1 = ev.timing()["fit_ctas"]
    with forced(GSB_SCORE_CFG=None, GSB_NO_SCORES="1"):
        with E.PLSEvaluator(X, y, sv, **KW) as ev:
            f0, s0 = ev.evaluate_population(subsets)
            ctas0 = ev.timing()["fit_ctas"]
            nfits = ev.info()["distinct_fits"]
    assert ctas0 == len(subsets) * nfits and ctas1 < ctas0      # different kernels ran
    assert (s0 == s1).all() and rel_err(f1, f0).max() <= 1e-9  # same arithmetic up to the summation order


def test_atoms_wider_than_a_slot_and_plans_without_atoms(engine, oracle_lib):
    """150 rows in 3 outer segments: atoms of 50 rows take two 32-row slots each.  190 rows in 5 segments need ten
    slots and 62 rows in 4 segments do not tile at all: both stay on the per-fit kernel."""
    from gaselect_b200 import evaluators as E
    for n, kw, scores in [(150, dict(numReplications=2, maxNComp=5, innerSegments=2, outerSegments=3), True),
                          (190, dict(numReplications=1, maxNComp=4, innerSegments=4, outerSegments=5), False),
                          (62, dict(numReplications=2, maxNComp=4, innerSegments=3, outerSegments=4), False)]:
        X, y = synth.nir_spectra(n, 90, seed=n)
        sv = oracle_lib.make_seedvec(n)
        subsets = synth.random_subsets(90, 24, 1, 80, seed=n + 1)
        want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=4)
        with forced(GSB_SCORE_CFG=None, GSB_NO_SCORES=None):
            with E.PLSEvaluator(X, y, sv, **kw) as ev:
                fit, status = ev.evaluate_population(subsets)
                per_fit = ev.timing()["fit_ctas"] == len(subsets) * ev.info()["distinct_fits"]
        assert (status == wstatus).all() and rel_err(fit[wstatus == 0], want[wstatus == 0]).max() <= REL_TOL
        assert per_fit != scores


def test_underflow_and_tiny_subsets(engine, oracle_lib):
    """Constant response: X'y vanishes and the reference throws (src/PLSSimpls.cpp:141-143) -> status 2 for every
    chromosome; single-column subsets run one component."""
    from gaselect_b200 import evaluators as E
    X, y = synth.nir_spectra(60, 50, seed=8)
    sv = oracle_lib.make_seedvec(9)
    subsets = [[1, 2, 3], [5], list(range(0, 50, 2)), list(range(50))]
    for cfg in ("16,1,0", "8,2,1"):
        with forced(GSB_SCORE_CFG=cfg, GSB_NO_SCORES=None):
            with E.PLSEvaluator(X, np.full(60, 1.5), sv, **KW) as ev:
                _, status = ev.evaluate_population(subsets)
                assert ev.timing()["fit_ctas"] < len(subsets) * ev.info()["distinct_fits"]
            assert status.tolist() == [2, 2, 2, 2]
            want, wstatus = oracle_lib.pls_evaluator(X, y, sv, **KW).evaluate(subsets)
            with E.PLSEvaluator(X, y, sv, **KW) as ev:
                fit, status = ev.evaluate_population(subsets)
            assert (status == wstatus).all() and rel_err(fit, want).max() <= REL_TOL
