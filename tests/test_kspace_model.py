"""The row-space ("kernel space") form of SIMPLS that K1k computes agrees with the oracle's column-space restatement:
this pins the FORMULAS of gaselect_b200/csrc/simpls_kspace.cu on the CPU (the GPU parity tests pin the kernel)."""
import numpy as np

from gaselect_b200 import synth
from helpers import rel_err
import kspace_model


import pytest


@pytest.mark.parametrize("n,kw,inner", [
    (60, dict(numReplications=2, maxNComp=6, innerSegments=3, outerSegments=4), 3),            # segments that tile
    (62, dict(numReplications=2, maxNComp=4, innerSegments=3, outerSegments=4), 3),            # ragged segments
    (50, dict(numReplications=3, maxNComp=5, innerSegments=4, outerSegments=1, testSetSize=0.3), 4),  # single CV
])
def test_row_space_simpls_matches_the_oracle(oracle_lib, n, kw, inner):
    X, y = synth.nir_spectra(n, 260, seed=41)
    sv = oracle_lib.make_seedvec(77)
    ev = oracle_lib.pls_evaluator(X, y, sv, **kw)
    seg = [np.asarray(s, dtype=np.int64) for s in ev.segmentation()]
    R, O = kw["numReplications"], kw["outerSegments"]
    I = len(seg) // (2 * R * O) - 1                                  # effective inner segments (src/PLSEvaluator.cpp:34)
    subsets = synth.random_subsets(260, 5, 1, 230, seed=3) + [np.array([7], dtype=np.uint32)]
    want, status = ev.evaluate(subsets)
    assert (status == 0).all()
    got = [kspace_model.evaluate(X, y, s, seg, R, O, I, kw["maxNComp"], 1.0 / np.sqrt(I)) for s in subsets]
    assert rel_err(got, want).max() <= 1e-9
