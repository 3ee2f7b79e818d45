"""The row-space ("kernel space") form of SIMPLS that K1k computes agrees with the oracle's column-space restatement:
this pins the FORMULAS of gaselect_b200/csrc/simpls_kspace.cu on the CPU (the GPU parity tests pin the kernel)."""
import numpy as np

from gaselect_b200 import synth
from helpers import rel_err
import kspace_model


def test_row_space_simpls_matches_the_oracle(oracle_lib):
    X, y = synth.nir_spectra(60, 260, seed=41)
    sv = oracle_lib.make_seedvec(77)
    kw = dict(numReplications=2, maxNComp=6, innerSegments=3, outerSegments=4)
    ev = oracle_lib.pls_evaluator(X, y, sv, **kw)
    seg = [np.asarray(s, dtype=np.int64) for s in ev.segmentation()]
    subsets = synth.random_subsets(260, 6, 1, 230, seed=3) + [np.array([7], dtype=np.uint32)]
    want, status = ev.evaluate(subsets)
    assert (status == 0).all()
    got = [kspace_model.evaluate(X, y, s, seg, 2, 4, 3, 6, 1.0 / np.sqrt(3)) for s in subsets]
    assert rel_err(got, want).max() <= 1e-9
