"""The engine's own evidence against out-of-bounds writes (compute-sanitizer is not available on the pool the kernels were
developed on): with GSB_GUARD=1 every device array carries a 256-byte guard zone behind its last element and the K1
kernels a 64-byte canary behind their shared-memory carve; every evaluate call checks all of them and fails when one was
overwritten.  Run over the shapes that stress the carves and scratch sizes: every K1 kernel (kernel-space with 1-4 CTAs
per chromosome and the largest K, per-fit with prediction in the kernel and in its own kernel, spill path, generic path
with its loadings in shared memory and in the global scratch, score-space fallback), K2, K3 (shared-memory and global
triangle), populations of one chromosome and of several hundred, repeated calls with growing and shrinking batches."""
import numpy as np
import pytest

from gaselect_b200 import synth
from test_gpu_scores import forced

pytestmark = pytest.mark.gpu


def _run(make, batches):
    with make() as ev:
        for subsets in batches:
            fit, status = ev.evaluate_population(subsets)   # raises EngineError("GSB_GUARD: ...") on an overwritten guard
            assert np.isin(status, (0, 1, 2, 4)).all()


def test_no_kernel_writes_out_of_bounds(engine):
    from gaselect_b200 import evaluators as E
    with forced(GSB_GUARD="1", GSB_NO_KSPACE=None, GSB_KSPACE_SPLIT=None, GSB_KSPACE_MINK=None):
        X, y = synth.nir_spectra(150, 600, seed=2)
        wide = synth.random_subsets(600, 300, 1, 230, seed=1)
        few = synth.random_subsets(600, 3, 41, 400, seed=2)
        kw = dict(numReplications=3, maxNComp=10, innerSegments=4, outerSegments=5)
        _run(lambda: E.PLSEvaluator(X, y, 11, **kw), [few, wide, wide[:37], [np.zeros(0, dtype=np.uint32)] + wide[:150], few])
        for split in ("1", "3", "4"):
            with forced(GSB_KSPACE_SPLIT=split, GSB_KSPACE_MINK="0"):
                _run(lambda: E.PLSEvaluator(X, y, 11, **kw), [wide[:60]])
        X160, y160 = synth.nir_spectra(160, 300, seed=3)   # the largest K the kernel-space kernel takes
        _run(lambda: E.PLSEvaluator(X160, y160, 5, **kw), [synth.random_subsets(300, 40, 30, 300, seed=4)])
        with forced(GSB_NO_KSPACE="1"):                     # score-space fallback and the per-fit path at config-3 shape
            _run(lambda: E.PLSEvaluator(X, y, 11, **kw), [wide[:80]])
        with forced(GSB_NO_KSPACE="1", GSB_NO_SCORES="1"):
            _run(lambda: E.PLSEvaluator(X, y, 11, **kw), [wide[:80], synth.random_subsets(600, 12, 213, 416, seed=6)])
            _run(lambda: E.PLSEvaluator(X, y, 11, maxNComp=0, numReplications=2), [synth.random_subsets(600, 20, 5, 200, seed=7)])
            _run(lambda: E.PLSEvaluator(X, y, 11, maxNComp=14, numReplications=2), [synth.random_subsets(600, 20, 5, 120, seed=8)])
        Xb, yb = synth.nir_spectra(400, 900, seed=4)        # 80-row blocks: fit-only kernel + predictBlocksKernel, spill path
        _run(lambda: E.PLSEvaluator(Xb, yb, 7, **kw), [synth.random_subsets(900, 60, 1, 416, seed=5), synth.random_subsets(900, 7, 300, 500, seed=9)])
        _run(lambda: E.BICEvaluator(X, y, 3, numSegments=5, maxNComp=8), [wide[:50]])
        Xl, yl = synth.gaussian(300, 400, seed=12)
        _run(lambda: E.LMEvaluator(Xl, yl), [synth.random_subsets(400, 80, 1, 298, seed=10), [np.zeros(0, dtype=np.uint32)]])


def test_the_guard_notices_an_overwritten_zone(engine):
    """The check itself: the test writes one element beyond the fitness vector through the device pointer the API hands
    out; the next evaluate call reports it."""
    from gaselect_b200 import evaluators as E, _capi
    with forced(GSB_GUARD="1"):
        X, y = synth.nir_spectra(60, 120, seed=3)
        with E.PLSEvaluator(X, y, 5, numReplications=2, maxNComp=4, innerSegments=3, outerSegments=3) as ev:
            subsets = synth.random_subsets(120, 32, 2, 60, seed=1)
            idx, off = _capi.to_csr(subsets)
            ev.upload_population(idx, off)
            ev.evaluate_resident()
            f_ptr, s_ptr, npop = ev.device_results()
            import torch
            # 32 doubles were reserved for the fitness vector: write the 33rd
            class _Dev:
                def __init__(self, ptr, n):
                    self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
            t = torch.as_tensor(_Dev(f_ptr, npop + 1), device="cuda:0")
            t[npop] = 1.0
            torch.cuda.synchronize()
            with pytest.raises(_capi.EngineError, match="GSB_GUARD"):
                ev.evaluate_resident()
