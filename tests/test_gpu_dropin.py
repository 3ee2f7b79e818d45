"""GPU drop-in test: the reference's OWN GA drivers (SingleThreadPopulation /
MultiThreadedPopulation, compiled from the reference sources into oracle/_ref/libgaselect_ref.so)
run on top of the product's reference-side binding (gaselect_b200/host/GpuEvaluator, an
`Evaluator` subclass over the C ABI) and must reproduce the golden GA runs that the reference
produced with its own PLSEvaluator: same selected subsets (bit-exact), fitness within 1e-8."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from gaselect_b200 import synth
from helpers import rel_err

pytestmark = pytest.mark.gpu

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def dropin(engine, ref_lib):
    from oracle import bindings
    if os.path.isdir("/root/reference/src"):
        bindings.build("dropin")
    path = os.path.join(ROOT, "oracle", "_ref", "libgaselect_dropin.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libgaselect_dropin.so not available")
    dll = C.CDLL(path)
    dll.dropin_pls_create.restype = C.c_void_p
    dll.dropin_pls_create.argtypes = [_f64p, _f64p, C.c_int, C.c_int, C.c_int, C.c_int, _u32p, C.c_int, C.c_int,
                                      C.c_double, C.c_double, C.c_double]
    dll.dropin_lm_create.restype = C.c_void_p
    dll.dropin_lm_create.argtypes = [_f64p, _f64p, C.c_int, C.c_int, C.c_int]
    dll.dropin_evaluate_one.restype = C.c_int
    dll.dropin_evaluate_one.argtypes = [C.c_void_p, _u32p, C.c_int, C.POINTER(C.c_double)]
    dll.dropin_destroy.argtypes = [C.c_void_p]
    return dll


def _colmajor(X):
    return np.ascontiguousarray(np.asarray(X, dtype=np.float64).T).ravel()


@pytest.mark.parametrize("threads", [1, 4])
def test_reference_ga_driver_on_gpu_evaluator_reproduces_golden_run(dropin, ref_lib, golden, threads):
    run = next(r for r in golden["ga_runs"] if r["numThreads"] == threads)
    X, y = synth.closed_form()
    sv = ref_lib.make_seedvec(golden["ga_seed"])
    h = dropin.dropin_pls_create(_colmajor(X), np.ascontiguousarray(y), 40, 30, 2, 5, sv, 4, 3, 0.0, 1.0, 0.0)
    assert h
    try:
        r = ref_lib.ga_run(h, sv, 30, 40, 8, elitism=5, minVariables=3, maxVariables=10, numThreads=threads)
    finally:
        dropin.dropin_destroy(h)
    assert [s.tolist() for s in r["subsets"]] == run["subsets"]          # chromosomes bit-exact
    assert rel_err(r["fitness"], run["fitness"]).max() <= 1e-8
    assert rel_err(r["fitnessEvolution"][:, 0], run["best"]).max() <= 1e-8
    assert r["evaluations"] == run["evaluations"]


def test_virtual_interface_error_convention(dropin):
    X, y = synth.closed_form()
    h = dropin.dropin_lm_create(_colmajor(X), np.ascontiguousarray(y), 40, 30, 0)
    assert h
    try:
        f = C.c_double(0.0)
        assert dropin.dropin_evaluate_one(h, np.asarray([1, 2, 5, 8, 11, 13, 17, 19, 23, 29], dtype=np.uint32), 10, C.byref(f)) == 0
        assert f.value == pytest.approx(174.02302223845993, rel=1e-8)
        assert dropin.dropin_evaluate_one(h, np.zeros(1, dtype=np.uint32), 0, C.byref(f)) == 1  # EvaluatorException
    finally:
        dropin.dropin_destroy(h)
