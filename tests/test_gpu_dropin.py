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
        # LMEvaluator does not refuse an empty subset: the intercept-only model (src/LMEvaluator.cpp:27-67)
        assert dropin.dropin_evaluate_one(h, np.zeros(1, dtype=np.uint32), 0, C.byref(f)) == 0
        assert f.value == pytest.approx(-37.56167949519022, rel=1e-8)  # the reference build's value for this data
    finally:
        dropin.dropin_destroy(h)
    # PLSEvaluator throws on an empty subset (src/PLSEvaluator.cpp:72-75) -> EvaluatorException through the virtual interface
    sv = np.ascontiguousarray(np.arange(1, 625, dtype=np.uint32))
    h = dropin.dropin_pls_create(_colmajor(X), np.ascontiguousarray(y), 40, 30, 2, 5, sv, 4, 3, 0.0, 1.0, 0.0)
    assert h
    try:
        f = C.c_double(0.0)
        assert dropin.dropin_evaluate_one(h, np.zeros(1, dtype=np.uint32), 0, C.byref(f)) == 1  # EvaluatorException
        assert dropin.dropin_evaluate_one(h, np.asarray([1, 2, 5], dtype=np.uint32), 3, C.byref(f)) == 0 and f.value < 0.0
    finally:
        dropin.dropin_destroy(h)


def _batched_run(dropin, handle, seedvec, chromosomeSize, popSize, numGenerations, elitism=10, minVariables=1, maxVariables=None,
                 mutationProb=0.01, maxDup=0, badSolutionThreshold=2.0, crossover=0, fitnessScaling=0, numThreads=1):
    dropin.dropin_ga_run_batched.restype = C.c_int
    dropin.dropin_ga_run_batched.argtypes = (
        [C.c_void_p] + [C.c_int] * 6 + [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _u32p,
                                        C.c_int, C.c_longlong, C.POINTER(C.c_int), _u32p, _u32p, _f64p, C.c_int,
                                        C.POINTER(C.c_int), _f64p, np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS"),
                                        C.POINTER(C.c_longlong)])
    maxVariables = int(maxVariables or chromosomeSize)
    cap = popSize + elitism + 8
    idx_cap = cap * maxVariables
    n_res, n_evo, ns = C.c_int(0), C.c_int(0), C.c_longlong(0)
    res_idx = np.zeros(idx_cap, dtype=np.uint32)
    res_off = np.zeros(cap + 1, dtype=np.uint32)
    res_fit = np.zeros(cap, dtype=np.float64)
    evo = np.zeros(3 * (numGenerations + 2), dtype=np.float64)
    stats = np.zeros(4, dtype=np.uint64)
    rc = dropin.dropin_ga_run_batched(handle, chromosomeSize, popSize, numGenerations, elitism, minVariables, maxVariables,
                                      mutationProb, numThreads, maxDup, badSolutionThreshold, 1e-6, 0, crossover, fitnessScaling, seedvec,
                                      cap, idx_cap, C.byref(n_res), res_idx, res_off, res_fit, len(evo), C.byref(n_evo), evo,
                                      stats, C.byref(ns))
    assert rc == 0
    return {"subsets": [res_idx[res_off[i]:res_off[i + 1]].tolist() for i in range(n_res.value)],
            "fitness": res_fit[:n_res.value].copy(), "evolution": evo[:n_evo.value].reshape(-1, 3).copy(),
            "engine_evaluations": int(stats[0]), "engine_calls": int(stats[1]), "passes": int(stats[2]),
            "seconds": ns.value * 1e-9}


def test_batched_population_reproduces_golden_single_thread_run(dropin, ref_lib, golden):
    """The speculate-and-replay driver + one engine call per generation == SingleThreadPopulation."""
    run = next(r for r in golden["ga_runs"] if r["numThreads"] == 1)
    X, y = synth.closed_form()
    sv = ref_lib.make_seedvec(golden["ga_seed"])
    h = dropin.dropin_pls_create(_colmajor(X), np.ascontiguousarray(y), 40, 30, 2, 5, sv, 4, 3, 0.0, 1.0, 0.0)
    try:
        r = _batched_run(dropin, h, sv, 30, 40, 8, elitism=5, minVariables=3, maxVariables=10)
    finally:
        dropin.dropin_destroy(h)
    assert r["subsets"] == run["subsets"]
    assert rel_err(r["fitness"], run["fitness"]).max() <= 1e-8
    assert rel_err(r["evolution"][:, 0], run["best"]).max() <= 1e-8
    assert r["engine_calls"] <= 2 * (8 + 1) and r["engine_evaluations"] <= run["evaluations"]


@pytest.mark.parametrize("threshold,dup", [(2.0, 0), (0.1, 0), (2.0, 3)])
def test_batched_population_against_the_reference_driver_live(dropin, ref_lib, threshold, dup):
    """BASELINE config 1 shape (n=100, p=200, population 100, evaluatorPLS defaults), fewer generations:
    reference SingleThreadPopulation + reference PLSEvaluator on the CPU vs BatchedPopulation + engine.
    threshold 0.1 forces real rejections (replayed passes), dup 3 exercises duplicate elimination."""
    X, y = synth.gaussian(100, 200, seed=777)
    sv = ref_lib.make_seedvec(321)
    kw = dict(elitism=10, minVariables=5, maxVariables=30, badSolutionThreshold=threshold)
    ev = ref_lib.pls_evaluator(X, y, sv)  # evaluatorPLS() defaults: 30 replications, srCV, 7 segments
    want = ref_lib.ga_run(ev, sv, 200, 100, 6, numThreads=1, maxDuplicateEliminationTries=dup, **kw)
    h = dropin.dropin_pls_create(_colmajor(X), np.ascontiguousarray(y), 100, 200, 30, 0, sv, 7, 1, 0.0, 1.0, 0.0)
    try:
        got = _batched_run(dropin, h, sv, 200, 100, 6, maxDup=dup, **kw)
    finally:
        dropin.dropin_destroy(h)
    assert got["subsets"] == [s.tolist() for s in want["subsets"]]      # bit-exact chromosomes
    assert rel_err(got["fitness"], want["fitness"]).max() <= 1e-8
    assert rel_err(got["evolution"], want["fitnessEvolution"]).max() <= 1e-6  # mean / sd columns amplify 1e-8 differences
    if threshold >= 2.0:
        assert got["engine_calls"] == 6 + 1                               # one engine call per generation
    assert got["engine_evaluations"] <= want["evaluations"]


def test_batched_population_virtual_threads_reproduce_golden_four_thread_run(dropin, ref_lib, golden):
    """numThreads = 4: MultiThreadedPopulation's result depends on the number of workers (one random stream and
    one slice each); BatchedPopulation reproduces them as virtual threads, one engine call per generation."""
    run = next(r for r in golden["ga_runs"] if r["numThreads"] == 4)
    X, y = synth.closed_form()
    sv = ref_lib.make_seedvec(golden["ga_seed"])
    h = dropin.dropin_pls_create(_colmajor(X), np.ascontiguousarray(y), 40, 30, 2, 5, sv, 4, 3, 0.0, 1.0, 0.0)
    try:
        r = _batched_run(dropin, h, sv, 30, 40, 8, elitism=5, minVariables=3, maxVariables=10, numThreads=4)
    finally:
        dropin.dropin_destroy(h)
    assert r["subsets"] == run["subsets"]
    assert rel_err(r["fitness"], run["fitness"]).max() <= 1e-8
    assert rel_err(r["evolution"][:, 0], run["best"]).max() <= 1e-8


@pytest.mark.parametrize("threads,threshold,dup", [(3, 2.0, 0), (7, 0.1, 0), (4, 2.0, 2)])
def test_batched_population_virtual_threads_against_the_reference_driver_live(dropin, ref_lib, threads, threshold, dup):
    X, y = synth.gaussian(100, 200, seed=777)
    sv = ref_lib.make_seedvec(99)
    kw = dict(elitism=10, minVariables=5, maxVariables=30, badSolutionThreshold=threshold)
    ev = ref_lib.pls_evaluator(X, y, sv, numReplications=4, innerSegments=5)
    want = ref_lib.ga_run(ev, sv, 200, 100, 6, numThreads=threads, maxDuplicateEliminationTries=dup, **kw)
    h = dropin.dropin_pls_create(_colmajor(X), np.ascontiguousarray(y), 100, 200, 4, 0, sv, 5, 1, 0.0, 1.0, 0.0)
    try:
        got = _batched_run(dropin, h, sv, 200, 100, 6, maxDup=dup, numThreads=threads, **kw)
    finally:
        dropin.dropin_destroy(h)
    assert got["subsets"] == [s.tolist() for s in want["subsets"]]
    assert rel_err(got["fitness"], want["fitness"]).max() <= 1e-8
    assert rel_err(got["evolution"], want["fitnessEvolution"]).max() <= 1e-6
