"""ctypes bindings for the two CPU checkers -- TEST INFRASTRUCTURE, not product code.

* ``RefLib``    -> oracle/_ref/libgaselect_ref.so: the reference's own C++ compiled out of
                   tree from /root/reference/src (oracle/Makefile, oracle/ref_shim.cpp).
* ``OracleLib`` -> oracle/libgaselect_oracle.so: the plain-C restatement
                   (oracle/gaselect_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the
product package (gaselect_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libgaselect_ref.so")
ORACLE_SO = os.path.join(HERE, "libgaselect_oracle.so")

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

STAT = {"BIC": 0, "AIC": 1, "adjusted.r.squared": 2, "r.squared": 3}


def build(target="all"):
    """Run oracle/Makefile (``ref`` is a no-op when /root/reference is absent)."""
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def have_ref():
    return os.path.exists(REF_SO)


def seedvec_of(lib, seed):
    out = np.zeros(624, dtype=np.uint32)
    lib.seedvec(int(seed), out)
    return out


def to_csr(subsets):
    """list of ascending column-index lists -> (idx uint32, offsets uint32)."""
    offsets = np.zeros(len(subsets) + 1, dtype=np.uint32)
    for i, s in enumerate(subsets):
        offsets[i + 1] = offsets[i] + len(s)
    idx = np.zeros(max(int(offsets[-1]), 1), dtype=np.uint32)
    pos = 0
    for s in subsets:
        idx[pos:pos + len(s)] = np.asarray(s, dtype=np.uint32)
        pos += len(s)
    return idx, offsets


def _colmajor(X):
    X = np.asarray(X, dtype=np.float64)
    return np.ascontiguousarray(X.T).ravel()  # column-major bytes of X


class _Evaluator:
    """Common handle wrapper: segmentation() and evaluate(subsets, threads)."""

    def __init__(self, lib, prefix, handle):
        if not handle:
            raise ValueError("evaluator construction failed (invalid arguments)")
        self._lib, self._p, self._h = lib, prefix, handle

    def close(self):
        if self._h:
            getattr(self._lib, self._p + "evaluator_destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def segmentation(self):
        cnt = getattr(self._lib, self._p + "segmentation_count")(self._h)
        tot = getattr(self._lib, self._p + "segmentation_total")(self._h)
        rows = np.zeros(max(int(tot), 1), dtype=np.uint32)
        offsets = np.zeros(cnt + 1, dtype=np.uint32)
        getattr(self._lib, self._p + "segmentation_get")(self._h, rows, offsets)
        return [rows[offsets[i]:offsets[i + 1]].copy() for i in range(cnt)]

    def evaluate(self, subsets, threads=1, return_time=False):
        idx, offsets = to_csr(subsets)
        fit = np.zeros(len(subsets), dtype=np.float64)
        status = np.zeros(len(subsets), dtype=np.int32)
        ns = C.c_longlong(0)
        getattr(self._lib, self._p + "evaluate")(self._h, idx, offsets, len(subsets), int(threads), fit, status, C.byref(ns))
        if return_time:
            return fit, status, ns.value * 1e-9
        return fit, status


class _Lib:
    prefix = ""
    path = ""

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path + " missing: run `make -C oracle`")
        self.dll = C.CDLL(self.path)
        p = self.prefix
        d = self.dll
        self._sig(p + "rng_u32_stream", None, [C.c_uint32, C.c_int, _u32p])
        self._sig(p + "seedvec", None, [C.c_uint32, _u32p])
        self._sig(p + "rng_vec_stream", None, [_u32p, C.c_int, _u32p])
        self._sig(p + "rng_vec_uniform", None, [_u32p, C.c_double, C.c_double, C.c_int, _f64p])
        self._sig(p + "shuffle_all", None, [_u32p, C.c_int, C.c_int, _u32p])
        self._sig(p + "pls_evaluator_create", C.c_void_p,
                  [_f64p, _f64p, C.c_int, C.c_int, C.c_int, C.c_int, _u32p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double])
        self._sig(p + "fit_evaluator_create", C.c_void_p,
                  [_f64p, _f64p, C.c_int, C.c_int, C.c_int, _u32p, C.c_int, C.c_int, C.c_double])
        self._sig(p + "lm_evaluator_create", C.c_void_p, [_f64p, _f64p, C.c_int, C.c_int, C.c_int])
        self._sig(p + "evaluator_destroy", None, [C.c_void_p])
        self._sig(p + "segmentation_count", C.c_int, [C.c_void_p])
        self._sig(p + "segmentation_total", C.c_longlong, [C.c_void_p])
        self._sig(p + "segmentation_get", None, [C.c_void_p, _u32p, _u32p])
        self._sig(p + "evaluate", C.c_int,
                  [C.c_void_p, _u32p, _u32p, C.c_int, C.c_int, _f64p, _i32p, C.POINTER(C.c_longlong)])
        self._sig(p + "simpls", C.c_int,
                  [_f64p, _f64p, C.c_int, C.c_int, _u32p, C.c_int, _u32p, C.c_int, C.c_int, _f64p, _f64p])
        self._sig(p + "hardware_threads", C.c_int, [])

    def _sig(self, name, restype, argtypes):
        fn = getattr(self.dll, name)
        fn.restype = restype
        fn.argtypes = argtypes
        short = name[len(self.prefix):]
        if not hasattr(type(self), short):  # do not shadow the Python convenience wrappers
            setattr(self, short, fn)
        setattr(self, name, fn)

    # ---- convenience wrappers ------------------------------------------------------
    def rng_u32(self, seed, n):
        out = np.zeros(n, dtype=np.uint32)
        self.rng_u32_stream(int(seed), n, out)
        return out

    def make_seedvec(self, seed):
        return seedvec_of(self, seed)

    def rng_vec(self, seedvec, n):
        out = np.zeros(n, dtype=np.uint32)
        self.rng_vec_stream(seedvec, n, out)
        return out

    def rng_uniform(self, seedvec, lo, hi, n):
        out = np.zeros(n, dtype=np.float64)
        self.rng_vec_uniform(seedvec, float(lo), float(hi), n, out)
        return out

    def shuffle(self, seedvec, size, ncalls):
        out = np.zeros(size * ncalls, dtype=np.uint32)
        self.shuffle_all(seedvec, size, ncalls, out)
        return out.reshape(ncalls, size)

    def pls_evaluator(self, X, y, seedvec, numReplications=30, maxNComp=0, innerSegments=7, outerSegments=1,
                      testSetSize=0.0, sdfact=1.0, complexityPenalty=0.0):
        n, p = np.asarray(X).shape
        h = self.pls_evaluator_create(_colmajor(X), np.ascontiguousarray(y, dtype=np.float64), n, p,
                                      numReplications, maxNComp, seedvec, innerSegments, outerSegments,
                                      testSetSize, sdfact, complexityPenalty)
        return _Evaluator(self, self.prefix, h)

    def fit_evaluator(self, X, y, seedvec, numSegments=7, statistic="BIC", maxNComp=0, sdfact=1.0):
        n, p = np.asarray(X).shape
        h = self.fit_evaluator_create(_colmajor(X), np.ascontiguousarray(y, dtype=np.float64), n, p, maxNComp,
                                      seedvec, numSegments, STAT[statistic], sdfact)
        return _Evaluator(self, self.prefix, h)

    def lm_evaluator(self, X, y, statistic="BIC"):
        n, p = np.asarray(X).shape
        h = self.lm_evaluator_create(_colmajor(X), np.ascontiguousarray(y, dtype=np.float64), n, p, STAT[statistic])
        return _Evaluator(self, self.prefix, h)

    def simpls_fit(self, X, y, cols, ncomp, rows=None):
        n, p = np.asarray(X).shape
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        rows_a = np.ascontiguousarray(rows if rows is not None else [0], dtype=np.uint32)
        coef = np.zeros(len(cols) * ncomp, dtype=np.float64)
        b0 = np.zeros(ncomp, dtype=np.float64)
        rc = self.simpls(_colmajor(X), np.ascontiguousarray(y, dtype=np.float64), n, p, cols, len(cols), rows_a,
                         0 if rows is None else len(rows_a), ncomp, coef, b0)
        if rc != 0:
            raise RuntimeError("simpls failed rc=%d" % rc)
        return coef.reshape(ncomp, len(cols)).T.copy(), b0


class RefLib(_Lib):
    prefix = "ref_"
    path = REF_SO

    def __init__(self):
        super().__init__()
        self._sig("ref_ga_run", C.c_int,
                  [C.c_void_p] + [C.c_int] * 6 + [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                                   C.c_int, _u32p,
                                                   C.c_int, C.c_longlong, C.POINTER(C.c_int), _u32p, _u32p, _f64p,
                                                   C.c_int, C.POINTER(C.c_int), _f64p, C.POINTER(C.c_ulonglong),
                                                   C.POINTER(C.c_longlong)])
        self._sig("ref_chromosome_kat", C.c_int,
                  [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, _u32p, _u32p, _u32p, _i32p,
                   C.POINTER(C.c_uint32)])

    def ga_run(self, evaluator, seedvec, chromosomeSize, popSize, numGenerations, elitism=10, minVariables=1,
               maxVariables=None, mutationProb=0.01, numThreads=1, maxDuplicateEliminationTries=0,
               badSolutionThreshold=2.0, convergenceThreshold=1e-6, convergenceGenerations=0, crossover=0,
               fitnessScaling=0):
        """Run the reference GA; `evaluator` is an _Evaluator (or a raw Evaluator* handle)."""
        handle = evaluator._h if hasattr(evaluator, "_h") else evaluator
        cap = popSize + elitism + 8
        idx_cap = cap * max(int(maxVariables or chromosomeSize), 1)
        n_res = C.c_int(0)
        res_idx = np.zeros(idx_cap, dtype=np.uint32)
        res_off = np.zeros(cap + 1, dtype=np.uint32)
        res_fit = np.zeros(cap, dtype=np.float64)
        evo_cap = 3 * (numGenerations + 2)
        n_evo = C.c_int(0)
        evo = np.zeros(evo_cap, dtype=np.float64)
        count = C.c_ulonglong(0)
        ns = C.c_longlong(0)
        rc = self.ref_ga_run(handle, chromosomeSize, popSize, numGenerations, elitism, minVariables,
                             int(maxVariables or chromosomeSize), mutationProb, numThreads,
                             maxDuplicateEliminationTries, badSolutionThreshold, convergenceThreshold,
                             convergenceGenerations, crossover, fitnessScaling, seedvec,
                             cap, idx_cap, C.byref(n_res), res_idx, res_off, res_fit,
                             evo_cap, C.byref(n_evo), evo, C.byref(count), C.byref(ns))
        if rc != 0:
            raise RuntimeError("ref_ga_run rc=%d" % rc)
        subsets = [res_idx[res_off[i]:res_off[i + 1]].copy() for i in range(n_res.value)]
        return {
            "subsets": subsets,
            "fitness": res_fit[:n_res.value].copy(),
            "fitnessEvolution": evo[:n_evo.value].reshape(-1, 3).copy(),
            "evaluations": int(count.value),
            "seconds": ns.value * 1e-9,
        }

    def chromosome_kat(self, p, seedvec, popSize=40, minVariables=3, maxVariables=10, mutationProb=0.01, crossover=0):
        idx = np.zeros(8 * p, dtype=np.uint32)
        offsets = np.zeros(7, dtype=np.uint32)
        mutated = np.zeros(2, dtype=np.int32)
        nxt = C.c_uint32(0)
        rc = self.ref_chromosome_kat(p, popSize, minVariables, maxVariables, mutationProb, crossover, seedvec, idx,
                                     offsets, mutated, C.byref(nxt))
        if rc != 0:
            raise RuntimeError("ref_chromosome_kat rc=%d" % rc)
        lists = [idx[offsets[i]:offsets[i + 1]].tolist() for i in range(6)]
        return lists, mutated.tolist(), int(nxt.value)


class OracleLib(_Lib):
    prefix = "orc_"
    path = ORACLE_SO
