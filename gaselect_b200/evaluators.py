"""Host-side mirror of the reference's Evaluator plugin surface, backed by the CUDA engine.

Same names, argument meaning and error behaviour as the reference's C++ evaluators:

* ``PLSEvaluator``  <- src/PLSEvaluator.h:25-74 / src/PLSEvaluator.cpp:27-91 (R: ``evaluatorPLS``)
* ``BICEvaluator``  <- src/BICEvaluator.h:24-72 / src/BICEvaluator.cpp:22-91 (R: ``evaluatorFit``)
* ``LMEvaluator``   <- src/LMEvaluator.h:18-50  / src/LMEvaluator.cpp:17-67  (R: ``evaluatorLM``)

``evaluate(columnSubset)`` scores one subset (``Evaluator::evaluate(arma::uvec&)``,
src/Evaluator.h:30) and raises ``EvaluatorException`` where the reference throws
``Evaluator::EvaluatorException`` (src/Evaluator.h:21-25); ``evaluate_population`` is the batch
loop of src/GenAlg.cpp:312-325 as ONE engine call and returns (fitness, status) so that a GA
driver can re-draw failed slots itself (src/MultiThreadedPopulation.cpp:202-206).  Status 4
(``STATUS_FLAGGED``) is not a failure: the fitness is valid, but a comparison of the one-SE rule
(src/PLSEvaluator.cpp:284-308) was a near-tie within the relative band ``tieBand`` (default 1e-9), so
the reference -- summing in another order -- may fit a different number of components.
``getSegmentation()`` is src/Evaluator.h:35-37.  Everything numerical happens in the sm_100a
kernels behind include/gaselect_b200.h; nothing here computes a fit on the CPU.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import EngineError  # noqa: F401  (re-exported)


class EvaluatorException(RuntimeError):
    """Evaluator::EvaluatorException, src/Evaluator.h:21-25."""


_STATUS_MESSAGE = {
    _capi.STATUS_EMPTY: "Can not evaluate empty variable subset",  # src/PLSEvaluator.cpp:74
    _capi.STATUS_UNDERFLOW: "Can not perform PLS regression: Underflow error",  # src/PLSEvaluator.cpp:339
    _capi.STATUS_SINGULAR: "The system could not be solved",  # src/LMEvaluator.cpp:61
}


def _colmajor(X):
    X = np.asarray(X, dtype=np.float64)
    if X.ndim != 2:
        raise ValueError("X must be a matrix")
    return np.ascontiguousarray(X.T).ravel(), X.shape[0], X.shape[1]


class Evaluator:
    """Common engine handle: data, segmentation, batch evaluation."""

    _kind = 0

    def __init__(self, X, y, seed, device=0, devices=None, **cfg):
        """device: one CUDA ordinal; devices: a list of ordinals for a multi-device engine (every batch is dealt to
        the GPUs inside the engine; same results as on one GPU)."""
        lib = _capi.lib()
        c = _capi.Config()
        c.struct_size = C.sizeof(_capi.Config)
        c.device = int(device)
        if devices is not None:
            devices = [int(d) for d in devices]
            if not 1 <= len(devices) <= _capi.MAX_DEVICES:
                raise ValueError("devices must hold 1..%d ordinals" % _capi.MAX_DEVICES)
            c.num_devices = len(devices)
            for i, d in enumerate(devices):
                c.devices[i] = d
        c.evaluator = self._kind
        c.statistic = cfg.get("statistic", 0)
        c.num_replications = cfg.get("num_replications", 1)
        c.inner_segments = cfg.get("inner_segments", 1)
        c.outer_segments = cfg.get("outer_segments", 1)
        c.max_ncomp = cfg.get("max_ncomp", 0)
        c.test_set_size = cfg.get("test_set_size", 0.0)
        c.sdfact = cfg.get("sdfact", 1.0)
        c.complexity_penalty = cfg.get("complexity_penalty", 0.0)
        c.tie_band = cfg.get("tie_band", 0.0)
        self._h = _capi._H()
        self._lib = lib
        rc = lib.gsb_create(C.byref(self._h), C.byref(c))
        if rc != _capi.GSB_OK:
            _capi.check(None, rc)
        try:
            Xc, n, p = _colmajor(X)
            yv = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
            if yv.shape[0] != n:
                raise ValueError("y must have one entry per row of X")
            self.n, self.p = n, p
            self._check(lib.gsb_set_data(self._h, Xc, yv, n, p))
            if self._kind != _capi.EVAL_LM:
                if np.isscalar(seed):
                    seed = _capi.seed_vector(seed)
                seedvec = np.ascontiguousarray(seed, dtype=np.uint32)
                if seedvec.shape != (_capi.SEED_WORDS,):
                    raise ValueError("seed must be one integer or %d words (RNG::SEED_SIZE)" % _capi.SEED_WORDS)
                self._check(lib.gsb_init_segmentation(self._h, seedvec))
        except Exception:
            self.close()
            raise

    # ---- lifetime ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.gsb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        _capi.check(self._h, rc)

    # ---- Evaluator interface ---------------------------------------------------------------
    def getSegmentation(self):
        """std::vector<arma::uvec> Evaluator::getSegmentation(), src/Evaluator.h:35-37."""
        nvec, total = C.c_int32(0), C.c_int64(0)
        self._check(self._lib.gsb_get_segmentation(self._h, None, None, C.byref(nvec), C.byref(total)))
        rows = np.zeros(max(int(total.value), 1), dtype=np.uint32)
        offsets = np.zeros(nvec.value + 1, dtype=np.uint32)
        self._check(self._lib.gsb_get_segmentation(self._h, rows.ctypes.data_as(C.c_void_p),
                                                   offsets.ctypes.data_as(C.c_void_p), C.byref(nvec), C.byref(total)))
        return [rows[offsets[i]:offsets[i + 1]].copy() for i in range(nvec.value)]

    def setSegmentation(self, segmentation):
        """Adopt an externally built segmentation (reference order train, test, ...)."""
        idx, offsets = _capi.to_csr(segmentation)
        self._check(self._lib.gsb_set_segmentation(self._h, idx, offsets, len(segmentation)))

    def prepare(self):
        """Upload the data and run the one-off Gram build; implicit in the first evaluate."""
        self._check(self._lib.gsb_prepare(self._h))
        return self.info()

    def info(self):
        inf = _capi.Info()
        self._check(self._lib.gsb_get_info(self._h, C.byref(inf)))
        return inf.as_dict()

    def timing(self):
        t = _capi.Timing()
        self._check(self._lib.gsb_last_timing(self._h, C.byref(t)))
        return t.as_dict()

    def debug_trace(self, n=64):
        """clock64() stamps of one K1 CTA (profiling aid, needs GSB_TRACE=1 in the environment)."""
        out = np.zeros(n, dtype=np.int64)
        self._check(self._lib.gsb_debug_trace(self._h, out, n))
        return out

    def evaluate_population(self, subsets):
        """fitness, status of every subset; one engine call (src/GenAlg.cpp:312-325)."""
        idx, offsets = _capi.to_csr(subsets)
        return self.evaluate_csr(idx, offsets)

    def evaluate_csr(self, idx, offsets):
        npop = len(offsets) - 1
        fitness = np.zeros(npop, dtype=np.float64)
        status = np.zeros(npop, dtype=np.int32)
        self._check(self._lib.gsb_evaluate_indices(self._h, idx, offsets, npop, fitness, status))
        return fitness, status

    def evaluate_bitsets(self, bitsets):
        """Chromosome bitsets in the reference layout (src/Chromosome.cpp:419-438): array of
        shape (npop, ceil(p / 64)) uint64."""
        bits = np.ascontiguousarray(bitsets, dtype=np.uint64)
        npop, words = bits.shape
        fitness = np.zeros(npop, dtype=np.float64)
        status = np.zeros(npop, dtype=np.int32)
        self._check(self._lib.gsb_evaluate_bitsets(self._h, bits.ravel(), words, npop, fitness, status))
        return fitness, status

    def evaluate(self, columnSubset):
        """double Evaluator::evaluate(arma::uvec &columnSubset), src/Evaluator.h:30."""
        fitness, status = self.evaluate_population([np.asarray(columnSubset, dtype=np.uint32)])
        if status[0] not in (_capi.STATUS_OK, _capi.STATUS_FLAGGED):  # FLAGGED: valid fitness, near-tie in the one-SE rule
            raise EvaluatorException(_STATUS_MESSAGE.get(int(status[0]), "evaluation failed"))
        return float(fitness[0])

    # ---- resident population (bench / multi-GPU host) ----------------------------------------
    def upload_population(self, idx, offsets):
        self._check(self._lib.gsb_population_upload(self._h, idx, offsets, len(offsets) - 1))

    def evaluate_resident(self):
        self._check(self._lib.gsb_population_evaluate(self._h))

    def enqueue_resident(self):
        """Enqueue only (no synchronisation): K1/K2 (or K3) on the engine's stream."""
        self._check(self._lib.gsb_population_enqueue(self._h))

    def set_stream(self, cuda_stream):
        """Run the engine's device work on the caller's cudaStream_t (e.g. torch's current stream)."""
        self._check(self._lib.gsb_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def download_results(self, fitness, status):
        self._check(self._lib.gsb_population_download(self._h, fitness, status))

    def device_results(self):
        """(fitness device address, status device address, npop)."""
        f, s, n = C.c_void_p(), C.c_void_p(), C.c_int32(0)
        self._check(self._lib.gsb_population_device_results(self._h, C.byref(f), C.byref(s), C.byref(n)))
        return f.value, s.value, n.value

    def simpls(self, cols, ncomp, rows=None):
        """The hidden C_simpls entry (src/GenAlg.cpp:351-371): coefficients (k x ncomp,
        cumulative columns) and intercepts of one SIMPLS fit."""
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        k = len(cols)
        coef = np.zeros(k * ncomp, dtype=np.float64)
        b0 = np.zeros(ncomp, dtype=np.float64)
        if rows is None:
            rp, nr = None, 0
        else:
            rows = np.ascontiguousarray(rows, dtype=np.uint32)
            rp, nr = rows.ctypes.data_as(C.c_void_p), len(rows)
        self._check(self._lib.gsb_simpls(self._h, cols, k, rp, nr, int(ncomp), coef, b0))
        return coef.reshape(ncomp, k).T.copy(), b0


class PLSEvaluator(Evaluator):
    """PLSEvaluator(pls, numReplications, maxNComp, seed, verbosity, innerSegments, outerSegments,
    testSetSize, sdfact, complexityPenalty), src/PLSEvaluator.cpp:27-57.  Defaults are those of
    R's evaluatorPLS (R/Evaluator.R:200-201)."""

    _kind = _capi.EVAL_PLS_CV

    def __init__(self, X, y, seed, numReplications=30, maxNComp=0, innerSegments=7, outerSegments=1,
                 testSetSize=0.0, sdfact=1.0, complexityPenalty=0.0, device=0, devices=None, tieBand=0.0):
        super().__init__(X, y, seed, device=device, devices=devices, tie_band=float(tieBand), num_replications=int(numReplications), max_ncomp=int(maxNComp),
                         inner_segments=int(innerSegments), outer_segments=int(outerSegments),
                         test_set_size=float(testSetSize), sdfact=float(sdfact),
                         complexity_penalty=float(complexityPenalty))


class BICEvaluator(Evaluator):
    """BICEvaluator(pls, maxNComp, seed, verbosity, numSegments, stat, sdfact),
    src/BICEvaluator.cpp:22-44 (R: evaluatorFit, R/Evaluator.R:296-340)."""

    _kind = _capi.EVAL_PLS_FIT

    def __init__(self, X, y, seed, numSegments=7, statistic="BIC", maxNComp=0, sdfact=1.0, device=0, devices=None, tieBand=0.0):
        super().__init__(X, y, seed, device=device, devices=devices, tie_band=float(tieBand), inner_segments=int(numSegments), statistic=_capi.STAT[statistic],
                         max_ncomp=int(maxNComp), sdfact=float(sdfact))


class LMEvaluator(Evaluator):
    """LMEvaluator(X, y, statistic, verbosity), src/LMEvaluator.cpp:17-25 (R: evaluatorLM)."""

    _kind = _capi.EVAL_LM

    def __init__(self, X, y, statistic="BIC", device=0, devices=None):
        super().__init__(X, y, None, device=device, devices=devices, statistic=_capi.STAT[statistic])
