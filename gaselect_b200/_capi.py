"""ctypes binding of the C ABI in include/gaselect_b200.h (libgaselect_b200.so).

This is the only way the Python host layer reaches the engine.  There is no fallback: when the
shared library is missing the import of the evaluators fails with instructions to build it, and
when no CUDA device is present every evaluate call raises ``EngineError`` (GSB_ERR_CUDA).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgaselect_b200.so")

GSB_OK, GSB_ERR_INVALID_ARGUMENT, GSB_ERR_STATE, GSB_ERR_CUDA, GSB_ERR_MEMORY = range(5)
EVAL_PLS_CV, EVAL_LM, EVAL_PLS_FIT = 1, 2, 3
STAT = {"BIC": 0, "AIC": 1, "adjusted.r.squared": 2, "r.squared": 3}
STATUS_OK, STATUS_EMPTY, STATUS_UNDERFLOW, STATUS_SINGULAR, STATUS_FLAGGED, STATUS_RETRY = range(6)
SEED_WORDS = 624
MAX_DEVICES = 16


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("evaluator", C.c_int32), ("statistic", C.c_int32),
        ("num_replications", C.c_int32), ("inner_segments", C.c_int32), ("outer_segments", C.c_int32),
        ("max_ncomp", C.c_int32), ("test_set_size", C.c_double), ("sdfact", C.c_double),
        ("complexity_penalty", C.c_double), ("num_devices", C.c_int32), ("devices", C.c_int32 * 16), ("reserved", C.c_int32),
        ("tie_band", C.c_double),
    ]


class Info(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("p", C.c_int32), ("max_ncomp", C.c_int32), ("inner_segments", C.c_int32),
        ("outer_segments", C.c_int32), ("num_replications", C.c_int32), ("segmentation_vectors", C.c_int32),
        ("fits_per_evaluation", C.c_int32), ("distinct_fits", C.c_int32), ("distinct_blocks", C.c_int32),
        ("prediction_tasks", C.c_int32), ("sm_count", C.c_int32), ("gram_bytes", C.c_int64),
        ("block_bytes", C.c_int64), ("sdfact_effective", C.c_double), ("setup_ms", C.c_double),
        ("gram_dmma_ms", C.c_double), ("gram_flops", C.c_double),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class Timing(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_double), ("fit_kernel_ms", C.c_double), ("reduce_kernel_ms", C.c_double),
        ("d2h_ms", C.c_double), ("total_ms", C.c_double), ("fit_ctas", C.c_int64), ("kernel_launches", C.c_int32),
        ("retries", C.c_int32), ("fit_tensor_gflop", C.c_double),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_H = C.c_void_p

# name -> (restype, argtypes); every symbol include/gaselect_b200.h declares
SIGNATURES = {
    "gsb_abi_version": (C.c_int, []),
    "gsb_device_count": (C.c_int, [C.POINTER(C.c_int32)]),
    "gsb_create": (C.c_int, [C.POINTER(_H), C.POINTER(Config)]),
    "gsb_destroy": (None, [_H]),
    "gsb_last_error": (C.c_int, [_H, C.c_char_p, C.c_size_t]),
    "gsb_set_data": (C.c_int, [_H, _f64p, _f64p, C.c_int32, C.c_int32]),
    "gsb_seed_vector": (C.c_int, [C.c_uint32, _u32p]),
    "gsb_init_segmentation": (C.c_int, [_H, _u32p]),
    "gsb_set_segmentation": (C.c_int, [_H, _u32p, _u32p, C.c_int32]),
    "gsb_get_segmentation": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "gsb_prepare": (C.c_int, [_H]),
    "gsb_get_info": (C.c_int, [_H, C.POINTER(Info)]),
    "gsb_evaluate_indices": (C.c_int, [_H, _u32p, _u32p, C.c_int32, _f64p, _i32p]),
    "gsb_evaluate_bitsets": (C.c_int, [_H, _u64p, C.c_int32, C.c_int32, _f64p, _i32p]),
    "gsb_population_upload": (C.c_int, [_H, _u32p, _u32p, C.c_int32]),
    "gsb_population_evaluate": (C.c_int, [_H]),
    "gsb_population_enqueue": (C.c_int, [_H]),
    "gsb_set_stream": (C.c_int, [_H, C.c_void_p]),
    "gsb_population_download": (C.c_int, [_H, _f64p, _i32p]),
    "gsb_population_device_results": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]),
    "gsb_last_timing": (C.c_int, [_H, C.POINTER(Timing)]),
    "gsb_deal": (C.c_int, [_u32p, C.c_int32, C.c_int32, C.c_int32, _i32p]),
    "gsb_debug_trace": (C.c_int, [_H, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"), C.c_int32]),
    "gsb_simpls": (C.c_int, [_H, _u32p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, _f64p, _f64p]),
    "gsb_rng_stream": (C.c_int, [_u32p, C.c_int32, _u32p]),
    "gsb_shuffle_all": (C.c_int, [_u32p, C.c_int32, C.c_int32, _u32p]),
}


class EngineError(RuntimeError):
    """A nonzero gsb_error other than an invalid argument (the C++ host turns these into
    std::runtime_error, which surfaces as an R error through BEGIN_RCPP, src/GenAlg.cpp:53,223)."""

    def __init__(self, code, message):
        super().__init__("gaselect_b200 engine error %d: %s" % (code, message))
        self.code = code


_lib = None


def lib():
    """The loaded shared library.  Loud failure when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `make -C gaselect_b200/csrc` (or python -c 'import __graft_entry__ as g; "
                "g.build()'). There is no CPU fallback for this path." % LIB_PATH)
        dll = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(dll, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = dll
    return _lib


def last_error(handle):
    buf = C.create_string_buffer(1024)
    lib().gsb_last_error(handle, buf, len(buf))
    return buf.value.decode("utf-8", "replace")


def check(handle, rc):
    if rc == GSB_OK:
        return
    msg = last_error(handle)
    if rc == GSB_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)  # std::invalid_argument of the reference constructors
    raise EngineError(rc, msg)


def seed_vector(seed):
    """The two-stage seeding of src/GenAlg.cpp:99-103."""
    out = np.zeros(SEED_WORDS, dtype=np.uint32)
    rc = lib().gsb_seed_vector(int(seed) & 0xFFFFFFFF, out)
    if rc != GSB_OK:
        raise EngineError(rc, "gsb_seed_vector")
    return out


def to_csr(subsets):
    """list of column-index sequences -> (idx uint32, offsets uint32)."""
    offsets = np.zeros(len(subsets) + 1, dtype=np.uint32)
    if len(subsets):
        offsets[1:] = np.cumsum([len(s) for s in subsets], dtype=np.uint64).astype(np.uint32)
    idx = np.zeros(max(int(offsets[-1]), 1), dtype=np.uint32)
    if int(offsets[-1]):
        idx[: int(offsets[-1])] = np.concatenate([np.asarray(s, dtype=np.uint32).ravel() for s in subsets if len(s)])
    return idx, offsets
