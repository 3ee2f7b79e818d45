"""Shared helpers for the parity tests: rebuild a golden case's inputs from its spec."""
import numpy as np

from gaselect_b200 import synth


def make_data(spec):
    spec = dict(spec)
    kind = spec.pop("kind")
    return {"closed_form": synth.closed_form, "gaussian": synth.gaussian, "nir": synth.nir_spectra}[kind](**spec)


def make_subsets(spec, p):
    if spec["kind"] == "list":
        return [np.asarray(s, dtype=np.uint32) for s in spec["lists"]]
    return synth.random_subsets(p, spec["count"], spec["kmin"], spec["kmax"], spec["seed"])


def make_checker_evaluator(lib, spec, X, y, seedvec):
    """Evaluator on a CPU checker library (oracle.bindings.RefLib / OracleLib)."""
    spec = dict(spec)
    kind = spec.pop("kind")
    if kind == "pls":
        return lib.pls_evaluator(X, y, seedvec, **spec)
    if kind == "fit":
        return lib.fit_evaluator(X, y, seedvec, **spec)
    return lib.lm_evaluator(X, y, **spec)


def seg_digest(seg):
    return [int((np.asarray(s, dtype=np.uint64) * np.arange(1, len(s) + 1, dtype=np.uint64)).sum() % (1 << 61)) for s in seg]


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
