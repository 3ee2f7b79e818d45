"""Synthetic inputs shared by tests, bench.py and the CPU baselines.

Everything is derived from SplitMix64 so that the same bytes can be regenerated anywhere
(no library RNG).  Matrices are returned as C-ordered (n, p) float64 numpy arrays; the
engine and the checkers take column-major bytes (``np.ascontiguousarray(X.T)``).
"""
import numpy as np

_MASK = (1 << 64) - 1


class SplitMix64:
    """Vectorised SplitMix64 stream: state += golden gamma; output = mix(state)."""

    def __init__(self, seed):
        self.state = np.uint64(seed & _MASK)

    def u64(self, n):
        with np.errstate(over="ignore"):
            steps = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + self.state
            self.state = steps[-1] if n else self.state
            z = steps
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        return z

    def uniform(self, n, lo=0.0, hi=1.0):
        u = (self.u64(n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        return lo + u * (hi - lo)

    def integers(self, n, lo, hi):
        """n integers in [lo, hi] inclusive."""
        return (lo + np.floor(self.uniform(n) * (hi - lo + 1))).astype(np.int64)


def closed_form(n=40, p=30):
    """The closed-form, well-conditioned data of SURVEY.md Appendix C.0."""
    i = np.arange(n, dtype=np.float64)[:, None]
    j = np.arange(p, dtype=np.float64)[None, :]
    X = np.sin(0.37 * (i + 1) * (j + 1)) + 0.05 * np.mod(7 * i + 3 * j, 11) + 0.5 * np.cos(0.11 * (i + 1) + 0.05 * j)
    ii = np.arange(n, dtype=np.float64)
    y = 1.5 + X[:, 2] - 2.0 * X[:, 5] + 0.5 * X[:, 11] + 0.25 * X[:, p - 1] + 0.1 * np.cos(1.3 * (ii + 1))
    return np.ascontiguousarray(X), np.ascontiguousarray(y)


def nir_spectra(n, p, seed=777, noise=1e-2):
    """NIR-like spectra: baseline + slope + 6 Gaussian bands + noise.

    noise = 1e-2 is the parity-gated setting; 1e-3 / 1e-4 are stress sets on which the
    reference itself is not reproducible to 1e-8."""
    g = SplitMix64(seed)
    lam = np.arange(p, dtype=np.float64) / max(p - 1, 1)
    mu = g.uniform(6, 0.0, 1.0)
    sig = g.uniform(6, 0.03, 0.15)
    conc = g.uniform(n * 6, 0.2, 1.2).reshape(n, 6)
    base = g.uniform(n, 1.0, 1.5)
    slope = g.uniform(n, -0.15, 0.15)
    bands = np.exp(-0.5 * ((lam[None, :] - mu[:, None]) / sig[:, None]) ** 2)  # 6 x p
    eps = noise * (g.uniform(n * p * 3).reshape(n, p, 3).sum(axis=2) - 1.5)
    X = base[:, None] + slope[:, None] * lam[None, :] + conc @ bands + eps
    ynoise = g.uniform(n * 3).reshape(n, 3).sum(axis=1) - 1.5
    y = 10.0 + 3.0 * conc[:, 0] - 2.0 * conc[:, 2] + conc[:, 4] + 0.05 * ynoise
    return np.ascontiguousarray(X), np.ascontiguousarray(y)


def gaussian(n, p, seed=777, informative=8, noise=0.5):
    """Gaussian predictors with column SDs cycling 1..5 (as examples/genAlg.R draws them)
    by Box-Muller on SplitMix64; y depends on `informative` evenly spaced columns."""
    g = SplitMix64(seed)
    m = n * p
    u1 = np.maximum(g.uniform(m), 1e-300)
    u2 = g.uniform(m)
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    X = z.reshape(n, p) * (1.0 + np.mod(np.arange(p), 5))[None, :]
    cols = np.linspace(0, p - 1, informative).astype(np.int64)
    beta = g.uniform(informative, -1.0, 1.0)
    e1 = np.maximum(g.uniform(n), 1e-300)
    e2 = g.uniform(n)
    y = X[:, cols] @ beta + noise * np.sqrt(-2.0 * np.log(e1)) * np.cos(2.0 * np.pi * e2)
    return np.ascontiguousarray(X), np.ascontiguousarray(y)


def random_subsets(p, count, kmin, kmax, seed=1):
    """`count` ascending column subsets with sizes uniform in [kmin, kmax] (initial-population
    shape of src/Chromosome.cpp:107-129)."""
    g = SplitMix64(seed)
    sizes = g.integers(count, kmin, kmax)
    out = []
    for k in sizes:
        keys = g.uniform(p)
        out.append(np.sort(np.argsort(keys, kind="stable")[: int(k)]).astype(np.uint32))
    return out


def sweep_subsets(p, count, kbar, seed=1):
    """cfg5 populations: k ~ round(U(0.5, 1.5) * kbar), clipped to [1, p]."""
    g = SplitMix64(seed)
    sizes = np.clip(np.rint(g.uniform(count, 0.5, 1.5) * kbar).astype(np.int64), 1, p)
    out = []
    for k in sizes:
        keys = g.uniform(p)
        out.append(np.sort(np.argsort(keys, kind="stable")[: int(k)]).astype(np.uint32))
    return out
