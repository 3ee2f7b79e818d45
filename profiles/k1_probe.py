"""Small driver for profiling K1: one population of fixed subset size on the cfg3 data set.
usage: python profiles/k1_probe.py K [POP] [STEPS]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gaselect_b200 import synth, evaluators as E, _capi

k = int(sys.argv[1]); pop = int(sys.argv[2]) if len(sys.argv) > 2 else 296; steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
X, y = synth.nir_spectra(150, 1000, seed=777)
subsets = synth.random_subsets(1000, pop, k, k, seed=1)
idx, off = _capi.to_csr(subsets)
ev = E.PLSEvaluator(X, y, 123, numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5)
ev.upload_population(idx, off)
ts = []
for _ in range(steps):
    ev.evaluate_resident(); ts.append(ev.timing()["fit_kernel_ms"])
fits = pop * 150
print("k=%d pop=%d K1 ms=%s  us/fit-slot(148 SMs)=%.2f ctas=%d" % (k, pop, ["%.3f" % t for t in ts], min(ts) * 1e3 * 148 / fits, ev.timing()["fit_ctas"]))
tr = ev.debug_trace(64)
if tr[32:].any():
    t0 = tr[0]
    print('tile0:', (tr[:32][tr[:32] > 0] - t0).tolist())
    print('tile1:', (tr[32:][tr[32:] > 0] - t0).tolist())
    tr = tr[:32]
if tr.any():
    d = np.diff(tr[tr > 0])
    print("phase clocks:", d.tolist(), "total", int(tr[tr > 0][-1] - tr[0]))
