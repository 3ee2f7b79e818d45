"""K0 (one-off Gram build) timing at a BASELINE shape.   usage: python profiles/k0_probe.py [n p]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaselect_b200 import synth, evaluators as E
n, p = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (500, 5000)
X, y = synth.nir_spectra(n, p, seed=778)
for rep in range(2):
    ev = E.PLSEvaluator(X, y, 123, numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5)
    i = ev.prepare()
    print("n=%d p=%d: K0 %.2f ms, DMMA SYRK %.3f ms = %.2f TFLOP/s on %.2f GFLOP (lower triangle), %d units; tables %.1f GB" % (
        n, p, i["setup_ms"], i["gram_dmma_ms"], i["gram_flops"] / i["gram_dmma_ms"] / 1e9, i["gram_flops"] / 1e9,
        1 + i["distinct_blocks"], i["gram_bytes"] / 1e9))
    ev.close()
