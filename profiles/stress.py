"""Parity on the stress sets of SURVEY.md 8(d)/Appendix B: NIR-like spectra with spectral noise 1e-2 (the
parity-gated setting), 1e-3 and 1e-4, 300 random subsets, k in [5, 60], config-3 CV with 2 replications.
Engine vs the reference build, next to the reference's own reproducibility band (the plain-C oracle, which
differs from the reference build only in summation order).  Also: how many of the deviations above 1e-8 (and above 1e-4: a
flipped optNComp) carry GSB_STATUS_FLAGGED, the near-tie flag of the one-SE rule, at the default band (1e-9) and at 1e-6."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gaselect_b200 import synth, evaluators as E
from oracle import bindings
ref, orc = bindings.RefLib(), bindings.OracleLib()
kw = dict(numReplications=2, maxNComp=10, innerSegments=4, outerSegments=5)
sv = ref.make_seedvec(123)
subsets = synth.random_subsets(1000, 300, 5, 60, seed=17)
print("noise | engine vs reference: max rel err, #>1e-8 | oracle vs reference (reproducibility band): max rel err, #>1e-8")
for noise in (1e-2, 1e-3, 1e-4):
    X, y = synth.nir_spectra(150, 1000, seed=777, noise=noise)
    fr, sr = ref.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=os.cpu_count())
    fo, so = orc.pls_evaluator(X, y, sv, **kw).evaluate(subsets, threads=os.cpu_count())
    with E.PLSEvaluator(X, y, sv, **kw) as ev:
        fg, sg = ev.evaluate_population(subsets)
    with E.PLSEvaluator(X, y, sv, tieBand=1e-6, **kw) as ev:
        fg6, sg6 = ev.evaluate_population(subsets)
    eg = np.abs(fg - fr) / np.abs(fr); eo = np.abs(fo - fr) / np.abs(fr)
    ok = np.isin(sg, (0, 4)) == (sr == 0)
    print("%5.0e | %.2e, %3d of 300 | %.2e, %3d of 300   (statuses consistent: %s)" % (noise, eg.max(), int((eg > 1e-8).sum()), eo.max(), int((eo > 1e-8).sum()), bool(ok.all())))
    print("      | near-tie flag, band 1e-9: %d flagged, %d of the %d deviations > 1e-8, %d of the %d > 1e-4 | band 1e-6: %d flagged, %d of > 1e-8, %d of > 1e-4" % (
        int((sg == 4).sum()), int(((sg == 4) & (eg > 1e-8)).sum()), int((eg > 1e-8).sum()), int(((sg == 4) & (eg > 1e-4)).sum()), int((eg > 1e-4).sum()),
        int((sg6 == 4).sum()), int(((sg6 == 4) & (eg > 1e-8)).sum()), int(((sg6 == 4) & (eg > 1e-4)).sum())))
