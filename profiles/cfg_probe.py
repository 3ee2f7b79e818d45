"""Engine timing for the BASELINE evaluator configurations other than the bench workload.
usage: python profiles/cfg_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gaselect_b200 import synth, evaluators as E

def run(name, ev, subsets, reps=5):
    t0 = time.perf_counter(); info = ev.prepare(); tp = time.perf_counter() - t0
    ev.evaluate_population(subsets)
    ts, tk = [], []
    for _ in range(reps):
        t0 = time.perf_counter(); f, s = ev.evaluate_population(subsets); ts.append(time.perf_counter() - t0)
        t = ev.timing(); tk.append(t["fit_kernel_ms"] + t["reduce_kernel_ms"])
    print("%-34s prepare %.3f s (K0 %.2f ms) | %d chromosomes: call %.2f ms, kernels %.2f ms -> %.0f evals/s e2e, %.0f device | launches %d, fits/eval %d" % (
        name, tp, info["setup_ms"], len(subsets), 1e3 * min(ts), min(tk), len(subsets) / min(ts), len(subsets) / (min(tk) * 1e-3),
        ev.timing()["kernel_launches"], info["fits_per_evaluation"]))
    ev.close()

X, y = synth.gaussian(100, 200, seed=777)
run("cfg1 evaluatorPLS defaults p=200", E.PLSEvaluator(X, y, 123), synth.random_subsets(200, 100, 5, 30, seed=7))
X, y = synth.gaussian(200, 500, seed=777)
run("cfg2 evaluatorLM BIC p=500", E.LMEvaluator(X, y, "BIC"), synth.random_subsets(500, 200, 5, 50, seed=6))
run("cfg2' evaluatorFit BIC p=500", E.BICEvaluator(X, y, 123, numSegments=7, statistic="BIC"), synth.random_subsets(500, 200, 5, 50, seed=6))
X, y = synth.nir_spectra(150, 1000, seed=777)
run("cfg3 evaluatorPLS rdCV p=1000", E.PLSEvaluator(X, y, 123, numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5), synth.random_subsets(1000, 500, 10, 200, seed=42))
X, y = synth.nir_spectra(500, 5000, seed=778)
run("cfg4 evaluatorPLS rdCV p=5000 /8", E.PLSEvaluator(X, y, 123, numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5), synth.random_subsets(5000, 250, 20, 200, seed=43), reps=3)
