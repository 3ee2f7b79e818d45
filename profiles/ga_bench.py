"""End-to-end genAlg() timing: the reference's SingleThreadPopulation + PLSEvaluator on one host core
(and MultiThreadedPopulation on all cores) vs BatchedPopulation + the B200 engine, same seed, and a
check that both select the same subsets.   usage: python profiles/ga_bench.py [cfg1|cfg3] [generations]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from gaselect_b200 import synth
from oracle import bindings
import test_gpu_dropin as T

which = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
ref = bindings.RefLib()
dll = C.CDLL(os.path.join(os.path.dirname(bindings.REF_SO), "libgaselect_dropin.so"))
dll.dropin_pls_create.restype = C.c_void_p
dll.dropin_pls_create.argtypes = [T._f64p, T._f64p, C.c_int, C.c_int, C.c_int, C.c_int, T._u32p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
dll.dropin_destroy.argtypes = [C.c_void_p]
sv = ref.make_seedvec(123)
if which == "cfg1":   # BASELINE configs[0]: n=100 p=200, population 100, 20 generations, evaluatorPLS defaults
    n, p, pop, gens = 100, 200, 100, int(sys.argv[2]) if len(sys.argv) > 2 else 20
    X, y = synth.gaussian(n, p, seed=777)
    ekw = dict(numReplications=30, maxNComp=0, innerSegments=7, outerSegments=1)
    gkw = dict(elitism=10, minVariables=5, maxVariables=30)
else:                 # BASELINE configs[2]: n=150 p=1000 rdCV 10 x 5/4, <= 10 components, population 500
    n, p, pop, gens = 150, 1000, 500, int(sys.argv[2]) if len(sys.argv) > 2 else 5
    X, y = synth.nir_spectra(n, p, seed=777)
    ekw = dict(numReplications=10, maxNComp=10, innerSegments=4, outerSegments=5)
    gkw = dict(elitism=10, minVariables=10, maxVariables=200)
cores = os.cpu_count()
ev = ref.pls_evaluator(X, y, sv, **ekw)
skip1 = "--skip-1core" in sys.argv  # long runs: compare against the multi-threaded reference driver only
if skip1:
    r1, t1 = None, float("nan")
else:
    t0 = time.perf_counter(); r1 = ref.ga_run(ev, sv, p, pop, gens, numThreads=1, **gkw); t1 = time.perf_counter() - t0
h = dll.dropin_pls_create(T._colmajor(X), np.ascontiguousarray(y), n, p, ekw["numReplications"], ekw["maxNComp"], sv,
                          ekw["innerSegments"], ekw["outerSegments"], 0.0, 1.0, 0.0)
t0 = time.perf_counter(); g = T._batched_run(dll, h, sv, p, pop, gens, **gkw); tg = time.perf_counter() - t0
warm = []
for _ in range(3):  # warm engine: three runs, the median is reported (the engine runs come BEFORE the reference's 16-thread run: measured
    t0 = time.perf_counter(); gT = T._batched_run(dll, h, sv, p, pop, gens, numThreads=cores, **gkw); warm.append(time.perf_counter() - t0)  # after it, in the same process, they took 0.24-0.83 s instead of 0.15)
tgT = sorted(warm)[1]
dll.dropin_destroy(h)
noref = "--no-ref" in sys.argv       # engine timing only (the reference's 16-thread run takes 35 s at config 3)
if noref:
    rT, tT = None, float("nan")
else:
    t0 = time.perf_counter(); rT = ref.ga_run(ev, sv, p, pop, gens, numThreads=cores, **gkw); tT = time.perf_counter() - t0
sameT = None if rT is None else gT["subsets"] == [s.tolist() for s in rT["subsets"]]
print("%s: %d generations x %d chromosomes, %s evaluations" % (which, gens, pop, "?" if rT is None else rT["evaluations"]))
if r1 is not None:
    same = g["subsets"] == [s.tolist() for s in r1["subsets"]]
    err = float(np.max(np.abs(g["fitness"] - r1["fitness"]) / np.abs(r1["fitness"])))
    print("  reference SingleThreadPopulation, 1 core : %8.2f s  (%7.1f evals/s)" % (t1, r1["evaluations"] / t1))
else:
    same, err = None, float("nan")
if rT is not None:
    print("  reference MultiThreadedPopulation, %2d thr : %8.2f s  (%7.1f evals/s)  [different random streams by design]" % (cores, tT, rT["evaluations"] / tT))
print("  BatchedPopulation + B200 engine           : %8.2f s  (%7.1f evals/s)  engine calls %d, engine evaluations %d, passes %d" %
      (tg, g["engine_evaluations"] / tg, g["engine_calls"], g["engine_evaluations"], g["passes"]))
print("  BatchedPopulation, %2d virtual threads (engine warm, median of 3): %7.3f s (%.0f evals/s end to end)  engine calls %d, engine evaluations %d; same subsets as the %d-thread reference run: %s" %
      (cores, tgT, gT["engine_evaluations"] / tgT, gT["engine_calls"], gT["engine_evaluations"], cores, sameT))
print("  same subsets as the 1-core reference run: %s, max rel fitness error %.2e" % (same, err))
import hashlib
print("  devices %s; digest of the selected subsets and their fitness (virtual-thread run): %s" % (os.environ.get("GASELECT_B200_DEVICES", "0"),
      hashlib.sha1(repr(gT["subsets"]).encode() + np.asarray(gT["fitness"], dtype=np.float64).tobytes()).hexdigest()[:16]))
