"""Generates tests/golden/host_ga_digests.json: outcome digests of the REFERENCE's own GA drivers (SingleThreadPopulation,
MultiThreadedPopulation with real threads; compiled from /root/reference/src where they lie) on the fake evaluator of
ga_host_check.cpp, for every case tests/test_host_ga.py runs BatchedPopulation on.  The multi-threaded driver's
hand-rolled condition-variable hand-over can miss a wake-up when an evaluation takes no time at all, so every run is
guarded by a watchdog and repeated until two runs agree.   usage: python tests/host_ga/make_digests.py"""
import json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import test_host_ga as T

exe = T.build()
out = {}
for case in T.CASES:
    seen = {}
    for attempt in range(40):
        try:
            r = subprocess.run([exe, "ref"] + [str(a) for a in case], capture_output=True, text=True, timeout=20)
        except subprocess.TimeoutExpired:
            continue
        if r.returncode != 0:
            continue
        d = r.stdout.split()[1]
        seen[d] = seen.get(d, 0) + 1
        if seen[d] >= 2:
            out[T.key(case)] = d
            break
    else:
        raise SystemExit("no two agreeing runs for %r: %r" % (case, seen))
    print(T.key(case), out[T.key(case)], flush=True)
json.dump(out, open(os.path.join(os.path.dirname(HERE), "golden", "host_ga_digests.json"), "w"), indent=1, sort_keys=True)
