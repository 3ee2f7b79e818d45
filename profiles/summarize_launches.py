"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total and mean duration."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("gsb::<unnamed>::", "")
    val = float(r[-1].replace(",", ""))
    unit = r[-2]
    ns = val * {"ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(unit, 1.0)
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1; t[1] += ns
allns = sum(t[1] for t in tot.values())
for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-60s launches %5d  total %10.3f ms  mean %9.1f us  share %5.1f%%" % (name[:60], n, ns / 1e6, ns / n / 1e3, 100 * ns / allns))
