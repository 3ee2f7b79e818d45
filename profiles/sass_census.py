"""Per-kernel SASS mnemonic census of the shipped library: `cuobjdump -sass gaselect_b200/libgaselect_b200.so` folded to
instructions per kernel and the counts of the mnemonics that prove which units the code uses (DMMA = FP64 tensor cores,
UBLKCP / SYNCS = TMA bulk copies on mbarriers, LDGSTS = cp.async, LDTM / STTM = tensor memory, UTMALDG = tensor-map TMA).
usage: python profiles/sass_census.py > profiles/r02_sass_census.txt"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "gaselect_b200", "libgaselect_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
watch = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "LDGSTS", "UBLKCP", "SYNCS", "UTMALDG", "LDTM", "STTM", "BAR", "ATOMG", "RED", "CALL"]
cur, total, hist = None, collections.Counter(), collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur).replace("gsb::", "")
        cur = re.sub(r"\(.*\)$", "", cur)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        total[cur] += 1
        hist[cur][m.group(1)] += 1
print("SASS census of gaselect_b200/libgaselect_b200.so (sm_100a), %d kernels, %d instructions" % (len(total), sum(total.values())))
print("%-64s %7s  %s" % ("kernel", "instr", "  ".join("%s" % w for w in watch)))
allh = collections.Counter()
for k in sorted(total, key=lambda k: -total[k]):
    print("%-64s %7d  %s" % (k[:64], total[k], "  ".join("%*d" % (len(w), hist[k][w]) for w in watch)))
    allh.update(hist[k])
print("%-64s %7d  %s" % ("ALL", sum(total.values()), "  ".join("%*d" % (len(w), allh[w]) for w in watch)))
