"""Attribute the warp-stall samples of an ncu report to CUDA source lines.

usage: python profiles/sass_lines.py REPORT.ncu-rep CUBIN KERNEL_SUBSTRING [TOP]
The SASS page of the report gives samples per instruction; `nvdisasm -g` on the cubin extracted
from the shipped .so (cuobjdump -xelf all) gives the source line of each instruction, in the
same order."""
import csv, io, re, subprocess, sys, collections

rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = rows[1]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
inst = [r for r in rows[2:] if len(r) == len(hdr)]

dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, inside = [], None, False
for l in dis.splitlines():
    if l.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", l):
        inside = kern in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if inside and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
if len(lines) != len(inst):
    print("warning: %d instructions in the report, %d in the cubin" % (len(inst), len(lines)))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r, ln in zip(inst, lines):
    s = int(r[si]); tot += s
    a = agg[ln]; a[0] += s; a[1] += int(r[ii])
    for i, n in stalls:
        if r[i].isdigit(): a[2][n] += int(r[i])
src = {}
print("total samples", tot)
for ln, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if ln and ln[0] not in src:
        try: src[ln[0]] = open("gaselect_b200/csrc/" + ln[0]).read().splitlines()
        except OSError: src[ln[0]] = []
    text = src[ln[0]][ln[1] - 1].strip()[:70] if ln and len(src[ln[0]]) >= ln[1] else ""
    print("%5.1f%% inst=%10d %s:%s | %-70s | %s" % (100.0 * s / tot, n, ln[0] if ln else "?", ln[1] if ln else "?", text,
                                                  " ".join("%s=%d" % kv for kv in st.most_common(3))))
