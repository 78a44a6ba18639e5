"""Join an .ncu-rep's SASS-level stall samples with nvdisasm line info of the matching cubin.

usage: ncu_by_line.py REPORT.ncu-rep OBJECT.o MANGLED_SUBSTRING [min_samples]
The report's kernel must have been built from the same sources as OBJECT (instruction counts are checked).
Prints samples / executed warp-instructions per (file, line) and per file, sorted by line.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, key = sys.argv[1:4]
min_s = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + (sys.argv[5] if len(sys.argv) > 5 else key.split("ILi")[0][-16:])],
                     capture_output=True, text=True).stdout
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# instruction index -> (file, line) for the wanted kernel
lines, cur, on = [], ("?", 0), False
for l in dis:
    if l.startswith(".text."):
        on = key in l
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        lines.append((cur, l.split("*/", 1)[1].strip()))
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
data = [r for r in rows[2:] if len(r) == len(h) and r[0] != "Address"]
si, ie = h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, hh in enumerate(h) if hh.startswith("stall_") and "Not Issued" not in hh]
print(f"ncu instr {len(data)}, nvdisasm instr {len(lines)}")
n = min(len(data), len(lines))
agg = collections.OrderedDict()
for k in range(n):
    a = agg.setdefault(lines[k][0], [0, 0, collections.Counter()])
    a[0] += int(data[k][si]); a[1] += int(data[k][ie])
    for i in stall_cols:
        if data[k][i] not in ("", "0"):
            a[2][h[i][6:]] += int(data[k][i])
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
byfile = collections.Counter(); byfile_i = collections.Counter()
for (f, ln), a in agg.items():
    byfile[f] += a[0]; byfile_i[f] += a[1]
print(f"total samples {tot}, warp-instr {toti}")
for f, s in byfile.most_common():
    print(f"  {f:28s} samples {s:9d} ({100*s/tot:5.1f} %)  instr {byfile_i[f]:12d} ({100*byfile_i[f]/toti:5.1f} %)")
srcs = {}
for (f, ln), a in sorted(agg.items()):
    if a[0] < min_s:
        continue
    if f not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "varpro_b200", "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][ln - 1].strip()[:80] if 0 < ln <= len(srcs[f]) else ""
    why = ", ".join(f"{k}={v}" for k, v in a[2].most_common(3))
    print(f"{f:24s}:{ln:4d} s={a[0]:7d} ({100*a[0]/tot:4.1f}%) i={a[1]:10d} ({100*a[1]/toti:4.1f}%) {text:80s} [{why}]")
