"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and max time per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
    name = r[ki].split("(")[0][:110]
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += us; a[2] = max(a[2], us)
tot = sum(a[1] for a in agg.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else ""))
print("# per-launch times are cold-cache and serialised; the SHARE of each kernel is what counts")
print(f"{'launches':>8s} {'total_us':>12s} {'share':>7s} {'max_us':>10s}  kernel")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[0]:8d} {a[1]:12.1f} {100 * a[1] / tot:6.1f}% {a[2]:10.1f}  {name}")
