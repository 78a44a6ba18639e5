"""Per-kernel SASS statistics of the built library: instruction count, local-memory ops, TMA ops."""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "varpro_b200/libvarpro_b200.so"
filt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None
stats = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = dict(n=0, local=0, tma=0, dfma=0, lds=0, shfl=0, bar=0)
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        st = stats[cur]
        st["n"] += 1
        if re.search(r"\b(STL|LDL)\b", line): st["local"] += 1
        if "UBLKCP" in line: st["tma"] += 1
        if re.search(r"\bD(FMA|ADD|MUL)\b", line): st["dfma"] += 1
        if re.search(r"\bLDS\b", line): st["lds"] += 1
        if "SHFL" in line: st["shfl"] += 1
        if "BAR.SYNC" in line: st["bar"] += 1
for k, v in stats.items():
    if filt in k:
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:90]
        print(f"{v['n']:6d} instr  local={v['local']:3d} tma={v['tma']} dfp={v['dfma']:4d} lds={v['lds']:3d} shfl={v['shfl']:3d} bar={v['bar']:2d}  {name}")
