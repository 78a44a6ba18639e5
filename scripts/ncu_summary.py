"""Summarise an .ncu-rep: key raw metrics per launch, plus per-kernel stall breakdown and hot instructions."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'smsp__cycles_elapsed.avg.per_second',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum']
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print('----')
    for i in idx:
        print(f"  {hdr[i]:68s} {r[i][:70]} {units[i]}")

names = sorted({r[hdr.index('Kernel Name')].split('(')[0] for r in rows[2:]})
for name in names:
    key = name.replace('void ', '').split('<')[0]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + key.split('::')[-1]],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) < 3:
        continue
    h = srows[1]
    data = []
    for r in srows[2:]:
        if len(r) != len(h):
            continue
        if r[0] == 'Address':
            break
        data.append(r)
    si = h.index('# Samples')
    ie = h.index('Instructions Executed')
    tot = sum(int(r[si]) for r in data)
    print(f"\n== {name}: {len(data)} SASS instr, {sum(int(r[ie]) for r in data)} warp-instr executed, {tot} samples")
    stall_cols = [i for i, hh in enumerate(h) if hh.startswith('stall_') and 'Not Issued' not in hh]
    agg = {h[i]: 0 for i in stall_cols}
    for r in data:
        for i in stall_cols:
            try:
                agg[h[i]] += int(r[i])
            except ValueError:
                pass
    print("  stalls:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = sorted(range(len(data)), key=lambda n: -int(data[n][si]))[:topn]
    for n in sorted(order):
        r = data[n]
        why = [(h[i][6:], r[i]) for i in stall_cols if r[i] not in ('0', '')]
        why.sort(key=lambda t: -int(t[1]))
        print(f"  #{n:5d} samples={r[si]:>5s} exec={r[ie]:>7s}  {r[1][:70]:70s} {why[:3]}")
