#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + per-region stall breakdown from the source page.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out_prefix]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
prefix = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['gpu__time_duration.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.per_cycle_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__cycles_elapsed.max', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']
lines = ["metric,unit," + ",".join("launch%d" % (i + 1) for i in range(len(rows) - 2))]
for i, h in enumerate(hdr):
    if h in keep:
        lines.append("%s,%s,%s" % (h, units[i], ",".join(r[i] for r in rows[2:])))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hidx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hidx[0]]
end = hidx[1] - 1 if len(hidx) > 1 else len(rows)
body = rows[hidx[0] + 1:end]
col = {n: i for i, n in enumerate(h)}


def f(r, n):
    try:
        return float(r[col[n]])
    except Exception:
        return 0.0


stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(f(r, '# Samples') for r in body)
toti = sum(f(r, 'Instructions Executed') for r in body)
mx = max(f(r, 'Instructions Executed') for r in body)
hot = [r for r in body if f(r, 'Instructions Executed') >= 0.95 * mx]
hs = sum(f(r, '# Samples') for r in hot)
lines.append("")
lines.append("source page: %d SASS rows, %d samples, %.4g warp instructions" % (len(body), tot, toti))
lines.append("hot loop: %d instructions/iter, %.1f%% of executed instructions, %.1f%% of samples" %
             (len(hot), 100 * sum(f(r, 'Instructions Executed') for r in hot) / toti, 100 * hs / tot))
agg = {n: sum(f(r, n) for r in body) for n in stalls}
lines.append("all samples by stall: " + ", ".join("%s %.1f%%" % (k.replace('stall_', ''), 100 * v / tot)
             for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
agg = {n: sum(f(r, n) for r in hot) for n in stalls}
lines.append("hot-loop samples by stall: " + ", ".join("%s %.1f%%" % (k.replace('stall_', ''), 100 * v / hs)
             for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
segs, cur = [], None
for idx, r in enumerate(body):
    lvl = f(r, 'Instructions Executed')
    if cur is None or abs(lvl - cur['lvl']) > 0.2 * max(lvl, cur['lvl'], 1):
        cur = {'lvl': lvl, 'start': idx, 'n': 0, 'samples': 0, 'st': {}}
        segs.append(cur)
    cur['n'] += 1
    cur['samples'] += f(r, '# Samples')
    cur['end'] = idx
    for s in stalls:
        cur['st'][s] = cur['st'].get(s, 0) + f(r, s)
lines.append("regions with > 0.8% of samples (SASS row range, executions per instruction, top stalls):")
for s in sorted([s for s in segs if s['samples'] > 0.008 * tot], key=lambda s: -s['samples'])[:20]:
    top = sorted(s['st'].items(), key=lambda kv: -kv[1])[:3]
    lines.append("  rows %5d-%5d n=%4d execs=%10.0f samples=%5.1f%%  %s" % (
        s['start'], s['end'], s['n'], s['lvl'], 100 * s['samples'] / tot,
        ", ".join("%s %.0f%%" % (k.replace('stall_', ''), 100 * v / max(s['samples'], 1)) for k, v in top)))
txt = "\n".join(lines)
print(txt)
if prefix:
    open(prefix + "_summary.txt", "w").write(txt + "\n")
