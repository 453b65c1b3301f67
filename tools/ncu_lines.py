#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall-sample share of a profiled kernel
(ncu --import-source on, compiled with -lineinfo; every source file of the translation unit).
usage: ncu_lines.py rep [n_ctas] [top] [--by-line]"""
import csv
import os
import subprocess
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
rep = args[0]
nw = float(args[1]) if len(args) > 1 else 8192.0
top = int(args[2]) if len(args) > 2 else 60
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
lines, fname, col = [], "?", None
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == "File Name" and len(r) > 1:
        fname = os.path.basename(r[1])
        continue
    if r[0] == "Line No":
        col = {}
        for i, n in enumerate(r):
            col.setdefault(n, i)
        continue
    if r[0] == "Address":          # the SASS-only section that follows the per-file ones
        col = None
        continue
    if col and r[0].strip().isdigit():
        try:
            lines.append((fname, int(r[0]), r[1], float(r[col["Instructions Executed"]]), float(r[col["# Samples"]])))
        except (ValueError, IndexError):
            pass
ti = sum(l[3] for l in lines) or 1.0
ts = sum(l[4] for l in lines) or 1.0
print("total warp instr/CTA %.0f, samples %d" % (ti / nw, ts))
key = (lambda l: (l[0], l[1])) if "--by-line" in sys.argv else (lambda l: -l[3])
for fn, ln, s, e, smp in sorted(lines, key=key)[:top]:
    print("%-20s %5d  inst/CTA %8.1f (%4.1f%%)  smp %4.1f%%  %s" % (fn, ln, e / nw, 100 * e / ti, 100 * smp / ts, s.strip()[:100]))
