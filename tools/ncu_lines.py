#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall-sample share of a profiled kernel
(ncu --import-source on, compiled with -lineinfo).  usage: ncu_lines.py rep [n_ctas] [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; nw = float(sys.argv[2]) if len(sys.argv) > 2 else 8192.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
h = rows[hi[0]]
end = hi[1] - 2 if len(hi) > 1 else len(rows)
col = {}
for i, n in enumerate(h):
    col.setdefault(n, i)
lines = []
for r in rows[hi[0] + 1:end]:
    if r and r[0].strip().isdigit():
        try:
            lines.append((int(r[0]), r[1], float(r[col["Instructions Executed"]]), float(r[col["# Samples"]])))
        except ValueError:
            pass
ti = sum(l[2] for l in lines); ts = sum(l[3] for l in lines)
print("total warp instr/CTA %.0f, samples %d" % (ti / nw, ts))
for ln, s, e, smp in sorted(lines, key=lambda l: -l[2])[:top]:
    print("%5d  inst/CTA %8.1f (%4.1f%%)  smp %4.1f%%  %s" % (ln, e / nw, 100 * e / ti, 100 * smp / ts, s.strip()[:110]))
