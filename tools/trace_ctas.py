#!/usr/bin/env python
"""Per-CTA phase timeline of k_walkers (needs a -DSPAMM_TRACE build: tools/build_variant.sh trace -DSPAMM_TRACE;
SPAMM_B200_LIB=variants/libtrace.so python tools/trace_ctas.py W [split])."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200 import _lib  # noqa: E402
from spamm_b200.synth import full_model_workload  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 512
split = int(sys.argv[2]) if len(sys.argv) > 2 else 0
model, params = full_model_workload(n_pix=8192, n_walkers=W, seed=42)
lib = _lib.load()
p = torch.from_numpy(np.asarray(params)).cuda()
out = torch.empty(W, dtype=torch.float64, device="cuda")
model.plan.set_walker_split(split)
for _ in range(5):
    model.plan.lnposterior(p, out=out)
torch.cuda.synchronize()
n_cta = W * (2 if split == 3 else max(split, 1))
buf = torch.zeros(n_cta * 8, dtype=torch.int64, device="cuda")
lib.spamm_debug_set_trace.argtypes = [ctypes.c_void_p]
_lib.check(lib.spamm_debug_set_trace(ctypes.c_void_p(buf.data_ptr())))
model.plan.lnposterior(p, out=out)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(n_cta, 8).astype(np.int64)
t0 = t[:, 0].min()
start, p0, p1, p3, end, w0end = [(t[:, k] - t0) / 1e3 for k in range(6)]
sm = t[:, 7]
print("W=%d split=%d  CTAs=%d  (us relative to the first CTA start)" % (W, split, n_cta))
print("kernel span %.1f us" % end.max())


def stats(name, x):
    print("  %-34s mean %6.1f  min %6.1f  p50 %6.1f  p90 %6.1f  max %6.1f" %
          (name, x.mean(), x.min(), np.median(x), np.percentile(x, 90), x.max()))


groups = [("all", np.ones(n_cta, bool))]
if split == 3:
    groups = [("red parts", np.arange(n_cta) < W), ("blue parts", np.arange(n_cta) >= W)]
for gname, g in groups:
    print(gname)
    stats("start", start[g])
    stats("P0 (params, prior)", (p0 - start)[g])
    stats("P1 (tables, BC stage)", (p1 - p0)[g])
    stats("P3 (line finish, moments)", (p3 - p1)[g])
    stats("tasks (last warp out)", (end - p3)[g])
    stats("tasks (warp 0 out)", (w0end - p3)[g])
    stats("CTA total", (end - start)[g])
    stats("end", end[g])
per_sm = np.bincount(sm, minlength=148)
busy_end = np.array([end[sm == s].max() if (sm == s).any() else 0 for s in range(148)])
print("CTAs per SM: min %d max %d; SM finish time: min %.1f p50 %.1f max %.1f" %
      (per_sm.min(), per_sm.max(), busy_end.min(), np.median(busy_end), busy_end.max()))
