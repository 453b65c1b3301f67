#!/usr/bin/env python
"""Quick kernel-iteration loop: time k_walkers on config 5 (8192 walkers x 8192 px) and check the
cfg5 golden ln-posteriors, so a kernel edit is judged on speed and parity in one GPU call."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from spamm_b200.synth import full_model_workload, build_model  # noqa: E402
from helpers import Case, lnpost_cases, rel_err  # noqa: E402

z, _ = lnpost_cases()
worst = 0.0
for name in ("cfg5_full", "full_fp6", "cfg4_nc_bc", "cfg2_nc_fe"):
    c = Case(z, name)
    m = build_model(list(c.comps), c.wl, c.flux, c.err, max_walkers=256)
    got = m.lnposterior_batch(c.params)
    r = rel_err(got, c.lnpost).max()
    worst = max(worst, r)
    print("parity %-12s max rel err %.2e" % (name, r))
W = 8192
model, params = full_model_workload(n_pix=8192, n_walkers=W, seed=42)
p = torch.from_numpy(np.asarray(params)).cuda()
out = torch.empty(W, dtype=torch.float64, device="cuda")
for _ in range(5):
    model.plan.lnposterior(p, out=out)
torch.cuda.synchronize()
ts = []
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        model.plan.lnposterior(p, out=out)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 10)
print("k_walkers config5: %.4f ms (min of 5x10; all %s)  %.3g evals/s  checksum %.17g  parity worst %.2e"
      % (min(ts), " ".join("%.4f" % t for t in ts), W / min(ts) * 1e3, float(out.sum().cpu()), worst))
