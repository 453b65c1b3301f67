#!/usr/bin/env python
"""One ln-posterior batch of W walkers: one CTA per walker (split 0) against the automatic choice, which for a
few waves of CTAs runs the walkers of the last, partial wave as two independent CTAs each (mixed launch)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200.synth import full_model_workload  # noqa: E402

sizes = tuple(int(x) for x in os.environ.get("SIZES", "512,768,1024,1536,2048,3072,4096,8192").split(","))
model, params = full_model_workload(n_pix=8192, n_walkers=max(sizes), seed=42)
p_all = torch.from_numpy(np.asarray(params)).cuda()
for W in sizes:
    p = p_all[:W].contiguous()
    out = torch.empty(W, dtype=torch.float64, device="cuda")
    row, ref = [], None
    for split in (0, 1):
        model.plan.set_walker_split(split)
        for _ in range(5):
            model.plan.lnposterior(p, out=out)
        torch.cuda.synchronize()
        got = out.cpu().numpy().copy()
        ref = got if ref is None else ref
        assert np.array_equal(ref, got), "split %d changes the result at W=%d" % (split, W)
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                model.plan.lnposterior(p, out=out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 50 * 1e3)
        row.append(best)
    print("W=%5d  one CTA per walker %7.1f us | automatic (SPAMM_PARTS_MIXED=%s) %7.1f us   (8192-walker rate: %.1f)"
          % (W, row[0], os.environ.get("SPAMM_PARTS_MIXED", "1"), row[1], W * 0.849e3 / 8192))
