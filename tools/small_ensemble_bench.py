#!/usr/bin/env python
"""Latency of one ln-posterior batch on small ensembles (config 5 model, 8192 px):
one CTA per walker vs one thread-block cluster per walker (spamm_plan_set_walker_split)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200.synth import full_model_workload  # noqa: E402

model, params = full_model_workload(n_pix=8192, n_walkers=1024, seed=42)
p_all = torch.from_numpy(np.asarray(params)).cuda()
for W in (8, 16, 32, 64, 128, 256, 512, 1024):
    p = p_all[:W].contiguous()
    out = torch.empty(W, dtype=torch.float64, device="cuda")
    row = []
    for split in (0, 2, 4, 8, 1):
        model.plan.set_walker_split(split)
        for _ in range(5):
            model.plan.lnposterior(p, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            model.plan.lnposterior(p, out=out)
        e1.record()
        torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 50 * 1e3)
    print("W=%5d  us/batch: split off %.1f | x2 %.1f | x4 %.1f | x8 %.1f | auto %.1f" % ((W,) + tuple(row)))
