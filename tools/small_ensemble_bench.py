#!/usr/bin/env python
"""Latency of one ln-posterior batch for 8 ... 4096 walkers (config 5 model, 8192 px): one CTA per
walker vs one thread-block cluster of 2/4/8 CTAs per walker (spamm_plan_set_walker_split) vs the
automatic choice; every split must return the bits of the unsplit kernel."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200.synth import full_model_workload  # noqa: E402

sizes = (8, 16, 32, 64, 128, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096, 8192)
if os.environ.get("SIZES"):
    sizes = tuple(int(x) for x in os.environ["SIZES"].split(","))
check = os.environ.get("NOCHECK") is None
model, params = full_model_workload(n_pix=8192, n_walkers=max(sizes), seed=42)
p_all = torch.from_numpy(np.asarray(params)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("us per batch (mean of 50 back-to-back launches)")
for W in sizes:
    p = p_all[:W].contiguous()
    out = torch.empty(W, dtype=torch.float64, device="cuda")
    row, ref = [], None
    for split in (0, 2, 4, 8, 3, 1):
        model.plan.set_walker_split(split)
        for _ in range(5):
            model.plan.lnposterior(p, out=out)
        torch.cuda.synchronize()
        got = out.cpu().numpy().copy()
        if ref is None:
            ref = got
        assert not check or np.array_equal(ref, got, equal_nan=True), "split %d changes the result at W=%d" % (split, W)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            model.plan.lnposterior(p, out=out)
        e1.record()
        torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 50 * 1e3)
    print("W=%5d  split off %7.1f | x2 %7.1f | x4 %7.1f | x8 %7.1f | parts %7.1f | auto %7.1f   (ideal at 8192-walker rate: %.1f)"
          % ((W,) + tuple(row) + (W * 0.849e3 / 8192,)))
