#!/usr/bin/env python
"""Time the fused kernel on sub-models of config 5 (which part of the step costs what)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200.synth import synthetic_spectrum, config_grid, build_model
wl = config_grid(8192)
_, flux, err = synthetic_spectrum(("PL", "FE", "HOST", "BC"), wl)
W = 8192
for comps in (("PL", "FE", "HOST", "BC"), ("BC",), ("PL", "FE", "HOST"), ("PL",), ("FE", "HOST"), ("HOST",)):
    m = build_model(list(comps), wl, flux, err, max_walkers=W)
    np.random.seed(42)
    p = torch.from_numpy(np.array(m.initial_walkers(W))).cuda()
    out = torch.empty(W, dtype=torch.float64, device="cuda")
    for _ in range(3):
        m.plan.lnposterior(p, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m.plan.lnposterior(p, out=out)
    e1.record(); torch.cuda.synchronize()
    print("%-28s %.3f ms" % ("+".join(comps), e0.elapsed_time(e1) / 10))
