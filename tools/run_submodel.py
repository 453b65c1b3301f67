#!/usr/bin/env python
"""Run the fused kernel a few times on a sub-model of config 5 (for ncu):  run_submodel.py BC"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200.synth import synthetic_spectrum, config_grid, build_model
comps = sys.argv[1].split("+")
wl = config_grid(8192)
_, flux, err = synthetic_spectrum(("PL", "FE", "HOST", "BC"), wl)
W = 8192
m = build_model(comps, wl, flux, err, max_walkers=W)
np.random.seed(42)
p = torch.from_numpy(np.array(m.initial_walkers(W))).cuda()
out = torch.empty(W, dtype=torch.float64, device="cuda")
for _ in range(5):
    m.plan.lnposterior(p, out=out)
torch.cuda.synchronize()
print("ok", float(out[0]))
