#!/usr/bin/env python
"""Kernel time of the BASELINE.json configs 1-5 at their stated walker counts (device-resident params)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200.synth import synthetic_spectrum, config_grid, build_model  # noqa: E402

CONFIGS = (("cfg2 NC+Fe, 1024 walkers x 4096 px", ("PL", "FE"), 4096, 1024),
           ("cfg3 NC+host, 2048 walkers x 8192 px", ("PL", "HOST"), 8192, 2048),
           ("cfg4 NC+Balmer, 4096 walkers x 8192 px", ("PL", "BC"), 8192, 4096),
           ("cfg5 full, 8192 walkers x 8192 px", ("PL", "FE", "HOST", "BC"), 8192, 8192),
           ("full + MW extinction, 8192 x 8192", ("PL", "FE", "HOST", "BC", "MW_EXT"), 8192, 8192),
           ("NC+Fe+host, 8192 x 8192", ("PL", "FE", "HOST"), 8192, 8192),
           ("NC+Fe+Balmer, 8192 x 8192", ("PL", "FE", "BC"), 8192, 8192),
           ("NC+host+Balmer, 8192 x 8192", ("PL", "HOST", "BC"), 8192, 8192))
for name, comps, P, W in CONFIGS:
    wl = config_grid(P)
    base = tuple(c for c in comps if not c.endswith("_EXT"))
    _, flux, err = synthetic_spectrum(base, wl)
    m = build_model(list(base), wl, flux, err, max_walkers=W)
    if len(base) != len(comps):
        from spamm_b200.components.ReddeningLaw import Extinction
        m.components.append(Extinction(MW=True))
        m.data_spectrum = m.data_spectrum
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
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-42s %.3f ms  %.3g evals/s  specialised=%s" % (name, ms, W / ms * 1e3, m.plan.specialised))
