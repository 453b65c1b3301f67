#!/usr/bin/env python
"""Kernel time of NC + BalmerCombined with Gaussian line profiles (generic instantiation), 4096 walkers x 8192 px."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200.synth import synthetic_spectrum, config_grid, build_model  # noqa: E402
from spamm_b200.utils.parse_pars import parse_pars  # noqa: E402
from spamm_b200.components.BalmerContinuumCombined import BalmerCombined  # noqa: E402

P, W = 8192, 4096
wl = config_grid(P)
_, flux, err = synthetic_spectrum(("PL", "BC"), wl)
for lt in ("lorentzian", "gaussian"):
    m = build_model(["PL", "BC"], wl, flux, err, max_walkers=W)
    pars = parse_pars()["balmer_continuum"]
    pars["bc_line_type"] = lt
    m.components[-1] = BalmerCombined(pars=pars, BalmerContinuum=True, BalmerPseudocContinuum=True)
    m.data_spectrum = m.data_spectrum
    np.random.seed(42)
    p = torch.from_numpy(np.array(m.initial_walkers(W))).cuda()
    out = torch.empty(W, dtype=torch.float64, device="cuda")
    for _ in range(3):
        m.plan.lnposterior(p, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m.plan.lnposterior(p, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%-10s %.3f ms  %.3g evals/s  checksum %.17g" % (lt, ms, W / ms * 1e3, float(out.sum().cpu())))
