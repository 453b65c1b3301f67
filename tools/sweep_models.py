#!/usr/bin/env python
"""Exploratory version of tests/test_gpu_parity.py::test_randomised_models_vs_oracle: many more seeds,
prints the worst relative error and every failing model instead of asserting.
usage: sweep_models.py first_seed n_seeds"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as t  # noqa: E402
from helpers import rel_err  # noqa: E402
from oracle.spamm_numpy import OracleModel  # noqa: E402
from spamm_b200.plan import ModelPlan  # noqa: E402
from spamm_b200.Spectrum import Spectrum  # noqa: E402
from spamm_b200.components.ReddeningLaw import Extinction  # noqa: E402

s0, n = int(sys.argv[1]), int(sys.argv[2])
worst, bad = 0.0, 0
for seed in range(s0, s0 + n):
    wl, comps, ext_law, bc, bpc, line_type, broken, rng = t._sweep_case(seed)
    P = len(wl)
    lo, hi = [], []
    for c in comps:
        l, h = t._SWEEP_BOUNDS[c]
        if c == "PL" and broken:
            l, h = [wl[0], 0., -3., -3.], [wl[-1], 1e-13, 3., 3.]
        lo += list(l)
        hi += list(h)
    lo, hi = np.array(lo), np.array(hi)
    params = lo + (hi - lo) * rng.rand(8, len(lo))
    params[7, rng.randint(len(lo))] = 2 * hi.max()
    d = 1e-14 * (wl / wl[P // 2]) ** -1.5 * (1 + 0.05 * rng.standard_normal(P))
    e = 0.05 * np.abs(d) + 1e-18
    o = OracleModel(wl, d, e, comps=comps, bc=bc, bpc=bpc, broken_pl=broken, line_type=line_type, ext_law=ext_law)
    o.set_bounds(lo, hi)
    with np.errstate(all="ignore"):
        want = o.ln_posterior_batch(params)
    sp = Spectrum(wl, d, e)
    cs = t._components([c for c in comps if c != "EXT"], bc=bc, bpc=bpc, broken=broken, line_type=line_type)
    if ext_law:
        cs.append(Extinction(**{ext_law: True}))
    tag = (seed, P, round(wl[0]), round(wl[-1]), comps, bc, bpc, line_type, broken, ext_law)
    try:
        for cc in cs:
            cc.initialize(sp)
        plan = ModelPlan(cs, sp, with_priors=False, bounds=(lo, hi), max_walkers=8)
        got = plan.lnposterior_host(params)
        status = plan.status()
    except Exception as ex:  # noqa: BLE001
        print("EXC", tag, repr(ex)[:200])
        bad += 1
        continue
    r = rel_err(got, want).max()
    if not (r < 1e-10) or status != 0:
        bad += 1
        print("BAD", tag, "rel %.3e status %d" % (r, status), got, want)
    else:
        worst = max(worst, r)
print("seeds %d..%d: %d bad, worst passing rel err %.3e" % (s0, s0 + n - 1, bad, worst))
