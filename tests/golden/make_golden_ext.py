#!/usr/bin/env python
"""Golden fixtures for the Extinction component (spamm/components/ReddeningLaw.py), generated
from the UNMODIFIED reference under oracle/refshim.py (build container only):

  tests/golden/extinction.npz   extinction curves of the five laws on three grids, and
                                ln_posterior()/model_flux() of models that end with an Extinction
                                component (Model.py:358-361, 384-395)

    python tests/golden/make_golden_ext.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (loads the reference through the shim)
from oracle import refshim  # noqa: E402

LAWS = ("MW", "AGN", "LMC", "SMC", "Calzetti")


def make():
    from spamm.components.ReddeningLaw import Extinction
    ref = mg.ref
    out = {}
    grids = {"g4096": mg.grid(4096), "wide": np.linspace(900.0, 25000.0, 1500),
             "test7": np.arange(1000, 10000, 0.5)[::7]}
    for gname, wl in grids.items():
        sp = ref.Spectrum(wl, wl, wl)
        out["curve/%s/wl" % gname] = wl
        for law in LAWS:
            ext = Extinction(**{law: True})
            for ebv in (0.1, 0.7, 1.9):
                out["curve/%s/%s/%g" % (gname, law, ebv)] = np.array(ext.extinction(spectrum=sp, params=[ebv]),
                                                                     dtype=np.float64)
    # models ending with an Extinction component
    for name, comps, law, wl in (("full_mw", ["PL", "FE", "HOST", "BC"], "MW", mg.grid(2048)),
                                 ("nc_fe_calzetti", ["PL", "FE"], "Calzetti", mg.grid(1024)),
                                 ("nc_bc_agn", ["PL", "BC"], "AGN", mg.grid(1536)),
                                 ("nc_smc", ["PL"], "SMC", mg.grid(512)),
                                 ("nc_hg_lmc", ["PL", "HOST"], "LMC", mg.grid(1024))):
        d, e = mg.synth(wl, comps)
        ebv_truth = 0.3
        e0 = Extinction(**{law: True})
        curve = np.array(e0.extinction(spectrum=ref.Spectrum(wl, wl, wl), params=[ebv_truth]))
        d, e = d * curve, e * curve
        m = ref.Model()
        for c in comps:
            m.components.append(mg.make_component(c))
        m.components.append(Extinction(**{law: True}))
        m.data_spectrum = ref.Spectrum(wl, d, e)
        params = mg.draw_walkers(m, 14)
        lo, hi = mg.bounds_of(m)
        lo = np.append(lo, m.components[-1].EBV_min)
        hi = np.append(hi, m.components[-1].EBV_max)
        rng = np.random.RandomState(77)
        truth = np.concatenate([mg.TRUTH[c] for c in comps] + [[ebv_truth]])
        near = truth[None, :] * (1 + 0.01 * rng.standard_normal((6, len(truth))))
        near = np.clip(near, lo + 1e-3 * (hi - lo), hi - 1e-3 * (hi - lo))
        bad = near[:2].copy()
        bad[0, -1] = 2.5            # E(B-V) outside (0, 2)
        bad[1, -1] = 0.0            # on the bound: strict inequality
        params = np.concatenate([params, near, truth[None, :], bad], axis=0)
        lnp = np.array([ref.ln_posterior(p, m) for p in params], dtype=np.float64)
        mf = np.array([np.array(m.model_flux(p)) for p in params[:4]], dtype=np.float64)
        for k, v in dict(wl=wl, flux=d, err=e, params=params, lnpost=lnp, lo=lo, hi=hi, model_flux=mf,
                         comps=np.array(comps), law=np.array(law)).items():
            out["model/%s/%s" % (name, k)] = v
        print("ext case %-16s W=%3d P=%5d finite=%d" % (name, len(params), len(wl), np.isfinite(lnp).sum()))
    np.savez_compressed(os.path.join(HERE, "extinction.npz"), **out)
    print("extinction.npz", os.path.getsize(os.path.join(HERE, "extinction.npz")))


if __name__ == "__main__":
    with refshim.chdir_examples():
        make()
