"""Live cross-check of the numpy oracle against the UNMODIFIED reference
imported under oracle/refshim.py.  Only runs where /root/reference exists (the
build container); skipped on the GPU box."""
import warnings

import numpy as np
import pytest

from oracle import refshim
from oracle.spamm_numpy import OracleModel
from helpers import grid, TRUTH

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree absent")


def test_full_model_lnposterior_matches_reference():
    warnings.filterwarnings("ignore")
    ref = refshim.load()
    wl = grid(2048)
    comps = ["PL", "FE", "HOST", "BC"]
    o0 = OracleModel(wl, wl, wl, comps=comps)
    fl = [o0.component_flux(c, TRUTH[c]) for c in comps]
    d = np.sum(fl, axis=0)
    e = np.sqrt(np.sum([(0.05 * f) ** 2 for f in fl], axis=0))
    with refshim.chdir_examples():
        m = ref.Model()
        m.components.append(ref.NuclearContinuumComponent(broken=False))
        m.components.append(ref.FeComponent())
        m.components.append(ref.HostGalaxyComponent())
        m.components.append(ref.BalmerCombined(BalmerContinuum=True, BalmerPseudocContinuum=True))
        m.data_spectrum = ref.Spectrum(wl, d, e)
        np.random.seed(5)
        ws = []
        for _ in range(6):
            p = []
            for c in m.components:
                p += list(c.initial_values(m.data_spectrum))
            ws.append(p)
        want = np.array([ref.ln_posterior(np.array(p), m) for p in ws])
    o = OracleModel(wl, d, e, comps=comps)
    np.random.seed(5)
    ws2 = np.array([o.initial_values() for _ in range(6)])
    assert np.array_equal(np.array(ws), ws2)
    got = o.ln_posterior_batch(ws2)
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-13
