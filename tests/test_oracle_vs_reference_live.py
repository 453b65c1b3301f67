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


def test_bridge_builds_equivalent_model_from_reference_object():
    """a spamm_b200 Model made from the reference's own Model object
    has the same parameter layout, prior boxes, templates and ln-lambda grids (CPU-side check)."""
    warnings.filterwarnings("ignore")
    from spamm_b200.bridge import model_from_reference
    ref = refshim.load()
    wl = grid(1024)
    with refshim.chdir_examples():
        m = ref.Model()
        m.components.append(ref.NuclearContinuumComponent(broken=False))
        m.components.append(ref.FeComponent())
        m.components.append(ref.HostGalaxyComponent())
        m.components.append(ref.BalmerCombined(BalmerContinuum=True, BalmerPseudocContinuum=True))
        m.data_spectrum = ref.Spectrum(wl, 1e-15 * (1 + wl / 5e3), 1e-16 * np.ones_like(wl))
        np.random.seed(1)
        for c in m.components:
            c.initial_values(m.data_spectrum)               # resolves the string prior bounds
    g = model_from_reference(m)
    assert g.model_parameter_names() == m.model_parameter_names()
    for gc, rc in zip(g.components, m.components):
        p = np.array(rc.initial_values(m.data_spectrum))
        assert gc.ln_priors(p) == rc.ln_priors(p)
        lo, hi = gc.prior_bounds()
        assert all(not isinstance(v, str) for v in lo + hi)
    for a, b in zip(g.components[1].log_fe, m.components[1].log_fe):
        assert np.array_equal(a.spectral_axis, b.spectral_axis) and np.array_equal(a.flux, b.flux)
    for a, b in zip(g.components[2].log_host, m.components[2].log_host):
        assert np.array_equal(a.spectral_axis, b.spectral_axis) and np.array_equal(a.flux, b.flux)
