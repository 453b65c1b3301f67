"""`rebin_spec: True` on the CUDA path (parity unpinned -- see tests/test_rebin_spec.py): every np.interp
resampling of the hot path becomes the flux-conserving rebin (template -> ln-lambda grid at initialise,
convolved row -> data grid per walker, both resamplings of the Balmer continuum's log_conv).  Checked against
the numpy oracle running the extended-precision literal restatement of the rebin."""
import numpy as np
import pytest

from helpers import TRUTH, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture
def rebin_on(monkeypatch):
    from spamm_b200.components import ComponentBase
    monkeypatch.setitem(ComponentBase.PARS, "rebin_spec", True)


@pytest.mark.parametrize("comps", [["PL", "FE"], ["PL", "HOST"], ["PL", "BC"], ["PL", "FE", "HOST", "BC"]])
def test_rebin_spec_true_matches_the_oracle(rebin_on, comps):
    from oracle.spamm_numpy import OracleModel
    from spamm_b200.synth import build_model
    P = 700
    wl = 1500.0 + np.arange(P) * (4500.0 / P)
    o0 = OracleModel(wl, wl, wl, comps=comps, rebin_spec=True)
    fl = [o0.component_flux(c, TRUTH[c]) for c in comps]
    d = np.sum(fl, axis=0)
    e = np.sqrt(np.sum([(0.05 * f) ** 2 for f in fl], axis=0))
    o = OracleModel(wl, d, e, comps=comps, rebin_spec=True)
    np.random.seed(7)
    params = np.array([o.initial_values() for _ in range(5)])
    params[0] = np.concatenate([TRUTH[c] for c in comps])
    want = np.array([o.ln_posterior(p) for p in params])
    m = build_model(comps, wl, d, e, max_walkers=8)
    assert not m.components[-1].fast_interp
    got = m.lnposterior_batch(params)
    # (the composed log_conv operator runs on the generic instantiation; rebinned template rows are bank rows
    #  like any other and keep the specialised one)
    assert m.plan.specialised == ("BC" not in comps)
    assert rel_err(got, want).max() < 1e-10, (comps, got, want)
    assert m.plan.status() == 0
    # component by component, against the oracle's fluxes (<= 1e-12 of the peak)
    off = 0
    for c, comp in zip(comps, m.components):
        k = comp.parameter_count
        f_gpu = comp.flux(m.data_spectrum, params[1, off:off + k])
        f_ora = o.component_flux(c, params[1, off:off + k])
        assert np.max(np.abs(f_gpu - f_ora)) <= 1e-12 * np.max(np.abs(f_ora)), c
        off += k
    # and it is a different model from the default np.interp one (the branch is really taken)
    from spamm_b200.components import ComponentBase
    ComponentBase.PARS["rebin_spec"] = False
    m0 = build_model(comps, wl, d, e, max_walkers=8)
    plain = m0.lnposterior_batch(params)              # (the plan is built on first use: keep the flag off until here)
    ComponentBase.PARS["rebin_spec"] = True
    assert not np.array_equal(plain, got)
