"""The device status word must surface as an error: a walker that needs a table entry the plan does not
hold (a template width outside the broadening bank, a log_conv kernel wider than one pixel, a Balmer edge
beyond the staged continuum samples) gets NaN / garbage, which `nan_to_num` would turn into ln L = 0.0 --
far ABOVE every legitimate posterior value.  A sampler must never see that as a number."""
import numpy as np
import pytest

from helpers import Case, lnpost_cases

pytestmark = pytest.mark.gpu


def _plan_with_bounds(widen):
    """cfg5 model, prior box widened per parameter name -> (plan, params inside the reference box)."""
    from spamm_b200.plan import ModelPlan
    from spamm_b200.synth import build_model
    z, _ = lnpost_cases()
    c = Case(z, "cfg5_full")
    m = build_model(list(c.comps), c.wl, c.flux, c.err, max_walkers=64)
    names = m.model_parameter_names()
    base = m.plan
    lo, hi = np.array(base.lo), np.array(base.hi)
    for name, (l, h) in widen.items():
        i = names.index(name)
        lo[i], hi[i] = l, h
    plan = ModelPlan(m.components, m.data_spectrum, max_walkers=64, bounds=(lo, hi))
    good = c.params[np.isfinite(c.lnpost)][:8].copy()
    return m, plan, names, good


CASES = [
    ("fe_width", (100.0, 10000.0), 500.0, "broadening bank"),          # sqrt(500^2 - 900^2) = NaN
    ("bc_lwidth", (100.0, 1e9), 5e8, "log_conv"),                      # round(5 dpix) != 0
    ("bc_loffset", (-1e11, 10.0), -1e10, "staged"),                    # edge pixel far redward of the stage
]


@pytest.mark.parametrize("name,box,value,word", CASES)
def test_status_word_raises_from_every_entry_point(name, box, value, word):
    import torch
    from spamm_b200 import _lib
    m, plan, names, good = _plan_with_bounds({name: box})
    want = plan.lnposterior_host(good)                         # clean walkers: no error
    assert np.all(np.isfinite(want))
    bad = good.copy()
    bad[3, names.index(name)] = value
    # host-buffer entry point: checked at its own synchronisation
    with pytest.raises(_lib.SpammError, match=word):
        plan.lnposterior_host(bad)
    assert plan.status() == 0                                   # reported once, then cleared
    assert np.array_equal(plan.lnposterior_host(good), want)
    # device-buffer entry point is asynchronous: the error surfaces at check() or at the next call
    d = torch.from_numpy(bad).cuda()
    plan.lnposterior(d)
    with pytest.raises(_lib.SpammError, match=word):
        plan.check()
    plan.lnposterior(d)
    torch.cuda.synchronize()
    with pytest.raises(_lib.SpammError, match=word):
        plan.lnposterior(torch.from_numpy(good).cuda())
    torch.cuda.synchronize()
    plan.check()


def test_sampler_run_raises_on_status():
    from spamm_b200 import _lib
    from spamm_b200.sampler import EnsembleSampler
    m, plan, names, good = _plan_with_bounds({"fe_width": (100.0, 10000.0)})
    rng = np.random.RandomState(3)
    pos = good[0] * (1 + 1e-4 * rng.randn(64, plan.n_dim))
    pos[5, names.index("fe_width")] = 500.0
    s = EnsembleSampler(nwalkers=64, dim=plan.n_dim, plan=plan, seed=1)
    with pytest.raises(_lib.SpammError, match="broadening bank"):
        s.run_mcmc(pos, 3)


def test_big_plan_survives_a_later_small_plan():
    """The dynamic shared-memory limit of the fused kernel is a per-function attribute: building a small
    plan (a single component's flux()) must not lower it under a live, larger plan."""
    from spamm_b200.synth import build_model
    z, _ = lnpost_cases()
    c = Case(z, "cfg5_full")
    big = build_model(list(c.comps), c.wl, c.flux, c.err, max_walkers=64)
    want = big.lnposterior_batch(c.params)
    nc = big.components[0]
    nc.flux(big.data_spectrum, c.params[0, :2])               # builds an NC-only plan: a few hundred bytes of smem
    small = build_model(["PL"], c.wl[:256], c.flux[:256], c.err[:256], max_walkers=8)
    small.lnposterior_batch(c.params[:4, :2])
    assert np.array_equal(big.lnposterior_batch(c.params), want)


def test_model_pickles_after_component_flux():
    import pickle
    from spamm_b200.synth import build_model
    z, _ = lnpost_cases()
    c = Case(z, "cfg2_nc_fe")
    m = build_model(list(c.comps), c.wl, c.flux, c.err, max_walkers=64)
    want = m.lnposterior_batch(c.params[:4])
    for comp in m.components:
        k = comp.parameter_count
        off = m.plan.offsets[m.components.index(comp)]
        comp.flux(m.data_spectrum, c.params[0, off:off + k])  # caches a ctypes-backed plan on the component
    m2 = pickle.loads(pickle.dumps(m))
    assert np.array_equal(m2.lnposterior_batch(c.params[:4]), want)
