"""Pin the numpy oracle (oracle/spamm_numpy.py) against the committed golden
vectors produced by the unmodified reference (tests/golden/make_golden.py) and
against the reference's only shipped known-answer artefact
(examples/model.pickle.gz).  CPU only."""
import os

import numpy as np
import pytest

from oracle.spamm_numpy import OracleModel, Templates, sh95_bilinear, stretch_move_emcee
from helpers import GOLDEN, Case, lnpost_cases, rel_err, grid, TRUTH

TPL = Templates()


def test_kat_pickle_lnprob():
    """All sampled (params, lnprob) pairs stored by the reference's own emcee run."""
    z = np.load(os.path.join(GOLDEN, "kat_pickle.npz"))
    o = OracleModel(z["wl"], z["flux"], z["err"], comps=["PL"], templates=TPL)
    assert o.norm_wl == float(z["norm_wavelength"]) == 1581.0
    o.set_bounds(z["lo"], z["hi"])
    got = o.ln_posterior_batch(z["params"])
    assert len(got) > 4000
    assert rel_err(got, z["lnprob"]).max() < 5e-13
    # the two examples quoted in SURVEY.md App. C1
    v = o.ln_posterior([2.2265754358975538e-17, 1.412003010341106])
    assert abs(v - (-252266.26726928068)) < 1e-7
    v = o.ln_posterior([1.7987789432468136e-17, 2.2016429312441748])
    assert abs(v - 33766.71156171418) < 1e-8


def test_component_flux_golden():
    z = np.load(os.path.join(GOLDEN, "components.npz"))
    tags = sorted({tuple(k.split("/")[:2]) for k in z.files if k.count("/") == 2})
    assert len(tags) >= 20
    for gname, tag in tags:
        wl = z["%s/wl" % gname]
        params, flux = z["%s/%s/params" % (gname, tag)], z["%s/%s/flux" % (gname, tag)]
        c = tag.split("_")[0]
        kw = dict(bc=True, bpc=True)
        if "bcFalse" in tag:
            kw["bc"] = False
        if "bpcFalse" in tag:
            kw["bpc"] = False
        if "gaussian" in tag:
            kw["line_type"] = "gaussian"
        if "brokenTrue" in tag:
            kw["broken_pl"] = True
        o = OracleModel(wl, wl, wl, comps=[c], templates=TPL, **kw)
        for p, f in zip(params, flux):
            got = o.component_flux(c, p)
            scale = np.max(np.abs(f))
            assert np.max(np.abs(got - f)) <= 2e-14 * scale, (gname, tag)


@pytest.mark.parametrize("name", lnpost_cases()[1])
def test_lnpost_golden(name):
    z, _ = lnpost_cases()
    c = Case(z, name)
    o = OracleModel(c.wl, c.flux, c.err, comps=c.comps, bc=c.bc, bpc=c.bpc, templates=TPL)
    lo, hi = o.bounds()
    assert np.array_equal(lo, c.lo) and np.allclose(hi, c.hi, rtol=1e-15, atol=0)
    # limit the heavy full-model cases so the CPU suite stays fast
    n = len(c.params) if len(c.wl) < 8000 else min(len(c.params), 81)
    got = o.ln_posterior_batch(c.params[:n])
    assert rel_err(got, c.lnpost[:n]).max() < 1e-12, name
    if name == "q6_fe_nan":
        assert np.all(got == 0.0)


def test_walker_draws_match_reference_rng_order():
    """initial_values follows Model.run_mcmc's RNG call order, so seed 42
    reproduces the reference's walkers bit for bit."""
    z, _ = lnpost_cases()
    c = Case(z, "cfg5_full")
    o = OracleModel(c.wl, c.flux, c.err, comps=c.comps, templates=TPL)
    np.random.seed(42)
    w = np.array([o.initial_values() for _ in range(40)])
    assert np.array_equal(w, c.params[:40])


def test_survey_probe_values():
    """SURVEY.md Appendix C3 numbers (reference under shims, 18 000 px grid)."""
    wl = np.arange(1000, 10000, 0.5)
    o = OracleModel(wl, wl, wl, templates=TPL)
    f = o.component_flux("PL", TRUTH["PL"])
    assert abs(f.sum() - 3.6869106070220087e-10) < 1e-22
    f = o.component_flux("FE", TRUTH["FE"])
    assert abs(f.sum() - 8.237495159855799e-11) < 1e-23
    f = o.component_flux("HOST", TRUTH["HOST"])
    assert abs(f.sum() - 2.2356813788608354e-10) < 1e-22
    f = o.component_flux("BC", TRUTH["BC"])
    assert abs(f.sum() - 4.0471641186584273e-10) < 1e-22
    assert abs(f[5292] - 3e-14) < 1e-28


def test_sh95_table_clamps_and_is_linear():
    ne, T, zt = TPL.sh95_ne, TPL.sh95_T, TPL.sh95_z
    assert zt.shape == (48, 7, 10)
    v = sh95_bilinear(zt, ne, T, 1e10, 10000.)
    assert v[0] == 1.25e-28 and v[47] == 4.02e-25      # probe values from the pickle
    lo = sh95_bilinear(zt, ne, T, 1e2, 50250.)
    assert np.array_equal(lo, zt[:, 0, -1])             # clamped on both axes
    mid = sh95_bilinear(zt, ne, T, 5.5e8, 750.)
    exp = 0.25 * (zt[:, 0, 0] + zt[:, 1, 0] + zt[:, 0, 1] + zt[:, 1, 1])
    assert np.allclose(mid, exp, rtol=1e-14)


def test_stretch_move_recovers_gaussian():
    """emcee 2.2.1 stretch-move restatement samples a known Gaussian."""
    rng = np.random.RandomState(3)
    mu, sig = np.array([1.0, -2.0]), np.array([0.5, 2.0])
    f = lambda q: -0.5 * np.sum(((q - mu) / sig) ** 2, axis=1)  # noqa: E731
    p0 = mu + sig * rng.standard_normal((40, 2))
    chain, lnp, acc = stretch_move_emcee(f, p0, 600, rng=rng)
    s = chain[:, 100:, :].reshape(-1, 2)
    assert np.all(np.abs(s.mean(0) - mu) < 0.1 * sig)
    assert np.all(np.abs(s.std(0) / sig - 1) < 0.1)
    assert 0.5 < acc.mean() < 0.9


def test_extinction_oracle_vs_reference_fixtures():
    """Extinction component (ReddeningLaw.py): the five curves and models that end with an Extinction
    against fixtures made by the unmodified reference (tests/golden/make_golden_ext.py)."""
    from oracle.spamm_numpy import OracleModel, extinction_k
    z = np.load(os.path.join(GOLDEN, "extinction.npz"))
    n_curves = 0
    for key in z.files:
        if key.startswith("curve/") and key.count("/") == 3:
            _, g, law, ebv = key.split("/")
            o = OracleModel(z["curve/%s/wl" % g], z["curve/%s/wl" % g], z["curve/%s/wl" % g],
                            comps=["EXT"], ext_law=law)
            assert np.array_equal(o.ext_k, extinction_k(z["curve/%s/wl" % g], law))
            got = o.extinction([float(ebv)])
            assert rel_err(got, z[key]).max() < 1e-14, key
            n_curves += 1
    assert n_curves == 45
    names = sorted({k.split("/")[1] for k in z.files if k.startswith("model/")})
    assert len(names) == 5
    for n in names:
        g = lambda k: z["model/%s/%s" % (n, k)]  # noqa: E731
        o = OracleModel(g("wl"), g("flux"), g("err"), comps=[str(c) for c in g("comps")] + ["EXT"],
                        ext_law=str(g("law")))
        lo, hi = o.bounds()
        assert np.allclose(lo, g("lo"), rtol=1e-15, atol=0) and np.allclose(hi, g("hi"), rtol=1e-15, atol=0)
        got = o.ln_posterior_batch(g("params"))
        assert rel_err(got, g("lnpost")).max() < 1e-12, n
        mf = np.array([o.model_flux(p) for p in g("params")[:4]])
        assert np.max(np.abs(mf - g("model_flux"))) <= 2e-14 * np.max(np.abs(g("model_flux")))
