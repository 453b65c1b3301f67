"""GPU parity tests proper: the CUDA path (through the C ABI, via the
reference-shaped Python plugin layer) against the golden vectors produced by
the unmodified reference and against the numpy oracle on fresh seeded inputs.

Tolerance: BASELINE.json north_star -- ln-posterior relative 1e-10 in FP64.
Component fluxes are held to 1e-12 of the component's peak."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, Case, lnpost_cases, rel_err, grid, TRUTH

pytestmark = pytest.mark.gpu

LNPOST_RTOL = 1e-10
FLUX_RTOL = 1e-12


def _components(comps, bc=True, bpc=True, broken=False, line_type=None):
    from spamm_b200.components.NuclearContinuumComponent import NuclearContinuumComponent
    from spamm_b200.components.FeComponent import FeComponent
    from spamm_b200.components.HostGalaxyComponent import HostGalaxyComponent
    from spamm_b200.components.BalmerContinuumCombined import BalmerCombined
    from spamm_b200.utils.parse_pars import parse_pars
    out = []
    for c in comps:
        if c == "PL":
            out.append(NuclearContinuumComponent(broken=broken))
        elif c == "FE":
            out.append(FeComponent())
        elif c == "HOST":
            out.append(HostGalaxyComponent())
        elif c == "BC":
            pars = parse_pars()["balmer_continuum"]
            if line_type:
                pars["bc_line_type"] = line_type
            out.append(BalmerCombined(pars=pars, BalmerContinuum=bc, BalmerPseudocContinuum=bpc))
    return out


def _model(case, template_mode=None):
    from spamm_b200.Model import Model
    from spamm_b200.Spectrum import Spectrum
    m = Model()
    if template_mode is not None:
        m.template_mode = template_mode
    m.components += _components(case.comps, bc=case.bc, bpc=case.bpc)
    m.data_spectrum = Spectrum(case.wl, case.flux, case.err)
    return m


def test_component_flux_vs_reference_golden():
    from spamm_b200.Spectrum import Spectrum
    z = np.load(os.path.join(GOLDEN, "components.npz"))
    tags = sorted({tuple(k.split("/")[:2]) for k in z.files if k.count("/") == 2})
    worst = {}
    for gname, tag in tags:
        wl = z["%s/wl" % gname]
        params, flux = z["%s/%s/params" % (gname, tag)], z["%s/%s/flux" % (gname, tag)]
        c = tag.split("_")[0]
        comp = _components([c], bc="bcFalse" not in tag, bpc="bpcFalse" not in tag,
                           broken="brokenTrue" in tag,
                           line_type="gaussian" if "gaussian" in tag else None)[0]
        sp = Spectrum(wl, wl, wl)
        comp.initialize(sp)
        got = comp.flux_batch(sp, params)
        assert got.shape == flux.shape
        err = np.max(np.abs(got - flux), axis=1) / np.max(np.abs(flux), axis=1)
        worst[(gname, tag)] = err.max()
        assert err.max() < FLUX_RTOL, (gname, tag, err)
        # single-walker plugin call == row of the batched call (Component.flux API)
        one = comp.flux(sp, list(params[0]))
        assert np.array_equal(one, got[0])
    print("worst component flux errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:5])


@pytest.mark.parametrize("name", lnpost_cases()[1])
def test_lnposterior_vs_reference_golden(name):
    z, _ = lnpost_cases()
    c = Case(z, name)
    m = _model(c)
    got = m.lnposterior_batch(c.params)
    # the plan resolved the same prior boxes as the reference
    assert np.array_equal(m.plan.lo, c.lo)
    assert np.allclose(m.plan.hi, c.hi, rtol=1e-15, atol=0)
    err = rel_err(got, c.lnpost)
    assert err.max() < LNPOST_RTOL, (name, err.max())
    assert np.array_equal(np.isneginf(got), np.isneginf(c.lnpost))
    if name == "q6_fe_nan":
        assert np.all(got == 0.0)            # NaN flux -> lnL == 0.0 (Model.py:443)
    assert m.plan.status() == 0
    # ln_posterior(new_params, model): the per-walker entry point emcee calls
    from spamm_b200.Model import ln_posterior
    assert ln_posterior(c.params[0], m) == got[0]


def test_known_answer_pickle():
    """The reference's own stored emcee chain + lnprob (examples/model.pickle.gz)."""
    from spamm_b200.Model import Model
    from spamm_b200.Spectrum import Spectrum
    z = np.load(os.path.join(GOLDEN, "kat_pickle.npz"))
    m = Model()
    nc = _components(["PL"])[0]
    nc.norm_min, nc.norm_max = float(z["lo"][0]), float(z["hi"][0])
    nc.slope_min, nc.slope_max = float(z["lo"][1]), float(z["hi"][1])
    m.components.append(nc)
    m.data_spectrum = Spectrum(z["wl"], z["flux"], z["err"])
    got = m.lnposterior_batch(z["params"])
    assert rel_err(got, z["lnprob"]).max() < LNPOST_RTOL


def test_lnposterior_vs_oracle_fresh_inputs():
    """Seeded inputs the fixtures never saw, all four components, three grids."""
    from oracle.spamm_numpy import OracleModel
    rng = np.random.RandomState(99)
    for P, W in ((1024, 48), (3000, 24)):
        wl = np.sort(1000.0 + 9000.0 * rng.rand(P)) if P == 3000 else grid(P)
        comps = ["PL", "FE", "HOST", "BC"]
        o0 = OracleModel(wl, wl, wl, comps=comps)
        fl = [o0.component_flux(c, TRUTH[c]) for c in comps]
        d = np.sum(fl, axis=0) * (1 + 0.02 * rng.standard_normal(P))
        e = np.sqrt(np.sum([(0.05 * f) ** 2 for f in fl], axis=0))
        o = OracleModel(wl, d, e, comps=comps)
        np.random.seed(7 + P)
        params = np.array([o.initial_values() for _ in range(W)])
        params[::5, 17] = rng.uniform(-10, 10, len(params[::5]))
        want = o.ln_posterior_batch(params)
        c = Case.__new__(Case)
        c.wl, c.flux, c.err, c.comps, c.bc, c.bpc = wl, d, e, comps, True, True
        m = _model(c)
        got = m.lnposterior_batch(params)
        assert rel_err(got, want).max() < LNPOST_RTOL


def test_bank_equals_direct_broadening(monkeypatch):
    """The exact sigma bank and the per-walker broadening kernel give the same bits (compared through
    the generic instantiation of the fused kernel; the specialised one agrees to a few ulp)."""
    from spamm_b200 import _lib
    z, _ = lnpost_cases()
    c = Case(z, "cfg5_full")
    fast = _model(c, _lib.TEMPLATES_BANK).lnposterior_batch(c.params)
    monkeypatch.setenv("SPAMM_NO_FULL_KERNEL", "1")
    a = _model(c, _lib.TEMPLATES_BANK)
    b = _model(c, _lib.TEMPLATES_DIRECT)
    assert a.plan.template_mode == _lib.TEMPLATES_BANK and a.plan.bank_rows > 100
    assert b.plan.template_mode == _lib.TEMPLATES_DIRECT
    ga, gb = a.lnposterior_batch(c.params), b.lnposterior_batch(c.params)
    assert np.array_equal(ga, gb)
    fa, fb = a.model_flux_batch(c.params[:8]), b.model_flux_batch(c.params[:8])
    assert np.array_equal(fa, fb)
    assert rel_err(fast, ga).max() < 1e-13


def test_flux_outside_prior_uses_direct_kernel():
    """component.flux() with a width no prior admits (bank miss) still matches the oracle."""
    from oracle.spamm_numpy import OracleModel
    from spamm_b200.Spectrum import Spectrum
    wl = grid(2048)
    o = OracleModel(wl, wl, wl, comps=["FE"])
    p = np.array([[1e-14, 2e-14, 3e-14, 14000.0], [1e-14, 2e-14, 3e-14, 2500.0]])
    want = np.array([o.component_flux("FE", q) for q in p])
    fe = _components(["FE"])[0]
    sp = Spectrum(wl, wl, wl)
    fe.initialize(sp)
    got = fe.flux_batch(sp, p)
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < FLUX_RTOL


def test_model_api_pieces_agree():
    """prior + likelihood(model_flux) == ln_posterior, through the separate kernels."""
    z, _ = lnpost_cases()
    c = Case(z, "full_fp6")
    m = _model(c)
    got = m.lnposterior_batch(c.params)
    for i in (0, 3, len(c.params) - 1):
        lp = m.prior(c.params[i])
        assert (lp == 0) == np.isfinite(c.lnpost[i])
        if lp == 0:
            f = m.model_flux(c.params[i])
            v = m.likelihood(f) + lp
            assert abs(v - got[i]) <= 1e-12 * abs(got[i])
            # the step-wise form of model_flux (Model.add_component, Model.py:368-380)
            m.model_spectrum.flux = np.zeros(len(c.wl))
            p = np.copy(c.params[i])
            for comp in m.components:
                m.add_component(component=comp, parameters=p[0:comp.parameter_count])
                p = p[comp.parameter_count:]
            assert np.max(np.abs(m.model_spectrum.flux - f)) <= 1e-13 * np.max(np.abs(f))


def test_full_size_properties():
    """BASELINE config 5 size (8192 px; walkers reduced only where a [W,P] array is
    materialised): determinism, batch-independence, permutation equivariance, and
    exact linearity of the model flux in the normalisations."""
    import torch
    z, _ = lnpost_cases()
    c = Case(z, "cfg5_full")
    m = _model(c)
    rng = np.random.RandomState(5)
    lo, hi = m.plan.lo, m.plan.hi
    W = 8192
    params = lo + (hi - lo) * rng.rand(W, len(lo))
    params[:81] = c.params                           # includes out-of-prior rows
    a = m.lnposterior_batch(params)
    b = m.lnposterior_batch(params)
    assert np.array_equal(a, b)
    assert rel_err(a[:81], c.lnpost).max() < LNPOST_RTOL
    perm = rng.permutation(W)
    assert np.array_equal(m.lnposterior_batch(params[perm]), a[perm])
    assert np.array_equal(m.lnposterior_batch(params[100:357]), a[100:357])
    assert np.all(np.isfinite(a[81:])) and np.all(a[81:] < 0)
    # device-resident entry point gives the same bits as the host-buffer one
    d = torch.from_numpy(params).cuda()
    assert np.array_equal(m.plan.lnposterior(d).cpu().numpy(), a)
    # doubling every normalisation doubles the model flux exactly
    sub = params[81:81 + 64].copy()
    f1 = m.model_flux_batch(sub)
    norm_idx = [0, 2, 3, 4, 6, 7, 8, 9, 10, 11, 12, 14]
    sub2 = sub.copy()
    sub2[:, norm_idx] *= 2.0
    f2 = m.model_flux_batch(sub2)
    assert np.array_equal(f2, 2.0 * f1)
    assert m.plan.status() == 0


def test_empty_and_ragged_batches():
    z, _ = lnpost_cases()
    c = Case(z, "cfg2_nc_fe")
    m = _model(c)
    assert m.lnposterior_batch(np.zeros((0, len(c.lo)))).shape == (0,)
    one = m.lnposterior_batch(c.params[:1])
    assert rel_err(one, c.lnpost[:1]).max() < LNPOST_RTOL
    with pytest.raises(ValueError):
        m.lnposterior_batch(np.zeros((3, len(c.lo) + 1)))
    # more walkers than the plan's staging capacity: chunked inside the library
    m2 = _model(c)
    m2.max_walkers = 16
    big = np.tile(c.params, (3, 1))
    assert np.array_equal(m2.lnposterior_batch(big), np.tile(m.lnposterior_batch(c.params), 3))


def test_cluster_series_matches_direct_line_sum():
    """Far-field series for distant line clusters vs. every line summed directly
    (makelines, BC:315-333): pixel-wise relative agreement at machine precision over the
    whole prior range of widths/offsets, on the BASELINE grid and on a ragged one."""
    from spamm_b200.components.BalmerContinuumCombined import BalmerCombined
    from spamm_b200.Spectrum import Spectrum
    from spamm_b200.utils.parse_pars import parse_pars
    rng = np.random.RandomState(17)
    for wl in (grid(8192), np.sort(3000.0 + 5000.0 * rng.rand(2500))):
        sp = Spectrum(wl, wl, wl)
        pe = parse_pars()["balmer_continuum"]
        pe["bc_exact_line_sum"] = True
        exact = BalmerCombined(pars=pe, BalmerContinuum=False, BalmerPseudocContinuum=True)
        fast = BalmerCombined(BalmerContinuum=False, BalmerPseudocContinuum=True)
        assert exact.exact_line_sum and not fast.exact_line_sum
        W = 96
        p = np.column_stack([rng.uniform(1e-15, 7e-14, W), rng.uniform(500, 1e5, W), rng.uniform(0, 2, W),
                             rng.uniform(-10, 10, W), rng.uniform(100, 1e4, W), rng.uniform(2, 9, W)])
        p[:8, 4] = [100.0001, 100.5, 150., 250., 400., 9999., 120., 101.]     # narrowest / broadest lines
        p[:8, 3] = [-9.999, 9.999, 0., -5., 5., 0., 9.9, -9.9]
        fe = exact.flux_batch(sp, p)
        ff = fast.flux_batch(sp, p)
        nz = fe > 0
        assert np.array_equal(nz, ff > 0)
        rel = np.abs(ff[nz] - fe[nz]) / fe[nz]
        assert rel.max() < 2e-14, rel.max()


def test_lnposterior_exact_line_sum_mode_agrees():
    z, _ = lnpost_cases()
    c = Case(z, "cfg5_full")
    m = _model(c)
    m2 = _model(c)
    m2.components[-1].exact_line_sum = True
    m2.invalidate_plan()
    a, b = m.lnposterior_batch(c.params), m2.lnposterior_batch(c.params)
    assert rel_err(a, b).max() < 1e-12
    assert rel_err(b, c.lnpost).max() < LNPOST_RTOL


@pytest.mark.parametrize("name,W", [("cfg2_nc_fe", 1024), ("cfg3_nc_hg", 2048), ("cfg4_nc_bc", 4096)])
def test_baseline_configs_at_full_walker_counts(name, W):
    """BASELINE.json configs 2-4 at their quoted walker counts: a slice against the oracle,
    everything else through size-independent properties."""
    from oracle.spamm_numpy import OracleModel
    z, _ = lnpost_cases()
    c = Case(z, name)
    m = _model(c)
    m.max_walkers = W
    rng = np.random.RandomState(len(name))
    lo, hi = c.lo, c.hi
    params = lo + (hi - lo) * rng.rand(W, len(lo))
    got = m.lnposterior_batch(params)
    assert got.shape == (W,) and np.all(np.isfinite(got))
    assert np.array_equal(got, m.lnposterior_batch(params))
    o = OracleModel(c.wl, c.flux, c.err, comps=c.comps, bc=c.bc, bpc=c.bpc)
    pick = rng.choice(W, 6, replace=False)
    want = o.ln_posterior_batch(params[pick])
    assert rel_err(got[pick], want).max() < LNPOST_RTOL
    # a walker moved just outside any one bound is rejected without touching the rest
    D_CHECK = min(len(lo), 8)
    bad = params[:D_CHECK].copy()
    for j in range(D_CHECK):
        bad[j, j] = hi[j]
    res = m.lnposterior_batch(np.concatenate([bad, params[D_CHECK:64]]))
    assert np.all(np.isneginf(res[:D_CHECK])) and np.array_equal(res[D_CHECK:], got[D_CHECK:64])


def test_reference_test_grid_18000_pixels():
    """parameters.yaml testing grid: np.arange(1000, 10000, 0.5), all components."""
    from oracle.spamm_numpy import OracleModel
    wl = np.arange(1000, 10000, 0.5)
    comps = ["PL", "FE", "HOST", "BC"]
    o0 = OracleModel(wl, wl, wl, comps=comps)
    fl = [o0.component_flux(cn, TRUTH[cn]) for cn in comps]
    d = np.sum(fl, axis=0)
    e = np.sqrt(np.sum([(0.05 * f) ** 2 for f in fl], axis=0))
    o = OracleModel(wl, d, e, comps=comps)
    np.random.seed(42)
    params = np.array([o.initial_values() for _ in range(48)])
    params[0] = np.concatenate([TRUTH[cn] for cn in comps])
    c = Case.__new__(Case)
    c.wl, c.flux, c.err, c.comps, c.bc, c.bpc = wl, d, e, comps, True, True
    m = _model(c)
    got = m.lnposterior_batch(params)
    want = o.ln_posterior_batch(params[:5])
    assert rel_err(got[:5], want).max() < LNPOST_RTOL
    assert abs(got[0] - 6.087254312300e+05) < 1e-5 * 6.09e5      # SURVEY.md App. C3 probe value
    assert m.plan.status() == 0


def test_tiny_and_odd_sized_spectra():
    """P not a multiple of the 64-pixel warp task, P smaller than one task, spectra entirely
    blueward / redward of the Balmer edge."""
    from oracle.spamm_numpy import OracleModel
    rng = np.random.RandomState(3)
    for wl in (np.linspace(3000., 4500., 37), np.linspace(1200., 3300., 333), np.linspace(3800., 9000., 1001)):
        comps = ["PL", "FE", "HOST", "BC"]
        o0 = OracleModel(wl, wl, wl, comps=comps)
        fl = [o0.component_flux(cn, TRUTH[cn]) for cn in comps]
        d = np.sum(fl, axis=0) * (1 + 0.01 * rng.standard_normal(len(wl)))
        e = 0.05 * np.abs(d) + 1e-18
        lo = np.array([0, -3, 0, 0, 0, 901, 0, 0, 0, 0, 0, 0, 0, 30, 0, 500, 0, -10, 100, 2], dtype=float)
        hi = np.array([1e-13, 3, 1e-13, 1e-13, 1e-13, 1e4, 1e-14, 1e-14, 1e-14, 1e-14, 1e-14, 1e-14, 1e-14, 1e3,
                       1e-13, 1e5, 2, 10, 1e4, 9], dtype=float)
        params = lo + (hi - lo) * rng.rand(12, 20)
        o = OracleModel(wl, d, e, comps=comps)
        o.set_bounds(lo, hi)
        with np.errstate(all="ignore"):
            want = o.ln_posterior_batch(params)
        from spamm_b200.plan import ModelPlan
        from spamm_b200.Spectrum import Spectrum
        sp = Spectrum(wl, d, e)
        cs = _components(comps)
        for cc in cs:
            cc.initialize(sp)
        plan = ModelPlan(cs, sp, with_priors=False, bounds=(lo, hi), max_walkers=16)
        got = plan.lnposterior_host(params)
        assert rel_err(got, want).max() < LNPOST_RTOL, (len(wl), got, want)


def test_non_canonical_component_order():
    """Model.components order fixes the parameter layout and the summation order
    (Model.py:353-364); an order other than run_spamm's PL, FE, HOST, BC takes the generic kernel."""
    from oracle.spamm_numpy import OracleModel
    z, _ = lnpost_cases()
    c = Case(z, "full_fp6")
    for comps in (["HOST", "BC", "PL", "FE"], ["BC", "PL"], ["FE", "PL"]):
        o = OracleModel(c.wl, c.flux, c.err, comps=comps)
        np.random.seed(11)
        params = np.array([o.initial_values() for _ in range(10)])
        want = o.ln_posterior_batch(params)
        cc = Case.__new__(Case)
        cc.wl, cc.flux, cc.err, cc.comps, cc.bc, cc.bpc = c.wl, c.flux, c.err, comps, True, True
        m = _model(cc)
        got = m.lnposterior_batch(params)
        assert rel_err(got, want).max() < LNPOST_RTOL, comps
        f = m.model_flux_batch(params[:2])
        wantf = np.array([o.model_flux(p) for p in params[:2]])
        assert np.max(np.abs(f - wantf)) <= FLUX_RTOL * np.max(np.abs(wantf))


def test_degenerate_data_values_follow_nan_to_num():
    """Model.likelihood edge values (Model.py:440-443): a zero or NaN error / NaN flux pixel makes
    the sum NaN -> 0.0; negative errors are squared away; an overflowing chi^2 -> -DBL_MAX."""
    from oracle.spamm_numpy import OracleModel
    z, _ = lnpost_cases()
    c = Case(z, "cfg2_nc_fe")
    params = c.params[:12]
    for kind in ("neg_err", "zero_err", "nan_flux", "huge_flux"):
        d, e = c.flux.copy(), c.err.copy()
        if kind == "neg_err":
            e[::7] *= -1.0
        elif kind == "zero_err":
            e[100] = 0.0
        elif kind == "nan_flux":
            d[5] = np.nan
        else:
            d[7] = 1e160                       # ((d-m)/e)^2 overflows -> -inf -> -1.797e308
        cc = Case.__new__(Case)
        cc.wl, cc.flux, cc.err, cc.comps, cc.bc, cc.bpc = c.wl, d, e, c.comps, True, True
        m = _model(cc)
        m.components[0].norm_max, m.components[1].norm_max = float(c.hi[0]), float(c.hi[2])
        m.invalidate_plan()
        o = OracleModel(c.wl, d, e, comps=c.comps)
        o.set_bounds(c.lo, c.hi)
        with np.errstate(all="ignore"):
            want = o.ln_posterior_batch(params)
        got = m.lnposterior_batch(params)
        assert rel_err(got, want).max() < LNPOST_RTOL, (kind, got[:3], want[:3])
        if kind in ("zero_err", "nan_flux"):
            assert np.all(got[np.isfinite(want)] == 0.0)
        if kind == "huge_flux":
            assert np.all(got[np.isfinite(want)] == -1.7976931348623157e308)


def test_walker_split_over_cluster_is_bit_identical():
    """Small ensembles run one walker per thread-block cluster (2/4/8 CTAs share a walker's pixel
    tasks, chi^2 partials gathered through distributed shared memory in the fixed chunk order):
    every split must reproduce the one-CTA-per-walker result bit for bit, ln-posterior and flux,
    including out-of-prior walkers and both the canonical and the generic component order."""
    z, _ = lnpost_cases()
    for name, comps in (("full_fp6", None), ("cfg1_nc_fp1", None), ("full_fp6", ["HOST", "BC", "PL", "FE"])):
        c = Case(z, name)
        if comps is None:
            m = _model(c)
            params = c.params[:30].copy()
        else:
            from oracle.spamm_numpy import OracleModel
            o = OracleModel(c.wl, c.flux, c.err, comps=comps)
            np.random.seed(5)
            params = np.array([o.initial_values() for _ in range(30)])
            cc = Case.__new__(Case)
            cc.wl, cc.flux, cc.err, cc.comps, cc.bc, cc.bpc = c.wl, c.flux, c.err, comps, True, True
            m = _model(cc)
        params[3, -1] = 1e9                       # one walker outside its prior box -> -inf
        m.plan.set_walker_split(0)
        base = m.lnposterior_batch(params)
        basef = m.model_flux_batch(params[:3])
        assert np.isneginf(base[3]) and np.isfinite(base).sum() >= 20
        if comps is None:
            ok = np.isfinite(c.lnpost[:30])
            ok[3] = False
            assert rel_err(base[ok], c.lnpost[:30][ok]).max() < LNPOST_RTOL
        for split in (2, 3, 4, 8, 1):
            m.plan.set_walker_split(split)
            got = m.lnposterior_batch(params)
            assert np.array_equal(got, base), (name, split, np.abs(got - base).max())
            assert np.array_equal(m.model_flux_batch(params[:3]), basef), (name, split)


def test_full_model_specialisation_equals_generic_kernel(monkeypatch):
    """The headline model (NC + Fe x3 + host x7 + Balmer, Lorentzian, bank rows, prior box that bounds
    every exp argument) runs a compile-time specialised instantiation of the fused kernel without exp
    range checks and with reciprocal-multiplies in the blackbody; it must agree with the generic
    instantiation to a few ulp (and with the fixture), and be independent of the cluster split."""
    z, _ = lnpost_cases()
    c = Case(z, "cfg5_full")
    m_full = _model(c)
    got = m_full.lnposterior_batch(c.params)
    assert m_full.plan.specialised
    monkeypatch.setenv("SPAMM_NO_FULL_KERNEL", "1")
    m_gen = _model(c)
    ref = m_gen.lnposterior_batch(c.params)
    assert not m_gen.plan.specialised
    monkeypatch.delenv("SPAMM_NO_FULL_KERNEL")
    assert rel_err(got, ref).max() < 1e-13
    assert np.array_equal(np.isneginf(got), np.isneginf(ref))
    assert rel_err(got, c.lnpost).max() < LNPOST_RTOL
    for split in (2, 3, 8):
        m_full.plan.set_walker_split(split)
        assert np.array_equal(m_full.lnposterior_batch(c.params), got)


def test_extinction_component_vs_reference_golden():
    """Extinction (ReddeningLaw.py): curves of the five laws from the device, and ln-posterior /
    model flux of models ending with an Extinction, against the reference-made fixtures."""
    from spamm_b200.Model import Model
    from spamm_b200.Spectrum import Spectrum
    from spamm_b200.components.ReddeningLaw import Extinction
    z = np.load(os.path.join(GOLDEN, "extinction.npz"))
    for g in ("g4096", "wide", "test7"):
        wl = z["curve/%s/wl" % g]
        sp = Spectrum(wl, wl, wl)
        for law in ("MW", "AGN", "LMC", "SMC", "Calzetti"):
            ext = Extinction(**{law: True})
            for ebv in ("0.1", "0.7", "1.9"):
                got = ext.extinction(spectrum=sp, params=[float(ebv)])
                assert rel_err(got, z["curve/%s/%s/%s" % (g, law, ebv)]).max() < 1e-13, (g, law, ebv)
    assert np.all(Extinction().extinction(spectrum=sp, params=[0.5]) == 1.0)
    names = sorted({k.split("/")[1] for k in z.files if k.startswith("model/")})
    for n in names:
        g = lambda k: z["model/%s/%s" % (n, k)]  # noqa: E731
        m = Model()
        m.components += _components([str(c) for c in g("comps")])
        m.components.append(Extinction(**{str(g("law")): True}))
        m.data_spectrum = Spectrum(g("wl"), g("flux"), g("err"))
        assert m.total_parameter_count == g("params").shape[1]
        got = m.lnposterior_batch(g("params"))
        assert rel_err(got, g("lnpost")).max() < LNPOST_RTOL, (n, got, g("lnpost"))
        f = m.model_flux_batch(g("params")[:4])
        assert np.max(np.abs(f - g("model_flux"))) <= FLUX_RTOL * np.max(np.abs(g("model_flux")))
        # the prior, component by component, through the reference-shaped host API
        p = g("params")[0]
        assert np.isfinite(m.prior(p)) == np.isfinite(g("lnpost")[0])
        # the reference's own loop (Model.py:353-364): add_component for each flux, reddening for Extinction
        m.model_spectrum.flux = np.zeros(len(g("wl")))
        q = np.copy(p)
        for comp in m.components:
            step = m.reddening if comp.name == "Extinction" else m.add_component
            step(component=comp, parameters=q[0:comp.parameter_count])
            q = q[comp.parameter_count:]
        assert np.max(np.abs(m.model_spectrum.flux - g("model_flux")[0])) <= FLUX_RTOL * np.max(np.abs(g("model_flux")[0]))
    # Extinction anywhere but last is rejected
    bad = Model()
    bad.components += [Extinction(MW=True)] + _components(["PL"])
    bad.data_spectrum = Spectrum(g("wl"), g("flux"), g("err"))
    with pytest.raises(Exception):
        bad.lnposterior_batch(g("params")[:1, ::-1].copy())


def test_specialised_kernel_on_awkward_grids(monkeypatch):
    """Even-sized grids that the specialised (FULL) instantiation accepts but that stress its edges:
    entirely blueward / redward of the Balmer edge, a single task, a ragged last task, the edge inside
    the first and the last task; each against the oracle, the generic instantiation and every walker
    split."""
    from oracle.spamm_numpy import OracleModel
    from spamm_b200.plan import ModelPlan
    from spamm_b200.Spectrum import Spectrum
    rng = np.random.RandomState(17)
    comps = ["PL", "FE", "HOST", "BC"]
    grids = (np.linspace(1200., 3300., 640), np.linspace(3800., 9000., 1002), np.linspace(3600., 3700., 64),
             np.linspace(3000., 4200., 450), np.linspace(3640., 5200., 1090), np.linspace(1500., 3650., 700),
             1000.0 + np.arange(18000) * 0.5)
    lo = np.array([0, -3, 0, 0, 0, 901, 0, 0, 0, 0, 0, 0, 0, 30, 0, 500, 0, -10, 100, 2], dtype=float)
    hi = np.array([1e-13, 3, 1e-13, 1e-13, 1e-13, 1e4, 1e-14, 1e-14, 1e-14, 1e-14, 1e-14, 1e-14, 1e-14, 1e3,
                   1e-13, 1e5, 2, 10, 1e4, 9], dtype=float)
    for wl in grids:
        o0 = OracleModel(wl, wl, wl, comps=comps)
        fl = [o0.component_flux(cn, TRUTH[cn]) for cn in comps]
        d = np.sum(fl, axis=0) * (1 + 0.01 * rng.standard_normal(len(wl)))
        e = 0.05 * np.abs(d) + 1e-18
        n = 6 if len(wl) > 5000 else 14
        params = lo + (hi - lo) * rng.rand(n, 20)
        params[1, 15] = 600.0                        # cold: large h nu / kT
        params[2, 19] = 8.7                          # SH95 table off its density clamp
        o = OracleModel(wl, d, e, comps=comps)
        o.set_bounds(lo, hi)
        with np.errstate(all="ignore"):
            want = o.ln_posterior_batch(params)
        sp = Spectrum(wl, d, e)

        def plan_for():
            cs = _components(comps)
            for cc in cs:
                cc.initialize(sp)
            return ModelPlan(cs, sp, with_priors=False, bounds=(lo, hi), max_walkers=16)
        full = plan_for()
        assert full.specialised
        got = full.lnposterior_host(params)
        assert rel_err(got, want).max() < LNPOST_RTOL, (len(wl), wl[0], got, want)
        monkeypatch.setenv("SPAMM_NO_FULL_KERNEL", "1")
        gplan = plan_for()
        assert not gplan.specialised
        gen = gplan.lnposterior_host(params)
        monkeypatch.delenv("SPAMM_NO_FULL_KERNEL")
        assert rel_err(got, gen).max() < 1e-12, (len(wl), wl[0])
        for split in (0, 2, 3, 4, 8):
            full.set_walker_split(split)
            assert np.array_equal(full.lnposterior_host(params), got), (len(wl), split)
        assert full.status() == 0


@pytest.mark.parametrize("W", [120, 333, 512, 700, 1024, 1536, 2100])
def test_automatic_split_of_mid_sized_ensembles_is_bit_identical(W):
    """Between ~120 walkers and four waves of CTAs the automatic choice runs some or all walkers of a launch as
    two independent CTAs (red part: pixels from the Balmer edge on + line table; blue part: the pixels before it
    + the continuum stage; per-chunk chi^2 partials through a scratch row, summed by the part that arrives
    second in the unsplit kernel's order) -- for a few waves only the walkers of the last, partial wave (mixed
    grid).  Same bits as one CTA per walker, out-of-prior walkers and the proposal-in-kernel form included."""
    from spamm_b200.synth import full_model_workload
    model, params = full_model_workload(n_pix=8192, n_walkers=W, seed=42)
    params[::17, 1] = 1e3                          # some walkers outside the slope prior -> -inf, early exit
    plan = model.plan
    plan.set_walker_split(0)
    base = plan.lnposterior_host(params)
    assert np.isneginf(base[::17]).all() and np.isfinite(base).sum() > W // 2
    for split in (3, 1):
        plan.set_walker_split(split)
        got = plan.lnposterior_host(params)
        assert np.array_equal(got, base), (W, split)
    assert plan.status() == 0


@pytest.mark.parametrize("name", ["cfg2_nc_fe", "cfg3_nc_hg", "cfg4_nc_bc", "cfg5_full"])
def test_specialised_instantiations_of_the_baseline_configs(name, monkeypatch):
    """BASELINE configs 2-5 each run their own compile-time specialised instantiation (NC+Fe, NC+host,
    NC+Balmer, full); each must agree with the generic instantiation to a few ulp, with the fixture,
    and be independent of the walker split."""
    z, _ = lnpost_cases()
    c = Case(z, name)
    m = _model(c)
    assert m.plan.specialised
    got = m.lnposterior_batch(c.params)
    assert rel_err(got, c.lnpost).max() < LNPOST_RTOL
    monkeypatch.setenv("SPAMM_NO_FULL_KERNEL", "1")
    g = _model(c)
    assert not g.plan.specialised
    ref = g.lnposterior_batch(c.params)
    monkeypatch.delenv("SPAMM_NO_FULL_KERNEL")
    assert rel_err(got, ref).max() < 1e-13
    for split in (2, 3, 4, 8):
        m.plan.set_walker_split(split)
        assert np.array_equal(m.lnposterior_batch(c.params), got)


def test_specialised_extinction_model(monkeypatch):
    """full model + trailing Extinction: specialised instantiation (exp-based extinction factor) vs the
    generic one and the reference fixture."""
    from spamm_b200.Model import Model
    from spamm_b200.Spectrum import Spectrum
    from spamm_b200.components.ReddeningLaw import Extinction
    z = np.load(os.path.join(GOLDEN, "extinction.npz"))
    g = lambda k: z["model/full_mw/%s" % k]  # noqa: E731

    def build():
        m = Model()
        m.components += _components([str(c) for c in g("comps")])
        m.components.append(Extinction(MW=True))
        m.data_spectrum = Spectrum(g("wl"), g("flux"), g("err"))
        return m
    m = build()
    got = m.lnposterior_batch(g("params"))
    assert m.plan.specialised
    assert rel_err(got, g("lnpost")).max() < LNPOST_RTOL
    monkeypatch.setenv("SPAMM_NO_FULL_KERNEL", "1")
    m2 = build()
    ref = m2.lnposterior_batch(g("params"))
    assert not m2.plan.specialised
    assert rel_err(got, ref).max() < 1e-13


_SWEEP_BOUNDS = {
    "PL": ([0., -3.], [1e-13, 3.]),
    "FE": ([0., 0., 0., 901.], [1e-13, 1e-13, 1e-13, 1e4]),
    "HOST": ([0.] * 7 + [30.], [1e-14] * 7 + [1e3]),
    "BC": ([0., 500., 0., -10., 100., 2.], [1e-13, 1e5, 2., 10., 1e4, 9.]),
    "EXT": ([0.], [2.]),
}


def _sweep_case(seed):
    """One seeded random model: grid (uniform / jittered / log-spaced, any size from a handful of pixels
    to a few tasks, anywhere in 1000-10000 A), component subset in any order, Balmer flags, line
    profile, broken power law, optional trailing Extinction."""
    rng = np.random.RandomState(1000 + seed)
    P = int(rng.choice([5, 17, 63, 64, 65, 130, 257, 600, 1023, 1500]))
    a = rng.uniform(1000., 7000.)
    b = min(a + rng.uniform(200., 6000.), 10000.)
    kind = seed % 3
    if kind == 0:
        wl = np.linspace(a, b, P)
    elif kind == 1:
        wl = np.sort(rng.uniform(a, b, P))
        wl = wl + np.arange(P) * 1e-6            # strictly ascending
    else:
        wl = np.exp(np.linspace(np.log(a), np.log(b), P))
    names = ["PL", "FE", "HOST", "BC"]
    k = rng.randint(1, 5)
    comps = [str(c) for c in rng.choice(names, size=k, replace=False)]
    if "FE" in comps and seed % 2 == 0:
        # cover the three Fe templates' normalisation wavelengths (otherwise Fe is NaN and ln L = 0.0:
        # the odd seeds keep that path)
        a, b = rng.uniform(1000., 1900.), rng.uniform(5700., 10000.)
        wl = np.linspace(a, b, P) if kind != 2 else np.exp(np.linspace(np.log(a), np.log(b), P))
        if kind == 1:
            wl = np.sort(rng.uniform(a, b, P)) + np.arange(P) * 1e-6
    if seed % 4 == 0:                            # run_spamm's order half of the time it matters
        comps = [c for c in names if c in comps]
    ext_law = None
    if seed % 5 == 0:
        comps.append("EXT")
        ext_law = ["MW", "AGN", "LMC", "SMC", "Calzetti"][(seed // 5) % 5]
    bc, bpc = [(True, True), (True, False), (False, True)][rng.randint(3)]
    line_type = "gaussian" if seed % 7 == 3 else "lorentzian"
    broken = bool(seed % 6 == 1)
    return wl, comps, ext_law, bc, bpc, line_type, broken, rng


@pytest.mark.parametrize("block", range(4))
def test_randomised_models_vs_oracle(block):
    """Seeded sweep over model shapes the fixtures do not list (40 models, 8 walkers each): every
    ln-posterior through the C ABI against the oracle, <= 1e-10 relative, -inf / 0.0 cases exact."""
    from oracle.spamm_numpy import OracleModel
    from spamm_b200.plan import ModelPlan
    from spamm_b200.Spectrum import Spectrum
    from spamm_b200.components.ReddeningLaw import Extinction
    for seed in range(10 * block, 10 * block + 10):
        wl, comps, ext_law, bc, bpc, line_type, broken, rng = _sweep_case(seed)
        P = len(wl)
        lo, hi = [], []
        for c in comps:
            l, h = _SWEEP_BOUNDS[c]
            if c == "PL" and broken:
                l, h = [wl[0], 0., -3., -3.], [wl[-1], 1e-13, 3., 3.]
            lo += list(l)
            hi += list(h)
        lo, hi = np.array(lo), np.array(hi)
        params = lo + (hi - lo) * rng.rand(8, len(lo))
        params[7, rng.randint(len(lo))] = hi[0] * 0 + 2 * hi.max()          # one walker outside the box
        d = 1e-14 * (wl / wl[P // 2]) ** -1.5 * (1 + 0.05 * rng.standard_normal(P))
        e = 0.05 * np.abs(d) + 1e-18
        o = OracleModel(wl, d, e, comps=comps, bc=bc, bpc=bpc, broken_pl=broken, line_type=line_type,
                        ext_law=ext_law)
        o.set_bounds(lo, hi)
        with np.errstate(all="ignore"):
            want = o.ln_posterior_batch(params)
        sp = Spectrum(wl, d, e)
        cs = _components([c for c in comps if c != "EXT"], bc=bc, bpc=bpc, broken=broken, line_type=line_type)
        if ext_law:
            cs.append(Extinction(**{ext_law: True}))
        for cc in cs:
            cc.initialize(sp)
        plan = ModelPlan(cs, sp, with_priors=False, bounds=(lo, hi), max_walkers=8)
        got = plan.lnposterior_host(params)
        tag = (seed, P, comps, bc, bpc, line_type, broken, ext_law)
        assert np.array_equal(np.isneginf(got), np.isneginf(want)), tag
        assert np.isneginf(got[7]), tag
        assert rel_err(got, want).max() < LNPOST_RTOL, (tag, got, want)


def test_gaussian_line_profile_cutoff_is_exact():
    """bc_line_type = gaussian (BC:318-319): the kernel visits only the lines within 27.4 widths of a task
    (the others are exactly 0.0 in the reference as well, exp underflow).  Narrow to broad widths, on a grid
    spanning the whole line forest, against the oracle's full 397 x P sum."""
    from oracle.spamm_numpy import OracleModel
    from spamm_b200.plan import ModelPlan
    from spamm_b200.Spectrum import Spectrum
    wl = grid(2048)
    comps = ["PL", "BC"]
    lo = np.array([0., -3., 0., 500., 0., -10., 100., 2.])
    hi = np.array([1e-13, 3., 1e-13, 1e5, 2., 10., 1e4, 9.])
    rng = np.random.RandomState(5)
    params = lo + (hi - lo) * rng.rand(12, 8)
    params[:, 6] = [100.001, 150., 300., 700., 1500., 3000., 6000., 9999., 120., 450., 2000., 5000.]   # lwidth
    params[:, 5] = np.linspace(-9.9, 9.9, 12)                                                            # loffset
    d = 1e-14 * (wl / 5500.) ** -1.5
    e = 0.05 * d
    for bc in (True, False):
        o = OracleModel(wl, d, e, comps=comps, bc=bc, bpc=True, line_type="gaussian")
        o.set_bounds(lo, hi)
        with np.errstate(all="ignore"):
            want = o.ln_posterior_batch(params)
            want_f = np.array([o.model_flux(p) for p in params])
        sp = Spectrum(wl, d, e)
        cs = _components(comps, bc=bc, bpc=True, line_type="gaussian")
        for cc in cs:
            cc.initialize(sp)
        plan = ModelPlan(cs, sp, with_priors=False, bounds=(lo, hi), max_walkers=16)
        got = plan.lnposterior_host(params)
        assert rel_err(got, want).max() < LNPOST_RTOL, (bc, got, want)
        got_f = plan.model_flux_host(params)
        assert np.max(np.abs(got_f - want_f) / np.max(np.abs(want_f), axis=1, keepdims=True)) < FLUX_RTOL


@pytest.mark.parametrize("comps", [["PL", "FE", "HOST"], ["PL", "FE", "BC"], ["PL", "HOST", "BC"]])
def test_specialised_three_component_models(comps, monkeypatch):
    """The three-component subsets in run_spamm's order (e.g. PL + FE + BC, the usual UV fit without a host)
    have their own specialised instantiations: against the oracle, the generic instantiation and every
    walker split."""
    from oracle.spamm_numpy import OracleModel
    wl = grid(2048)
    o0 = OracleModel(wl, wl, wl, comps=comps)
    fl = [o0.component_flux(c, TRUTH[c]) for c in comps]
    rng = np.random.RandomState(17)
    d = np.sum(fl, axis=0) * (1 + 0.02 * rng.standard_normal(len(wl)))
    e = np.sqrt(np.sum([(0.05 * f) ** 2 for f in fl], axis=0))
    o = OracleModel(wl, d, e, comps=comps)
    np.random.seed(23)
    params = np.array([o.initial_values() for _ in range(40)])
    params[3, 0] = -1.0                                          # one walker outside the prior box
    want = o.ln_posterior_batch(params)
    c = Case.__new__(Case)
    c.wl, c.flux, c.err, c.comps, c.bc, c.bpc = wl, d, e, comps, True, True
    m = _model(c)
    assert m.plan.specialised
    got = m.lnposterior_batch(params)
    assert np.isneginf(got[3]) and np.isneginf(want[3])
    assert rel_err(got, want).max() < LNPOST_RTOL
    monkeypatch.setenv("SPAMM_NO_FULL_KERNEL", "1")
    g = _model(c)
    assert not g.plan.specialised
    ref = g.lnposterior_batch(params)
    monkeypatch.delenv("SPAMM_NO_FULL_KERNEL")
    assert rel_err(got, ref).max() < 1e-13
    for split in (2, 3, 4, 8):
        m.plan.set_walker_split(split)
        assert np.array_equal(m.lnposterior_batch(params), got)


@pytest.mark.parametrize("comps", [["PL", "BC"], ["PL", "FE", "HOST", "BC"]])
@pytest.mark.parametrize("bc,bpc", [(True, False), (False, True)])
def test_specialised_balmer_with_one_part_off(comps, bc, bpc, monkeypatch):
    """run_spamm's "BC" without "BPC" is the Balmer continuum alone (run_spamm.py:118-121; the reference's
    examples/test_spamm.py fits exactly that), "BPC" alone the line forest alone: both have specialised
    instantiations for NC + Balmer and for the full model."""
    from oracle.spamm_numpy import OracleModel
    wl = grid(2048)
    o0 = OracleModel(wl, wl, wl, comps=comps)
    fl = [o0.component_flux(c, TRUTH[c]) for c in comps]
    rng = np.random.RandomState(29)
    d = np.sum(fl, axis=0) * (1 + 0.02 * rng.standard_normal(len(wl)))
    e = np.sqrt(np.sum([(0.05 * f) ** 2 for f in fl], axis=0))
    o = OracleModel(wl, d, e, comps=comps, bc=bc, bpc=bpc)
    np.random.seed(31)
    params = np.array([o.initial_values() for _ in range(40)])
    params[5, 1] = 3.5                                           # slope outside the prior box
    want = o.ln_posterior_batch(params)
    c = Case.__new__(Case)
    c.wl, c.flux, c.err, c.comps, c.bc, c.bpc = wl, d, e, comps, bc, bpc
    m = _model(c)
    assert m.plan.specialised
    got = m.lnposterior_batch(params)
    assert np.isneginf(got[5]) and np.isneginf(want[5])
    assert rel_err(got, want).max() < LNPOST_RTOL
    monkeypatch.setenv("SPAMM_NO_FULL_KERNEL", "1")
    g = _model(c)
    assert not g.plan.specialised
    ref = g.lnposterior_batch(params)
    monkeypatch.delenv("SPAMM_NO_FULL_KERNEL")
    assert rel_err(got, ref).max() < 1e-13
    for split in (2, 3, 4, 8):
        m.plan.set_walker_split(split)
        assert np.array_equal(m.lnposterior_batch(params), got)
    assert m.plan.status() == 0
