#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); imports the
reference through oracle/refshim.py and writes

  tests/golden/kat_pickle.npz   known-answer pairs from examples/model.pickle.gz
                                
  tests/golden/components.npz   per-component flux() of the reference on three
                                grids (NC, Fe, host, BC, BpC, BC+BpC, Gaussian
                                BpC, broken power law)
  tests/golden/lnpost.npz       ln_posterior() of the reference for BASELINE.json
                                configs 1-5 (reduced walker counts) + quirk cases

    python tests/golden/make_golden.py
"""
import gzip
import io
import os
import pickle
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import refshim  # noqa: E402

ref = refshim.load()

NC_TRUTH = [5e-15, 2.3]                                            # test_spamm.py:16-19
FE_TRUTH = [1.07988504e-14, 6.91877436e-15, 5e-15, 5450.]          # :24-29
HG_TRUTH = [5e-17, 5.5e-17, 5e-16, 1e-16, 2e-16, 3e-16, 4e-16, 515.]
BC_TRUTH = [3e-14, 50250., 1.0, 0.0, 5050., 5.5]                   # :32-39
TRUTH = {"PL": NC_TRUTH, "FE": FE_TRUTH, "HOST": HG_TRUTH, "BC": BC_TRUTH}

C_KMS = 299792.458


def grid(P):
    return 1000.0 + np.arange(P) * (9000.0 / P)


def make_component(c, bc=True, bpc=True, broken=False, line_type=None):
    if c == "PL":
        return ref.NuclearContinuumComponent(broken=broken)
    if c == "FE":
        return ref.FeComponent()
    if c == "HOST":
        return ref.HostGalaxyComponent()
    if c == "BC":
        pars = None
        if line_type is not None:
            from utils.parse_pars import parse_pars
            pars = parse_pars()["balmer_continuum"]
            pars["bc_line_type"] = line_type
        return ref.BalmerCombined(pars=pars, BalmerContinuum=bc, BalmerPseudocContinuum=bpc)
    raise KeyError(c)


def synth(wl, comps, truth=None):
    """examples/test_spamm.py:53-99: data = sum of component fluxes at the
    truths, err = 5 % per component in quadrature."""
    dummy = ref.Spectrum(wl, wl, wl)
    fluxes, errs = [], []
    for c in comps:
        comp = make_component(c)
        comp.initialize(dummy)
        f = np.array(comp.flux(dummy, np.array((truth or TRUTH)[c])))
        fluxes.append(f)
        errs.append(0.05 * f)
    return np.sum(fluxes, axis=0), np.sqrt(np.sum([e ** 2 for e in errs], axis=0))


def build_model(wl, d, e, comps, **kw):
    m = ref.Model()
    for c in comps:
        m.components.append(make_component(c, **kw))
    m.data_spectrum = ref.Spectrum(wl, d, e)
    return m


def draw_walkers(m, n, seed=42):
    """Model.run_mcmc :254-259."""
    np.random.seed(seed)
    out = []
    for _ in range(n):
        p = []
        for comp in m.components:
            p = p + list(comp.initial_values(m.data_spectrum))
        out.append(p)
    return np.array(out, dtype=np.float64)


def bounds_of(m):
    lo, hi = [], []
    for comp in m.components:
        if comp.name == "Nuclear":
            if comp.broken_pl:
                lo += [comp.wave_break_min, comp.norm_min, comp.slope_min, comp.slope_min]
                hi += [comp.wave_break_max, comp.norm_max, comp.slope_max, comp.slope_max]
            else:
                lo += [comp.norm_min, comp.slope_min]
                hi += [comp.norm_max, comp.slope_max]
        elif comp.name == "FeForest":
            n = len(comp.fe_templ)
            lo += [comp.norm_min] * n + [comp.width_min]
            hi += [comp.norm_max] * n + [comp.width_max]
        elif comp.name == "HostGalaxy":
            n = len(comp.host_gal)
            lo += [comp.norm_min] * n + [comp.stellar_disp_min]
            hi += [comp.norm_max] * n + [comp.stellar_disp_max]
        elif comp.name == "Balmer":
            lo += [comp.normalization_min, comp.Te_min, comp.tauBE_min, comp.loffset_min,
                   comp.lwidth_min, comp.logNe_min]
            hi += [comp.normalization_max, comp.Te_max, comp.tauBE_max, comp.loffset_max,
                   comp.lwidth_max, comp.logNe_max]
    return np.array(lo, dtype=np.float64), np.array(hi, dtype=np.float64)


def near_truth(comps, n, rng, lo, hi, frac=0.01):
    t = np.concatenate([TRUTH[c] for c in comps])
    # reduced models can push a truth outside its data-derived box (fnw, App. B Q12)
    t = np.where((t > lo) & (t < hi), t, 0.5 * (lo + hi))
    out = t * (1.0 + frac * rng.standard_normal((n, len(t))))
    out[:, t == 0.0] = 0.1 * rng.standard_normal((n, int((t == 0).sum())))
    return out


def fe_width_straddles(m, comps, base, rng):
    """Widths a hair either side of the integer sigma_norm boundaries
    (np.ceil, FeComponent.py:290 / HostGalaxyComponent.py:297)."""
    rows = []
    k0 = 0
    for c, comp in zip(comps, m.components):
        n = comp.parameter_count
        if c == "FE":
            for t, kk in [(0, 7), (0, 31), (1, 20), (2, 40), (2, 74)]:
                bin_size = comp.log_fe[t].spectral_axis[2] - comp.log_fe[t].spectral_axis[1]
                w = np.sqrt((kk * bin_size * C_KMS * 2. * np.sqrt(2. * np.log(2.))) ** 2 + 900. ** 2)
                for eps in (-1e-9, 1e-9):
                    r = base.copy()
                    r[k0 + n - 1] = w * (1 + eps)
                    rows.append(r)
        if c == "HOST":
            for kk in (1, 2):
                bin_size = comp.log_host[0].spectral_axis[2] - comp.log_host[0].spectral_axis[1]
                w = kk * bin_size * C_KMS * 2. * np.sqrt(2. * np.log(2.))
                for eps in (-1e-9, 1e-9):
                    r = base.copy()
                    r[k0 + n - 1] = w * (1 + eps)
                    rows.append(r)
        k0 += n
    return np.array(rows) if rows else np.zeros((0, len(base)))


def lnpost_case(out, name, wl, d, e, comps, n_prior, n_truth=0, n_out=0, straddle=False,
                extra=None, **kw):
    m = build_model(wl, d, e, comps, **kw)
    rng = np.random.RandomState(1234)
    sets = [draw_walkers(m, n_prior)]
    lo, hi = bounds_of(m)
    if n_truth:
        nt = near_truth(comps, n_truth, rng, lo, hi)
        sets.append(nt)
        if n_out:
            bad = nt[:n_out].copy()
            for i in range(n_out):
                j = rng.randint(bad.shape[1])
                bad[i, j] = hi[j] * 1.01 + 1.0 if i % 2 == 0 else lo[j] - abs(lo[j]) * 0.01 - 1e-30
            # exactly-on-the-bound rows (strict inequalities => -inf)
            edge = nt[:2].copy()
            edge[0, 1] = hi[1]
            edge[1, 1] = lo[1]
            sets += [bad, edge]
        if straddle:
            sets.append(fe_width_straddles(m, comps, nt[0], rng))
    if extra is not None:
        sets.append(np.atleast_2d(extra))
    params = np.concatenate(sets, axis=0)
    lnp = np.array([ref.ln_posterior(p, m) for p in params], dtype=np.float64)
    for k, v in dict(wl=wl, flux=d, err=e, params=params, lnpost=lnp, lo=lo, hi=hi,
                     comps=np.array(comps), bc=np.array(kw.get("bc", True)),
                     bpc=np.array(kw.get("bpc", True))).items():
        out["%s/%s" % (name, k)] = v
    print("lnpost case %-16s W=%3d  P=%5d  finite=%d  zero=%d" %
          (name, len(params), len(wl), np.isfinite(lnp).sum(), (lnp == 0).sum()))


# ----------------------------------------------------------------------------
def make_kat():
    class Stub(object):
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, st):
            if isinstance(st, dict):
                self.__dict__.update(st)

    class Q(np.ndarray):
        def __setstate__(self, st):
            np.ndarray.__setstate__(self, st[0])

    class U(pickle.Unpickler):
        def find_class(self, module, name):
            if module.startswith("numpy.core"):
                module = module.replace("numpy.core", "numpy._core")
            if module.startswith("numpy"):
                return super().find_class(module, name)
            if module == "astropy.units.quantity" and name == "Quantity":
                return Q
            if module.startswith(("spamm", "emcee", "astropy", "specutils")):
                return type(name, (Stub,), {})
            return super().find_class(module, name)

    raw = gzip.open(os.path.join(refshim.REFERENCE_ROOT, "examples/model.pickle.gz"), "rb").read()
    m = U(io.BytesIO(raw), encoding="latin1").load()
    ds, s, c = m._data_spectrum, m.sampler, m.components[0]
    chain, lnprob = np.asarray(s._chain), np.asarray(s._lnprob)
    flat = chain.reshape(-1, 2)
    lp = lnprob.reshape(-1)
    idx = np.unique(np.concatenate([np.arange(0, len(lp), 37), np.argsort(lp)[:50],
                                    np.argsort(lp)[-50:]]))
    # old-version convention F = norm (lam/lam0)^(+slope)  =>  slope1 = -slope (App. B Q16)
    params = np.stack([flat[idx, 0], -flat[idx, 1]], axis=1)
    samples = chain[:, 50:, :].reshape(-1, 2)                        # Samples.py:75
    med = np.median(samples, axis=0)
    # analysis.median_values :514-540 (index-based 68 % interval)
    srt = np.sort(samples, axis=0)
    n = len(srt)
    lo16 = srt[int(n * 0.16)]
    hi84 = srt[int(n * 0.84)]
    out = dict(wl=np.asarray(ds._dispersion, dtype=np.float64),
               flux=np.asarray(ds._data, dtype=np.float64),
               err=np.asarray(ds.flux_error, dtype=np.float64),
               norm_wavelength=np.float64(c._norm_wavelength),
               lo=np.array([c.normalization_min, -c.slope_max], dtype=np.float64),
               hi=np.array([c.normalization_max, -c.slope_min], dtype=np.float64),
               params=params, lnprob=lp[idx],
               n_total=np.int64(len(lp)), acceptance=np.float64(s.naccepted.mean() / s.iterations),
               post_median=med, post_p16=lo16, post_p84=hi84,
               walkers0=chain[:, 0, :])
    np.savez_compressed(os.path.join(HERE, "kat_pickle.npz"), **out)
    print("kat: %d of %d pairs; medians %s" % (len(idx), len(lp), med))


def make_components():
    out = {}
    rng = np.random.RandomState(7)
    fp6 = np.loadtxt(os.path.join(refshim.REFERENCE_ROOT,
                                  "Data/FakeData/PLcompOnly/fakepowlaw6_werr.dat"))[:, 0]
    grids = {"g8192": (grid(8192), 2), "g4096": (grid(4096), 3), "fp6": (fp6, 4)}
    for gname, (wl, ndraw) in grids.items():
        out["%s/wl" % gname] = wl
        d, e = synth(wl, ["PL", "FE", "HOST", "BC"])
        variants = [("PL", {}), ("FE", {}), ("HOST", {}), ("BC", dict(bc=True, bpc=False)),
                    ("BC", dict(bc=False, bpc=True)), ("BC", dict(bc=True, bpc=True))]
        if gname == "fp6":
            variants += [("BC", dict(bc=False, bpc=True, line_type="gaussian")),
                         ("PL", dict(broken=True))]
        for c, kw in variants:
            m = build_model(wl, d, e, [c], **kw)
            params = draw_walkers(m, ndraw, seed=11)
            params[0] = TRUTH[c] if not kw.get("broken") else [4000., 5e-15, 2.3, 1.1]
            if c == "BC" and ndraw > 1:
                params[1, 1], params[1, 5] = 12000., 8.6    # off the SH95 clamp
            if c == "BC" and ndraw > 2:
                params[2, 1], params[2, 5] = 700., 8.95
            comp = m.components[0]
            flux = np.array([np.array(comp.flux(m.data_spectrum, p)) for p in params])
            tag = c + "".join("_%s%s" % (k, v) for k, v in sorted(kw.items()))
            out["%s/%s/params" % (gname, tag)] = params
            out["%s/%s/flux" % (gname, tag)] = flux
            print("component %-6s %-28s %s" % (gname, tag, flux.shape))
    np.savez_compressed(os.path.join(HERE, "components.npz"), **out)


def make_lnpost():
    out = {}
    # config 1: NC on Data/FakeData/PLcompOnly/fakepowlaw1_werr.dat, 32 walkers
    fp1 = np.loadtxt(os.path.join(refshim.REFERENCE_ROOT,
                                  "Data/FakeData/PLcompOnly/fakepowlaw1_werr.dat"))
    lnpost_case(out, "cfg1_nc_fp1", fp1[:, 0], fp1[:, 1], fp1[:, 2], ["PL"], 32,
                extra=[[1.8e-17, 2.2], [1.7987789432468136e-17, 2.2016429312441748]])
    # config 2: NC + Fe, 4k pixels
    wl = grid(4096)
    d, e = synth(wl, ["PL", "FE"])
    lnpost_case(out, "cfg2_nc_fe", wl, d, e, ["PL", "FE"], 24, n_truth=8, n_out=4, straddle=True)
    # config 3: NC + host (all 7 Data/HostModels templates)
    wl = grid(8192)
    d, e = synth(wl, ["PL", "HOST"])
    lnpost_case(out, "cfg3_nc_hg", wl, d, e, ["PL", "HOST"], 24, n_truth=8, n_out=4, straddle=True)
    # config 4: NC + Balmer continuum with high-order lines
    d, e = synth(wl, ["PL", "BC"])
    lnpost_case(out, "cfg4_nc_bc", wl, d, e, ["PL", "BC"], 24, n_truth=8, n_out=4)
    # config 5: full model, 8192 px
    comps = ["PL", "FE", "HOST", "BC"]
    d, e = synth(wl, comps)
    lnpost_case(out, "cfg5_full", wl, d, e, comps, 40, n_truth=16, n_out=8, straddle=True,
                extra=np.concatenate([TRUTH[c] for c in comps]))
    # quirk Q6: Fe template missing from the grid => NaN => lnL == 0.0
    lnpost_case(out, "q6_fe_nan", fp1[:, 0], fp1[:, 1], fp1[:, 2], ["PL", "FE"], 6)
    # full model on a non-uniform-in-index real grid (fakepowlaw6: 1400-11999, 3 A)
    fp6 = np.loadtxt(os.path.join(refshim.REFERENCE_ROOT,
                                  "Data/FakeData/PLcompOnly/fakepowlaw6_werr.dat"))[:, 0]
    d, e = synth(fp6, comps)
    lnpost_case(out, "full_fp6", fp6, d, e, comps, 16, n_truth=8, n_out=2)
    np.savez_compressed(os.path.join(HERE, "lnpost.npz"), **out)


if __name__ == "__main__":
    with refshim.chdir_examples():
        make_kat()
        make_components()
        make_lnpost()
    for f in ("kat_pickle.npz", "components.npz", "lnpost.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
