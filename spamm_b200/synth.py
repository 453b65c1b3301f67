"""Synthetic-spectrum generators and the BASELINE.json workloads.

Mirrors examples/run_{nc,fe,hg,bc}.py::create_* and examples/test_spamm.py:53-99:
a component is instantiated, initialised on a dummy ``Spectrum(wl, wl, wl)``, its
flux at hand-picked truths becomes the data and 5 % of it the error; fluxes are
summed and errors added in quadrature.  All fluxes come from the GPU path."""
import numpy as np

from .Spectrum import Spectrum
from .Model import Model
from .components.NuclearContinuumComponent import NuclearContinuumComponent
from .components.FeComponent import FeComponent
from .components.HostGalaxyComponent import HostGalaxyComponent
from .components.BalmerContinuumCombined import BalmerCombined
from .utils.add_in_quadrature import add_in_quadrature
from .utils.parse_pars import parse_pars

TEST_WL = parse_pars()["testing"]
WL = np.arange(TEST_WL["wl_min"], TEST_WL["wl_max"], TEST_WL["wl_step"])

# examples/test_spamm.py:16-47 (host: 7-template vector)
NC_PARAMS = {"wl": WL, "slope1": 2.3, "norm_PL": 5e-15, "broken_pl": False}
FE_PARAMS = {"fe_norm_1": 1.07988504e-14, "fe_norm_2": 6.91877436e-15, "fe_norm_3": 5e-15,
             "fe_width": 5450, "no_templates": 3, "wl": WL}
BC_PARAMS = {"bc_norm": 3e-14, "bc_tauBE": 1., "bc_logNe": 5.5, "bc_loffset": 0., "bc_lwidth": 5050.,
             "bc_Te": 50250., "bc_lines": 201.5, "wl": WL}
HG_PARAMS = {"hg_norm_1": 5e-17, "hg_norm_2": 5.5e-17, "hg_norm_3": 5e-16, "hg_norm_4": 1e-16,
             "hg_norm_5": 2e-16, "hg_norm_6": 3e-16, "hg_norm_7": 4e-16, "hg_stellar_disp": 515,
             "no_templates": 7, "wl": WL}


def _with_wl(params, wl):
    p = dict(params)
    if wl is not None:
        p["wl"] = wl
    return p


def create_nc(nc_params=None, wl=None):
    p = _with_wl(nc_params or NC_PARAMS, wl)
    nc = NuclearContinuumComponent(broken=p["broken_pl"])
    spectrum = Spectrum(p["wl"], p["wl"], p["wl"])
    nc.initialize(spectrum)
    if p["broken_pl"]:
        comp_params = [p["wave_break"], p["norm_PL"], p["slope1"], p["slope2"]]
    else:
        comp_params = [p["norm_PL"], p["slope1"]]
    flux = nc.flux(spectrum, comp_params)
    return p["wl"], flux, flux * 0.05, p


def create_fe(fe_params=None, wl=None):
    p = _with_wl(fe_params or FE_PARAMS, wl)
    fe = FeComponent()
    spectrum = Spectrum(p["wl"], p["wl"], p["wl"])
    fe.initialize(spectrum)
    comp_params = [p["fe_norm_{}".format(x)] for x in range(1, p["no_templates"] + 1)] + [p["fe_width"]]
    flux = fe.flux(spectrum, comp_params)
    return p["wl"], flux, flux * 0.05, p


def create_hg(hg_params=None, wl=None):
    p = _with_wl(hg_params or HG_PARAMS, wl)
    hg = HostGalaxyComponent()
    spectrum = Spectrum(p["wl"], p["wl"], p["wl"])
    hg.initialize(spectrum)
    comp_params = [p["hg_norm_{}".format(x)] for x in range(1, p["no_templates"] + 1)] + [p["hg_stellar_disp"]]
    flux = hg.flux(spectrum, comp_params)
    return p["wl"], flux, flux * 0.05, p


def create_bc(bc_params=None, wl=None):
    p = _with_wl(bc_params or BC_PARAMS, wl)
    bc = BalmerCombined(BalmerContinuum=True, BalmerPseudocContinuum=True)
    spectrum = Spectrum(p["wl"], p["wl"], p["wl"])
    bc.initialize(spectrum)
    comp_params = [p[x] for x in ["bc_norm", "bc_Te", "bc_tauBE", "bc_loffset", "bc_lwidth", "bc_logNe"]]
    flux = bc.flux(spectrum, comp_params)
    return p["wl"], flux, flux * 0.05, p


_CREATE = {"PL": create_nc, "FE": create_fe, "HOST": create_hg, "BC": create_bc}


def synthetic_spectrum(components, wl):
    """(wl, flux, err) exactly as examples/test_spamm.py:98-99 combines them."""
    fluxes, errs = [], []
    for c in components:
        _, f, e, _ = _CREATE[c](wl=wl)
        fluxes.append(f)
        errs.append(e)
    return wl, np.sum(fluxes, axis=0), add_in_quadrature(errs)


def config_grid(n_pix):
    """BASELINE grids: lambda_i = 1000 + i * 9000 / P."""
    return 1000.0 + np.arange(n_pix) * (9000.0 / n_pix)


def build_model(components, wl, flux, err, max_walkers=8192, template_mode=None):
    """Model with the component list in run_spamm.py:100-122 order."""
    model = Model()
    model.max_walkers = max_walkers
    if template_mode is not None:
        model.template_mode = template_mode
    for c in components:
        if c == "PL":
            model.components.append(NuclearContinuumComponent(broken=False))
        elif c == "FE":
            model.components.append(FeComponent())
        elif c == "HOST":
            model.components.append(HostGalaxyComponent())
        elif c == "BC":
            model.components.append(BalmerCombined(BalmerContinuum=True, BalmerPseudocContinuum=True))
        else:
            raise KeyError(c)
    model.data_spectrum = Spectrum(wl, flux, err)
    return model


def full_model_workload(n_pix=8192, n_walkers=8192, seed=42, components=("PL", "FE", "HOST", "BC")):
    """BASELINE.json config 5: full NC+Fe+HG+BC model, 8192 walkers x 8k pixels;
    walkers drawn from the priors with np.random.seed(42) (Model.run_mcmc:255-259)."""
    wl = config_grid(n_pix)
    wl, flux, err = synthetic_spectrum(components, wl)
    model = build_model(list(components), wl, flux, err, max_walkers=n_walkers)
    np.random.seed(seed)
    params = np.array(model.initial_walkers(n_walkers), dtype=np.float64)
    return model, params
