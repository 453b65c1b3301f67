"""Build a spamm_b200 Model from a *reference* SPAMM Model object (duck-typed:
nothing from the reference is imported here).  Used with `GpuPool` to keep the
reference's own driver + emcee and swap only the ln-posterior evaluation
(INTEGRATION.md section 1)."""
import numpy as np

from .Model import Model
from .Spectrum import Spectrum
from .components.NuclearContinuumComponent import NuclearContinuumComponent
from .components.FeComponent import FeComponent
from .components.HostGalaxyComponent import HostGalaxyComponent
from .components.BalmerContinuumCombined import BalmerCombined
from .utils.parse_pars import BUNDLED


def _templates_from(ref_templates):
    return [Spectrum(spectral_axis=np.asarray(t.spectral_axis, dtype=np.float64),
                     flux=np.asarray(t.flux, dtype=np.float64),
                     flux_error=np.asarray(t.flux, dtype=np.float64)) for t in ref_templates]


def component_from_reference(rc):
    """Same-named component with the reference component's parameters, resolved
    prior attributes and (already loaded) templates in the reference's order."""
    pars = dict(rc.inputpars)
    if rc.name == "Nuclear":
        c = NuclearContinuumComponent(pars=pars, broken=rc.broken_pl)
        for a in ("norm_min", "norm_max", "slope_min", "slope_max", "wave_break_min", "wave_break_max"):
            setattr(c, a, getattr(rc, a))
    elif rc.name == "FeForest":
        pars["fe_templates"] = BUNDLED
        c = FeComponent(pars=pars)
        c.fe_templ = _templates_from(rc.fe_templ)           # the reference's own files and order
        c.model_parameter_names = list(rc.model_parameter_names)
        for a in ("norm_min", "norm_max", "width_min", "width_max", "templ_width"):
            setattr(c, a, getattr(rc, a))
    elif rc.name == "HostGalaxy":
        pars["hg_models"] = BUNDLED
        c = HostGalaxyComponent(pars=pars)
        c.host_gal = _templates_from(rc.host_gal)           # glob order of the reference process
        c.model_parameter_names = list(rc.model_parameter_names)
        for a in ("norm_min", "norm_max", "stellar_disp_min", "stellar_disp_max", "templ_stellar_disp"):
            setattr(c, a, getattr(rc, a))
    elif rc.name == "Balmer":
        pars.setdefault("sh95_coefficients", BUNDLED)
        c = BalmerCombined(pars=pars, BalmerContinuum=rc.BC, BalmerPseudocContinuum=rc.BpC)
        for a in ("normalization_min", "normalization_max", "Te_min", "Te_max", "tauBE_min", "tauBE_max",
                  "loffset_min", "loffset_max", "lwidth_min", "lwidth_max", "logNe_min", "logNe_max"):
            setattr(c, a, getattr(rc, a))
    else:
        raise ValueError("component %r has no GPU implementation" % (rc.name,))
    return c


def model_from_reference(ref_model, max_walkers=8192):
    m = Model()
    m.max_walkers = max_walkers
    for rc in ref_model.components:
        m.components.append(component_from_reference(rc))
    ds = ref_model.data_spectrum
    m.data_spectrum = Spectrum(np.asarray(ds.spectral_axis, dtype=np.float64),
                               np.asarray(ds.flux, dtype=np.float64),
                               np.asarray(ds.flux_error, dtype=np.float64))
    return m
