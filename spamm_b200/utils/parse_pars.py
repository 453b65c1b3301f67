"""Input parameters (priors, template locations, kernel sizes).

Mirrors the dictionary the reference builds from ``utils/parameters.yaml``
(utils/parse_pars.py:8-27): same sections and keys, so reference-style
``pars=parse_pars()["fe_forest"]`` constructor arguments keep working.  The
defaults live here as a Python literal; a user yaml file with the reference's
schema can be passed as ``par_file`` and is merged over them.

Differences that matter: the yaml is read with ``yaml.safe_load`` (the
reference's bare ``yaml.load`` fails on PyYAML >= 6), units stay plain strings
(no astropy), and the template locations default to the data bundled with this
package instead of cwd-relative ``../Data/...`` paths.
"""
import copy
import os

BUNDLED = "bundled"   # sentinel: read templates from spamm_b200/data/spamm_data.npz

_DEFAULTS = {
    "global": {"flux_unit": "erg cm-2 s-1 AA-1", "wl_unit": "AA"},
    "emission_lines": "",
    "host_galaxy": {
        "boxcar_width": 5,
        "hg_models": BUNDLED,
        "hg_norm_min": 0.,
        "hg_norm_max": "max_flux",
        "hg_stellar_disp_min": 30.,
        "hg_stellar_disp_max": 1000.,
        "hg_template_stellar_disp": 0.0,
        "hg_kernel_size_sigma": 10,
    },
    "boxcar_width": 5,
    "nuclear_continuum": {
        "boxcar_width": 5,
        "broken_pl": False,
        "pl_slope_min": -3.,
        "pl_slope_max": 3.,
        "pl_norm_min": 0.,
        "pl_norm_max": "fnw",
        "pl_wave_break_min": "min_wl",
        "pl_wave_break_max": "max_wl",
    },
    "balmer_continuum": {
        "bc_line_type": "lorentzian",
        "bc_lines_min": 3.,
        "bc_lines_max": 400.,
        "bc_norm_min": 0.,
        "bc_norm_max": "bcmax_flux",
        "bc_Te_min": 500.,
        "bc_Te_max": 100000.,
        "bc_tauBE_min": 0.,
        "bc_tauBE_max": 2.,
        "bc_loffset_min": -10.,
        "bc_loffset_max": 10.,
        "bc_lwidth_min": 100.,
        "bc_lwidth_max": 10000.,
        "bc_logNe_min": 2,
        "bc_logNe_max": 9,
        "sh95_coefficients": BUNDLED,
    },
    "fe_forest": {
        "boxcar_width": 10,
        "fe_templates": BUNDLED,
        "fe_template_width": 900.,
        "fe_norm_min": 0.,
        "fe_norm_max": "fnw",
        "fe_width_min": 901.,
        "fe_width_max": 10000.,
        "fe_line_type": "gaussian",
        "fe_kernel_size_sigma": 3,
    },
    "rebin_spec": False,
    "testing": {"wl_min": 1000, "wl_max": 10000, "wl_step": 0.5},
}


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v


def parse_pars(par_file=None):
    """Return the SPAMM input-parameter dictionary.

    Args:
        par_file (str): optional yaml file in the reference's schema.
    Returns:
        pars (dict)
    """
    pars = copy.deepcopy(_DEFAULTS)
    if par_file is not None:
        assert os.path.exists(par_file), \
            "Input parameters {0} is not in {1}".format(par_file, os.getcwd())
        import yaml
        with open(par_file, "r") as f:
            user = yaml.safe_load(f) or {}
        _merge(pars, user)
    return pars
