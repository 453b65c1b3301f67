"""Top-level driver with the reference's signature (spamm/run_spamm.py:26-167):
build the model from a component list, run the ensemble sampler on the GPU,
save {"model", "comp_params"} as a gzip pickle.  Plots are out of scope."""
import datetime
import gzip
import os
import pickle

import numpy as np

from .utils.parse_pars import parse_pars
from .Spectrum import Spectrum
from .Model import Model
from .components.NuclearContinuumComponent import NuclearContinuumComponent
from .components.HostGalaxyComponent import HostGalaxyComponent
from .components.FeComponent import FeComponent
from .components.BalmerContinuumCombined import BalmerCombined
from .components.ReddeningLaw import Extinction

# (the reference lists the five *_EXT reddening laws too but cannot run them: NameError at
# run_spamm.py:124; here they append an Extinction component as that line intends)
ACCEPTED_COMPS = ["PL", "FE", "HOST", "BC", "BPC", "CALZETTI_EXT", "SMC_EXT", "MW_EXT", "AGN_EXT", "LMC_EXT"]


def spamm(complist, inspectrum, par_file=None, n_walkers=30, n_iterations=500, outdir=None,
          picklefile=None, comp_params=None, seed=None, save=True):
    """
    Args:
        complist (list): component names, case insensitive: PL, FE, HOST, BC, BPC and at most one
            reddening law of CALZETTI_EXT, SMC_EXT, MW_EXT, AGN_EXT, LMC_EXT.
        inspectrum (Spectrum or tuple): a Spectrum, or (wavelength, flux[, flux_error]).
        par_file (str): parameter file in the reference's yaml schema.
        n_walkers, n_iterations (int): ensemble size and chain length.
        outdir, picklefile (str): where the gzip pickle goes (names default to the run time).
        comp_params (dict): known parameter values to carry along in the output.
        seed (int): RNG seed of the device sampler.  save (bool): write the pickle.
    Returns:
        dict with "model" and "comp_params".
    """
    t1 = datetime.datetime.now()
    pars = parse_pars() if par_file is None else parse_pars(par_file)
    complist = [x.upper() for x in complist]
    for x in complist:
        if x not in ACCEPTED_COMPS:
            raise ValueError("component {0!r} is not supported on the GPU path; accepted: {1}".format(
                x, ACCEPTED_COMPS))
    components = {k: (k in complist) for k in ACCEPTED_COMPS}

    if isinstance(inspectrum, Spectrum):
        spectrum = inspectrum
        wl, flux, flux_error = inspectrum.spectral_axis, inspectrum.flux, inspectrum.flux_error
    else:
        if len(inspectrum) == 3:
            wl, flux, flux_error = inspectrum
        else:
            wl, flux = inspectrum
            flux_error = np.asarray(flux) * 0.05
        spectrum = Spectrum(spectral_axis=wl, flux=flux, flux_error=flux_error)

    comp_params = {} if comp_params is None else comp_params
    for k, v in zip(("wl", "flux", "err", "components"), (wl, flux, flux_error, components)):
        comp_params.setdefault(k, v)

    model = Model()
    model.print_parameters = False
    model.max_walkers = max(int(n_walkers), 64)
    if components["PL"]:
        broken = comp_params.get("broken_pl", False) is True
        model.components.append(NuclearContinuumComponent(broken=broken, pars=pars["nuclear_continuum"]))
    if components["FE"]:
        model.components.append(FeComponent(pars=pars["fe_forest"]))
    if components["HOST"]:
        model.components.append(HostGalaxyComponent(pars=pars["host_galaxy"]))
    if components["BC"] or components["BPC"]:
        model.components.append(BalmerCombined(pars=pars["balmer_continuum"],
                                               BalmerContinuum=components["BC"],
                                               BalmerPseudocContinuum=components["BPC"]))
    if any(components[k] for k in ACCEPTED_COMPS[5:]):
        model.components.append(Extinction(MW=components["MW_EXT"], AGN=components["AGN_EXT"],
                                           LMC=components["LMC_EXT"], SMC=components["SMC_EXT"],
                                           Calzetti=components["CALZETTI_EXT"]))
    model.data_spectrum = spectrum

    model.run_mcmc(n_walkers=n_walkers, n_iterations=n_iterations, seed=seed)
    print("Mean acceptance fraction: {0:.3f}".format(np.mean(model.sampler.acceptance_fraction)))

    p_data = {"model": model, "comp_params": comp_params}
    from .distributed import world_info
    if save and world_info()[0] == 0:          # under torchrun every rank holds the same chain: rank 0 writes it
        now = datetime.datetime.now().strftime("%Y%m%d_%M%S")
        if picklefile is None:
            picklefile = "model_{0}.pickle.gz".format(now)
        else:
            picklefile = os.path.basename(picklefile)
            if not picklefile.endswith(".gz"):
                picklefile += ".gz" if picklefile.endswith((".pickle", ".p")) else ".pickle.gz"
        outdir = now if outdir is None else outdir
        os.makedirs(outdir, exist_ok=True)
        pname = os.path.join(outdir, picklefile)
        with gzip.open(pname, "wb") as model_output:
            model_output.write(pickle.dumps(p_data))
        print("Saved pickle file {0}".format(pname))
        p_data["pickle"] = pname
    print("executed in {}".format(datetime.datetime.now() - t1))
    return p_data


def parse_comps(argcomp):
    if len(argcomp) == 1 and "," in argcomp[0]:
        return [x for x in argcomp[0].split(",")]
    return argcomp
