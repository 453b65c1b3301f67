"""Model + ln-posterior with the reference's surface (spamm/Model.py): the
`components` list fixes the parameter layout, `data_spectrum = ...` initialises
the components, and `ln_posterior(new_params, model)` is what a sampler calls.

Underneath, every flux / likelihood / ln-posterior evaluation is one batched
CUDA call through the C ABI (include/spamm_b200.h).  New, batched entry points:
`Model.lnposterior_batch(params[W, D])`, `Model.model_flux_batch`, and
`GpuPool` (spamm_b200/sampler.py) for an unmodified emcee."""
import sys

import numpy as np

from .Spectrum import Spectrum
from . import _lib

iteration_count = 0


class MCMCDidNotConverge(Exception):
    pass


def ln_posterior(new_params, *args):
    """ln(prior) + ln(likelihood) of one parameter vector; -inf outside the
    priors (Model.py:42-79).  args[0] is the Model."""
    global iteration_count
    iteration_count = iteration_count + 1
    model = args[0]
    return float(model.lnposterior_batch(np.asarray(new_params, dtype=np.float64)[None, :])[0])


class Model(object):
    """
    Attributes:
        components (list): component plugins; list order = parameter order.
        data_spectrum (Spectrum): setting it initialises every component.
        model_spectrum (Spectrum): wavelength grid + last model flux.
        sampler: the ensemble sampler after run_mcmc.
    """

    def __init__(self, wavelength_start=1000, wavelength_end=10000, wavelength_delta=0.05, mpi=False):
        self._mask = None
        self._data_spectrum = None
        self.z = None
        self.components = []
        self.mpi = mpi
        self.sampler = None
        # the data spectrum defines the model grid (Model.py:189); no 180k-point
        # placeholder grid is allocated up front
        self._init_grid = (wavelength_start, wavelength_end, wavelength_delta)
        self.model_spectrum = Spectrum(spectral_axis=np.zeros(0), flux=np.zeros(0), flux_error=np.zeros(0))
        self.downsample_data_if_needed = False
        self.upsample_components_if_needed = False
        self.print_parameters = False
        self.max_walkers = 8192
        self.template_mode = _lib.TEMPLATES_AUTO
        self._plan = None

    # -- data ----------------------------------------------------------------------
    @property
    def data_spectrum(self):
        """All components of the model must be set before setting the data."""
        return self._data_spectrum

    @data_spectrum.setter
    def data_spectrum(self, new_data_spectrum):
        self._data_spectrum = new_data_spectrum
        if len(self.components) == 0:
            raise Exception("Components must be added before defining the data spectrum.")
        self.model_spectrum.spectral_axis = np.array(new_data_spectrum.spectral_axis)
        self.model_spectrum.flux = np.zeros(len(self.model_spectrum.spectral_axis))
        # Components coarser than the data (Model.py:196-240).  The shipped components report no grid
        # spacing (App. B Q9: the template components' `is_analytic` is a method, so `grid_spacing()` is
        # None), hence the model grid is the data grid; a plugin component that does report a spacing
        # coarser than the data takes the branches below.
        gs = 0
        worst_component = None
        for component in self.components:
            component.initialize(data_spectrum=new_data_spectrum)
            if component.grid_spacing() and component.grid_spacing() > gs:
                gs = component.grid_spacing()
                worst_component = component
        self.worst_component = worst_component
        if gs > new_data_spectrum.grid_spacing():
            if self.upsample_components_if_needed:
                pass            # components are evaluated on the data grid anyway
            elif self.downsample_data_if_needed:
                # The data are resampled onto the coarsest component's spacing and everything -- model grid,
                # components, likelihood -- moves to that grid.  (The reference's lines 219-231 index the
                # Spectrum object and store an unevaluated interp1d, so they raise; this is what they set
                # out to do, with the errors resampled linearly like the flux.)
                wl = np.asarray(new_data_spectrum.spectral_axis, dtype=np.float64)
                new_wl = np.arange(wl[0], wl[-1], gs)
                downsampled = Spectrum(spectral_axis=new_wl,
                                       flux=np.interp(new_wl, wl, np.asarray(new_data_spectrum.flux)),
                                       flux_error=np.interp(new_wl, wl, np.asarray(new_data_spectrum.flux_error)))
                self._data_spectrum = downsampled
                self.model_spectrum.spectral_axis = np.array(new_wl)
                self.model_spectrum.flux = np.zeros(len(new_wl))
                for component in self.components:
                    component.initialize(data_spectrum=downsampled)
            # else: the reference's `assert True, "... courser wavelength grid spacing ..."` never fires
        self._plan = None

    # -- plan ------------------------------------------------------------------------
    @property
    def plan(self):
        """Device plan, built on first use (after the prior sentinels can be resolved)."""
        if self._plan is None:
            if self._data_spectrum is None:
                raise Exception("Set data_spectrum before evaluating the model.")
            from .plan import ModelPlan
            self._plan = ModelPlan(self.components, self._data_spectrum, max_walkers=self.max_walkers,
                                   template_mode=self.template_mode)
        return self._plan

    def invalidate_plan(self):
        self._plan = None

    def __getstate__(self):
        """Pickle without the device plan (rebuilt on demand after loading)."""
        d = dict(self.__dict__)
        d["_plan"] = None
        d.pop("_exchange", None)
        return d

    # -- sampler -----------------------------------------------------------------------
    def initial_walkers(self, n_walkers):
        """Walker matrix drawn from the priors, component by component (Model.py:254-259)."""
        walkers_matrix = []
        for walker in range(n_walkers):
            walker_params = []
            for component in self.components:
                walker_params = walker_params + component.initial_values(self.data_spectrum)
            walkers_matrix.append(walker_params)
        return walkers_matrix

    def run_mcmc(self, n_walkers=100, n_iterations=100, seed=None):
        """Run the ensemble sampler (emcee 2.2.1 stretch move, on the device)."""
        from .sampler import EnsembleSampler
        walkers_matrix = self.initial_walkers(n_walkers)
        global iteration_count
        iteration_count = 0
        self.sampler = EnsembleSampler(nwalkers=n_walkers, dim=len(walkers_matrix[0]),
                                       lnpostfn=ln_posterior, args=[self], seed=seed)
        self.sampler.run_mcmc(walkers_matrix, n_iterations)

    # -- bookkeeping ------------------------------------------------------------------
    @property
    def total_parameter_count(self):
        return sum(c.parameter_count for c in self.components)

    def model_parameter_names(self):
        labels = []
        for c in self.components:
            labels = labels + [x for x in c.model_parameter_names]
        return labels

    # -- evaluation ------------------------------------------------------------------------
    def model_flux_batch(self, params):
        """[W, D] -> [W, P] summed component fluxes (Model.model_flux, Model.py:328-366)."""
        return self.plan.model_flux_host(params)

    def model_flux(self, params):
        if self.print_parameters:
            print("params = {0}".format(params))
        self.model_spectrum.flux = self.model_flux_batch(np.asarray(params, dtype=np.float64)[None, :])[0]
        return self.model_spectrum.flux

    def add_component(self, component, parameters):
        """Add one component's flux to `model_spectrum.flux` (Model.py:368-380); the step-wise form of
        `model_flux`, kept for callers that build the model spectrum themselves."""
        component_flux = component.flux(spectrum=self.data_spectrum, parameters=parameters)
        self.model_spectrum.flux = self.model_spectrum.flux + component_flux

    def reddening(self, component, parameters):
        """Multiply `model_spectrum.flux` by the component's extinction curve (Model.py:384-395)."""
        extinction = component.extinction(spectrum=self.data_spectrum, params=parameters)
        self.model_spectrum.flux = np.array(self.model_spectrum.flux) * extinction

    def likelihood(self, model_spectrum_flux):
        """ln L of a model flux on the data grid (Model.py:418-445)."""
        import torch
        m = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(model_spectrum_flux), dtype=np.float64)).cuda()
        return float(self.plan.likelihood(m).cpu().numpy()[0])

    def prior(self, params):
        """Sum of the components' ln-priors (Model.py:449-470)."""
        p = np.copy(params)
        ln_p = 0
        for component in self.components:
            ln_p += sum(component.ln_priors(params=p[0:component.parameter_count]))
            p = p[component.parameter_count:]
        return ln_p

    def lnposterior_batch_sharded(self, params, out=None):
        """`lnposterior_batch` with the ensemble sharded over the ranks of torch.distributed (one process per
        GPU): `params` is the same [W, D] array on every rank, each rank uploads and evaluates only its
        slice, the walker kernel stores the results into every rank's exchange region over NVLink, and every
        rank returns the whole [W] vector.  Collective: all ranks must call it together.  (This is what
        replaces emcee's pool.map task farm, Model.py:264-289.)"""
        from .distributed import world_info
        if world_info()[1] == 1:
            return self.plan.lnposterior_host(params, out=out)
        if getattr(self, "_exchange", None) is None:
            from .distributed import Exchange
            self._exchange = Exchange(self.max_walkers)
        return self.plan.lnposterior_sharded_host(params, self._exchange, out=out)

    def lnposterior_batch(self, params, out=None):
        """ln-posterior of a whole walker ensemble in one call: numpy [W, D] -> numpy [W]
        (`out`: optional preallocated result array; page-locked buffers avoid a staging copy)."""
        return self.plan.lnposterior_host(params, out=out)
