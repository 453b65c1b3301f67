"""Plugin contract of a model component -- same surface as the reference's
``Component`` ABC (spamm/components/ComponentBase.py:16-118): ``name``,
``model_parameter_names``, ``initial_values``, ``ln_priors``, ``flux``,
``initialize``, ``parameter_count``, ``parameter_index``, ``is_analytic``,
``fast_interp``.

What is new underneath: ``flux()`` runs on the GPU (one walker = a batched call
with W = 1), ``flux_batch()`` evaluates many parameter vectors in one launch,
and ``plan_fill()`` contributes the component's frozen state to the device plan
(include/spamm_b200.h: SpammPlanDesc)."""
from abc import ABC, abstractmethod

import numpy as np

from ..utils.parse_pars import parse_pars

PARS = parse_pars()


class Component(ABC):
    #: SpammComponentKind of the concrete class
    kind = None

    def __init__(self):
        self.z = None  # redshift
        self.reddening_law = None
        self.data_wavelength_grid = None
        self.interpolated_flux = None
        self._flux_plans = {}

    def parameter_index(self, parameter_name):
        for idx, pname in enumerate(self.model_parameter_names):
            if parameter_name == pname:
                return idx
        return None

    @property
    def is_analytic(self):
        raise NotImplementedError("Please define the 'is_analytic' property for the class "
                                  "'{0}'.".format(type(self).__name__))

    @property
    def parameter_count(self):
        """Number of parameters of this component."""
        if self.z:
            return len(self.model_parameter_names) + 1
        return len(self.model_parameter_names)

    @abstractmethod
    def initial_values(self, spectrum=None):
        """Return type must be a list (not an np.array)."""

    @abstractmethod
    def ln_priors(self, params):
        """Return a list of the ln of all of the priors."""

    def native_wavelength_grid(self):
        return None

    def grid_spacing(self):
        """Analytic components have no grid spacing (ComponentBase.py:86-92)."""
        return None

    def initialize(self, data_spectrum=None):
        pass

    @property
    def fast_interp(self):
        """np.interp instead of the flux-conserving rebin (ComponentBase.py:112-118)."""
        return PARS["rebin_spec"] is False

    # -- device path ---------------------------------------------------------
    def __getstate__(self):
        """Pickle without the cached device plans of flux() (ctypes handles; rebuilt on demand) -- the
        reference's components stay picklable after any call (run_spamm.py:157-160 pickles the Model)."""
        d = dict(self.__dict__)
        if "_flux_plans" in d:
            d["_flux_plans"] = {}
        return d

    @abstractmethod
    def prior_bounds(self, spectrum=None):
        """(lo, hi) lists of the flat-box priors, resolving the string
        sentinels against `spectrum` exactly like initial_values does."""

    @abstractmethod
    def plan_fill(self, desc, keep, spectrum):
        """Write this component's frozen state into a SpammPlanDesc."""

    def _single_plan(self, spectrum):
        from ..plan import ModelPlan
        wl = np.asarray(spectrum.spectral_axis, dtype=np.float64)
        key = (len(wl), float(wl[0]), float(wl[-1]), hash(wl.tobytes()))
        plan = self._flux_plans.get(key)
        if plan is None:
            if len(self._flux_plans) > 4:
                self._flux_plans.clear()
            plan = ModelPlan([self], spectrum, with_data=False, with_priors=False, max_walkers=64)
            self._flux_plans[key] = plan
        return plan

    def flux_batch(self, spectrum, parameters):
        """flux() for many parameter vectors at once: [W, k] -> ndarray [W, P]."""
        p = np.atleast_2d(np.asarray(parameters, dtype=np.float64))
        assert p.shape[1] == self.parameter_count, \
            "The wrong number of indices were provided: {0}".format(parameters)
        return self._single_plan(spectrum).model_flux_host(p)

    def flux(self, spectrum=None, parameters=None):
        """Flux of this component on the spectrum's wavelength grid (GPU)."""
        return self.flux_batch(spectrum, np.asarray(parameters, dtype=np.float64)[None, :])[0]
