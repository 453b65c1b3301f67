"""AGN nuclear continuum: power law or broken power law.
Reference: spamm/components/NuclearContinuumComponent.py."""
import numpy as np

from .ComponentBase import Component
from .. import _lib
from ..utils.runningmeanfast import runningMeanFast
from ..utils.parse_pars import parse_pars


class NuclearContinuumComponent(Component):
    """F_lambda = norm_PL (lambda/lambda_0)^(-slope1), lambda_0 = median wavelength of the
    data; broken: norm_PL (lambda/wave_break)^(-slope1|slope2) (NC:176-201)."""
    kind = _lib.COMP_NUCLEAR

    def __init__(self, pars=None, broken=None):
        super().__init__()
        self.name = "Nuclear"
        self.inputpars = parse_pars()["nuclear_continuum"] if pars is None else pars
        self.broken_pl = self.inputpars["broken_pl"] if broken is None else broken
        if not self.broken_pl:
            self.model_parameter_names = ["norm_PL", "slope1"]
        else:
            self.model_parameter_names = ["wave_break", "norm_PL", "slope1", "slope2"]
        self.norm_min = self.inputpars["pl_norm_min"]
        self.norm_max = self.inputpars["pl_norm_max"]
        self.slope_min = self.inputpars["pl_slope_min"]
        self.slope_max = self.inputpars["pl_slope_max"]
        self.wave_break_min = self.inputpars["pl_wave_break_min"]
        self.wave_break_max = self.inputpars["pl_wave_break_max"]

    @property
    def is_analytic(self):
        return True

    def _resolve(self, spectrum):
        """String sentinels -> numbers (side effect of initial_values, NC:98-109)."""
        if self.norm_max == "max_flux":
            self.norm_max = max(runningMeanFast(spectrum.flux, self.inputpars["boxcar_width"]))
        elif self.norm_max == "fnw":
            self.norm_max = spectrum.norm_wavelength_flux
        if self.broken_pl:
            if self.wave_break_min == "min_wl":
                self.wave_break_min = min(spectrum.spectral_axis)
            if self.wave_break_max == "max_wl":
                self.wave_break_max = max(spectrum.spectral_axis)

    def initial_values(self, spectrum):
        """Uniform draws from the priors, in the reference's RNG call order (NC:82-124)."""
        self._resolve(spectrum)
        pl_init = []
        if self.broken_pl:
            size = 2
            pl_init.append(np.random.uniform(low=self.wave_break_min, high=self.wave_break_max))
        else:
            size = 1
        pl_init.append(np.random.uniform(self.norm_min, high=self.norm_max))
        for slope in np.random.uniform(low=self.slope_min, high=self.slope_max, size=size):
            pl_init.append(slope)
        return pl_init

    def ln_priors(self, params):
        """Flat boxes with strict inequalities (NC:129-172)."""
        def box(lo, v, hi):
            return 0. if lo < v < hi else -np.inf
        out = []
        if self.broken_pl:
            out.append(box(self.wave_break_min, params[self.parameter_index("wave_break")],
                           self.wave_break_max))
        out.append(box(self.norm_min, params[self.parameter_index("norm_PL")], self.norm_max))
        out.append(box(self.slope_min, params[self.parameter_index("slope1")], self.slope_max))
        if self.broken_pl:
            out.append(box(self.slope_min, params[self.parameter_index("slope2")], self.slope_max))
        return out

    def prior_bounds(self, spectrum=None):
        if spectrum is not None:
            self._resolve(spectrum)
        if self.broken_pl:
            return ([self.wave_break_min, self.norm_min, self.slope_min, self.slope_min],
                    [self.wave_break_max, self.norm_max, self.slope_max, self.slope_max])
        return [self.norm_min, self.slope_min], [self.norm_max, self.slope_max]

    def plan_fill(self, desc, keep, spectrum):
        desc.nc_broken_pl = 1 if self.broken_pl else 0
        desc.nc_norm_wl = float(spectrum.norm_wavelength)

    def flux(self, spectrum, parameters=None):
        assert len(parameters) == len(self.model_parameter_names), \
            "The wrong number of indices were provided: {0}".format(parameters)
        return super().flux(spectrum, parameters)
