"""Extinction: a multiplicative reddening component (spamm/components/ReddeningLaw.py:126-216).

The model flux summed so far is multiplied by pow(10, -0.4 * E(B-V) * k(lambda)) (Model.py:358-361,
384-395); k(lambda) is one of the reference's five curves -- Calzetti (:7-20), LMC / Fitzpatrick 1986
(:22-46), MW / Seaton 1979 (:48-72), SMC / Gordon 2003 (:74-98), AGN / Gaskell & Benker 2007
(:100-123).  k is walker independent: the host evaluates it once per data grid in the reference's
operation order and the fused kernel applies it per walker on the device."""
import numpy as np

from .ComponentBase import Component
from .. import _lib


def _calzetti_k(wl):
    k = np.zeros(len(wl))                       # ext = 1 outside 0.12-2.2 um  <=>  k = 0
    for j in range(len(wl)):
        um = wl[j] / 10000.
        if (um >= 0.12) & (um < 0.63):
            k[j] = 2.659 * (-1.857 + 1.040 / um) + 4.05
        if (um >= 0.63) & (um <= 2.2):
            k[j] = 2.659 * (-2.156 + 1.509 / um - 0.198 / pow(um, 2) + 0.11 / pow(um, 3)) + 4.05
    return k


def _fm_k(wl, C1, C2, C3, C4uv, x_0, gamma, Rv):
    """Fitzpatrick-Massa form shared by the LMC, MW and SMC curves."""
    k = np.zeros(len(wl))
    for j in range(len(wl)):
        um = wl[j] / 10000.
        x = pow(um, -1)
        x2 = pow(x, 2)
        D = x2 / (pow(x2 - pow(x_0, 2), 2) + x2 * pow(gamma, 2))
        F = 0.5392 * pow((x - 5.9), 2) + 0.05644 * pow((x - 5.9), 3)
        C4 = C4uv if x >= 5.9 else 0.0
        k[j] = C1 + Rv + C2 * x + C3 * x * D + C4 * F
    return k


def _agn_k(wl):
    A, B, C, D, E, F, Rv = 0.000843, -0.02496, 0.2919, -1.815, 6.83, -7.92, 5.0
    x = pow(np.array(wl / 10000.), -1)          # evaluated on the whole array (:118-122)
    return A * x ** 5 + B * x ** 4 + C * x ** 3 + D * x ** 2 + E * x + F + Rv


class Extinction(Component):
    """One parameter, E(B-V), flat prior on (0, 2) (:155-185).  When several laws are switched on the
    last one in the order MW, AGN, LMC, SMC, Calzetti wins, as in the reference (:193-202); with none
    the extinction is 1 everywhere (:203-204)."""
    kind = _lib.COMP_EXTINCTION

    def __init__(self, MW=False, AGN=False, LMC=False, SMC=False, Calzetti=False):
        super().__init__()
        self.model_parameter_names = ["E(B-V)"]
        self.EBV_min = None
        self.EBV_max = None
        self.name = "Extinction"
        self.MW, self.AGN, self.LMC, self.SMC, self.Calzetti = MW, AGN, LMC, SMC, Calzetti
        self._k = None

    @property
    def is_analytic(self):
        return True

    def _resolve(self):
        if self.EBV_min is None or self.EBV_max is None:
            self.EBV_min = 0
            self.EBV_max = 2.

    def initial_values(self, spectrum=None):
        self._resolve()
        return [np.random.uniform(low=self.EBV_min, high=self.EBV_max)]

    def ln_priors(self, params):
        EBV = params[self.parameter_index("E(B-V)")]
        return [0 if self.EBV_min < EBV < self.EBV_max else -np.inf]

    def k_lambda(self, spectrum):
        """k(lambda) of the selected law on the spectrum's grid."""
        wl = np.asarray(spectrum.spectral_axis, dtype=np.float64)
        k = np.zeros(len(wl))
        if self.MW:
            k = _fm_k(wl, -0.38, 0.74, 3.96, 0.26, 4.595, 1.051, 3.1)
        if self.AGN:
            k = _agn_k(wl)
        if self.LMC:
            k = _fm_k(wl, -0.69, 0.89, 2.55, 0.50, 4.608, 0.994, 3.1)
        if self.SMC:                             # far-UV C4 = 0.26 inside the loop (:92), not 0.46
            k = _fm_k(wl, -4.96, 2.26, 0.39, 0.26, 4.6, 1.0, 2.74)
        if self.Calzetti:
            k = _calzetti_k(wl)
        return k

    def prior_bounds(self, spectrum=None):
        self._resolve()
        return [self.EBV_min], [self.EBV_max]

    def plan_fill(self, desc, keep, spectrum):
        k, kp = _lib.as_c(self.k_lambda(spectrum))
        keep.append(k)
        desc.ext_k = kp

    def flux(self, spectrum=None, parameters=None):
        """The component adds no flux of its own (:180-185)."""
        return [0.] * len(spectrum.spectral_axis)

    def extinction(self, spectrum=None, params=None):
        """Extinction curve for E(B-V) = params[0] on the spectrum's grid, evaluated on the GPU."""
        if spectrum is None:
            raise Exception("Need a data spectrum")
        p = np.asarray([params[self.parameter_index("E(B-V)")]], dtype=np.float64)[None, :]
        return self._single_plan(spectrum).model_flux_host(p)[0]
