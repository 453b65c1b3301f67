"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (spamm_b200/).

Shim loader that imports the *unmodified* reference package from
``/root/reference`` in a container that lacks its third-party dependencies
(astropy 3.x, specutils, emcee 2.2.1, pysynphot, pyfftw, matplotlib) and has a
scipy that dropped ``signal.gaussian`` / ``integrate.simps`` / callable
``interp2d``.  Only the *third-party* arithmetic is restated here; every line
of ``spamm/Model.py`` and ``spamm/components/*.py`` runs as shipped.

The reference tree only exists in the build container, so this module is used
(a) by ``tests/golden/make_golden.py`` to generate the committed fixtures and
(b) by not-gpu tests that are skipped when ``/root/reference`` is absent.

Third-party pieces restated (call sites in the reference):
  * astropy.modeling.powerlaws.PowerLaw1D / BrokenPowerLaw1D
        spamm/components/NuclearContinuumComponent.py:5,195,199
  * astropy.modeling.blackbody.blackbody_lambda
        spamm/components/BalmerContinuumCombined.py:11,435
  * astropy.constants c, h, k_B, Ryd (astropy 3.x => CODATA-2014)
        spamm/components/BalmerContinuumCombined.py:10,21-24
        spamm/components/FeComponent.py:10,270
        spamm/components/HostGalaxyComponent.py:13,275
  * scipy.signal.gaussian       FeComponent.py:292, HostGalaxyComponent.py:299
  * pickled scipy interp2d      BalmerContinuumCombined.py:288-290
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SPAMM_REFERENCE_ROOT", "/root/reference")

# CODATA-2014 (astropy 3.x default)
C_MS = 299792458.0
H_SI = 6.626070040e-34
KB_SI = 1.38064852e-23
RYD_M = 10973731.568508


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "spamm"))


class _Const(object):
    """Minimal stand-in for an astropy Constant/Quantity: .value, .cgs, .to()."""

    def __init__(self, value, conv):
        self.value = value
        self._conv = conv  # dict target -> value

    @property
    def cgs(self):
        return _Const(self._conv["cgs"], self._conv)

    def to(self, unit):
        return _Const(self._conv[str(unit)], self._conv)


class _Unit(object):
    """`array * Unit` and `Unit * array` hand back the bare ndarray."""
    __array_priority__ = 1e6

    def __init__(self, s=None):
        self.s = s

    def __mul__(self, other):
        return np.asarray(other)

    __rmul__ = __mul__

    def __eq__(self, other):
        return isinstance(other, _Unit) and self.s == other.s

    def __hash__(self):
        return hash(self.s)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _blackbody_lambda(in_x, temperature):
    """astropy 3.x blackbody_lambda for in_x in Angstrom:
    B_nu = 2 h nu^3 / (c^2 expm1(h nu / k T)) [cgs], B_lambda = B_nu * c_AA/lambda^2
    (freq = c/lambda via the spectral equivalency; only the shape matters
    downstream because BC_flux renormalises, BalmerContinuumCombined.py:444-446)."""
    lam = np.asarray(in_x, dtype=np.float64)
    h = H_SI * 1e7            # erg s
    kB = KB_SI * 1e7          # erg / K
    c_cgs = C_MS * 100.0      # cm / s
    c_AA = C_MS * 1e10        # AA / s
    freq = c_AA / lam
    log_boltz = h * freq / (kB * temperature)
    boltzm1 = np.expm1(log_boltz)
    bnu = 2.0 * h * freq ** 3 / (c_cgs ** 2 * boltzm1)
    return bnu * c_AA / lam ** 2


class _PowerLaw1D(object):
    def __init__(self, amplitude, x_0, alpha):
        self.amplitude, self.x_0, self.alpha = amplitude, x_0, alpha

    def __call__(self, x):
        xx = np.asarray(x, dtype=np.float64) / self.x_0
        return self.amplitude * xx ** (-self.alpha)


class _BrokenPowerLaw1D(object):
    def __init__(self, amplitude, x_break, alpha_1, alpha_2):
        self.amplitude, self.x_break = amplitude, x_break
        self.alpha_1, self.alpha_2 = alpha_1, alpha_2

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        alpha = np.where(x < self.x_break, self.alpha_1, self.alpha_2)
        xx = x / self.x_break
        return self.amplitude * xx ** (-alpha)


_installed = False


def install():
    """Insert the stub modules and make `import spamm` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    import scipy.signal
    import scipy.signal.windows
    import scipy.integrate
    import scipy.interpolate
    from scipy.interpolate import bisplev
    import yaml

    class Quantity(object):
        pass

    units = _mod("astropy.units", Unit=_Unit, Quantity=Quantity, AA=_Unit("AA"))
    _mod("astropy.units.quantity", Quantity=Quantity)

    class NDUncertainty(object):
        pass

    class StdDevUncertainty(NDUncertainty):
        def __init__(self, array=None, unit=None):
            self.array = array
            self.unit = unit

    nddata = _mod("astropy.nddata", NDUncertainty=NDUncertainty,
                  StdDevUncertainty=StdDevUncertainty)

    conv_c = {"cgs": C_MS * 100.0, "km/s": C_MS / 1000.0}
    conv_h = {"cgs": H_SI * 1e7}
    conv_k = {"cgs": KB_SI * 1e7}
    conv_R = {"cgs": RYD_M / 100.0, "1/Angstrom": RYD_M * 1e-10}
    constants = _mod("astropy.constants",
                     c=_Const(C_MS, conv_c), h=_Const(H_SI, conv_h),
                     k_B=_Const(KB_SI, conv_k), Ryd=_Const(RYD_M, conv_R))

    powerlaws = _mod("astropy.modeling.powerlaws", PowerLaw1D=_PowerLaw1D,
                     BrokenPowerLaw1D=_BrokenPowerLaw1D)
    blackbody = _mod("astropy.modeling.blackbody", blackbody_lambda=_blackbody_lambda)
    modeling = _mod("astropy.modeling", powerlaws=powerlaws, blackbody=blackbody)
    convolution = _mod("astropy.convolution", Gaussian1DKernel=object, convolve=None)
    _mod("astropy", units=units, nddata=nddata, constants=constants,
         modeling=modeling, convolution=convolution)

    class Spectrum1D(object):
        def __init__(self, *a, **k):
            pass

    class WCSWrapper(object):
        @staticmethod
        def from_array(a):
            return None

    wcsw = _mod("specutils.wcs.wcs_wrapper", WCSWrapper=WCSWrapper)
    wcs = _mod("specutils.wcs", wcs_wrapper=wcsw)
    _mod("specutils", Spectrum1D=Spectrum1D, wcs=wcs)

    _mod("pyfftw")
    plt = _mod("matplotlib.pyplot")
    _mod("matplotlib", pyplot=plt)
    _mod("emcee")

    if not hasattr(scipy.signal, "gaussian"):
        scipy.signal.gaussian = lambda M, std, sym=True: \
            scipy.signal.windows.gaussian(int(M), std, sym)
    if not hasattr(scipy.integrate, "simps"):
        scipy.integrate.simps = scipy.integrate.simpson
    try:
        import scipy.fftpack.helper  # noqa: F401
    except Exception:  # pragma: no cover
        import scipy.fft
        _mod("scipy.fftpack.helper", next_fast_len=scipy.fft.next_fast_len)

    _orig_load = yaml.load
    if not getattr(yaml.load, "_spamm_shim", False):
        def _load(stream, Loader=None):
            return _orig_load(stream, Loader=Loader or yaml.SafeLoader)
        _load._spamm_shim = True
        yaml.load = _load

    # old pickles reference scipy.interpolate.interpolate.interp2d; make the
    # class callable again through its stored FITPACK tck
    try:
        from scipy.interpolate import _interpolate as _ip
    except Exception:  # pragma: no cover
        _ip = scipy.interpolate.interpolate
    cls = getattr(_ip, "interp2d", None)
    if cls is None or not isinstance(cls, type):
        class cls(object):  # noqa: N801
            pass
        _ip.interp2d = cls
    cls.__call__ = lambda self, x, y: np.atleast_1d(bisplev(x, y, self.tck))
    if "scipy.interpolate.interpolate" not in sys.modules:
        try:
            import scipy.interpolate.interpolate  # noqa: F401
        except Exception:
            _mod("scipy.interpolate.interpolate", interp2d=cls)
    if not isinstance(getattr(sys.modules["scipy.interpolate.interpolate"],
                              "interp2d", None), type):
        sys.modules["scipy.interpolate.interpolate"].interp2d = cls

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


class chdir_examples(object):
    """The reference resolves template dirs and the SH95 pickle relative to
    the cwd (utils/parameters.yaml:20,66; BalmerContinuumCombined.py:288)."""

    def __enter__(self):
        self._old = os.getcwd()
        os.chdir(os.path.join(REFERENCE_ROOT, "examples"))

    def __exit__(self, *exc):
        os.chdir(self._old)


def load():
    """Return a namespace with the reference's own classes/functions."""
    install()
    with chdir_examples():
        from spamm.Spectrum import Spectrum
        from spamm import Model as model_mod
        from spamm.components.NuclearContinuumComponent import NuclearContinuumComponent
        from spamm.components.FeComponent import FeComponent
        from spamm.components.HostGalaxyComponent import HostGalaxyComponent
        from spamm.components.BalmerContinuumCombined import BalmerCombined
    ns = types.SimpleNamespace(
        Spectrum=Spectrum, Model=model_mod.Model, ln_posterior=model_mod.ln_posterior,
        model_module=model_mod,
        NuclearContinuumComponent=NuclearContinuumComponent, FeComponent=FeComponent,
        HostGalaxyComponent=HostGalaxyComponent, BalmerCombined=BalmerCombined)
    return ns
