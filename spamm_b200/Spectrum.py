"""Data container with the reference's Spectrum surface (spamm/Spectrum.py:19-128)
minus the specutils base class: unit-less ndarrays plus unit labels."""
import numpy as np

from .utils.parse_pars import parse_pars

_GLOBAL = parse_pars()["global"]
FLUX_UNIT = _GLOBAL["flux_unit"]
WL_UNIT = _GLOBAL["wl_unit"]


def _strip(x):
    """Accept astropy-like Quantities without importing astropy."""
    unit = getattr(x, "unit", None)
    val = getattr(x, "value", None)
    if unit is not None and val is not None:
        return np.asarray(val), unit
    arr = getattr(x, "array", None)          # NDUncertainty-like
    if arr is not None and not isinstance(x, np.ndarray):
        return np.asarray(arr), getattr(x, "unit", None)
    return x, None


class Spectrum(object):
    """
    Args:
        spectral_axis (array-like): Wavelength values.
        flux (array-like): Flux values.
        flux_error (array-like): Error on `flux` values.
        spectral_axis_unit, flux_unit: unit labels (strings).
    """

    def __init__(self, spectral_axis, flux, flux_error, spectral_axis_unit=WL_UNIT,
                 flux_unit=FLUX_UNIT, *args, **kwargs):
        spectral_axis, u = _strip(spectral_axis)
        if u is not None:
            spectral_axis_unit = u
        flux, u = _strip(flux)
        if u is not None:
            flux_unit = u
        flux_error, u = _strip(flux_error)
        if u is not None:
            assert u == flux_unit, "Flux and flux error units must match"
        assert len(spectral_axis) == len(flux), "Wavelength and flux arrays must be the same length"
        assert len(flux) == len(flux_error), "Flux and flux error arrays must be the same length"
        self.flux_unit = flux_unit
        self.flux_error = flux_error
        self._flux = flux
        self._norm_wavelength = None
        self._norm_wavelength_flux = None
        self._spectral_axis = spectral_axis
        self._spectral_axis_unit = spectral_axis_unit
        self.uncertainty = flux_error

    @property
    def norm_wavelength(self):
        """Median wavelength (Spectrum.py:84-87)."""
        if self._norm_wavelength is None:
            self._norm_wavelength = np.median(self.spectral_axis)
        return self._norm_wavelength

    @property
    def norm_wavelength_flux(self):
        """Flux at the normalization wavelength, linear interpolation (Spectrum.py:90-95)."""
        if self._norm_wavelength_flux is None:
            self._norm_wavelength_flux = np.interp(self.norm_wavelength, self.spectral_axis, self.flux)
        return self._norm_wavelength_flux

    def grid_spacing(self):
        """Spacing of the wavelength grid; uniform grids only (Spectrum.py:97-99)."""
        return self.spectral_axis[1] - self.spectral_axis[0]

    @property
    def spectral_axis(self):
        return self._spectral_axis

    @spectral_axis.setter
    def spectral_axis(self, new_wl):
        self._spectral_axis = new_wl
        self._norm_wavelength = None

    @property
    def flux(self):
        return self._flux

    @flux.setter
    def flux(self, new_flux):
        self._flux = new_flux
        self._norm_wavelength_flux = None

    def copy(self):
        return Spectrum(np.array(self.spectral_axis), np.array(self.flux), np.array(self.flux_error),
                        self._spectral_axis_unit, self.flux_unit)
