"""Shared machinery of the two template components (Fe forest, host galaxy):
N templates, Gaussian-broadened in ln-lambda space by a per-walker width and
scaled (reference: FeComponent.py / HostGalaxyComponent.py are copy-paste
twins).  The per-call arithmetic -- kernel, convolution, interpolation onto the
data grid, normalisation -- runs on the GPU (csrc/template_kernels.cuh: k_conv_rows,
k_interp_rows, k_norm_rows, k_walkers); this host class only prepares the
ln-lambda grids once, as ``initialize`` does in the reference."""
import numpy as np

from .ComponentBase import Component
from .. import _lib, constants
from ..Spectrum import Spectrum
from ..utils.runningmeanfast import runningMeanFast


class TemplateComponent(Component):
    # set by subclasses
    _prefix = None            # parameter-name prefix
    _width_name = None
    _clamp_edges = False
    _sort_templates = True

    def _load(self, location, family):
        from ..data import load_template_dir
        raw = load_template_dir(location, family, sort=self._sort_templates)
        assert len(raw) != 0, "No templates found in specified directory {0}".format(location)
        assert len(raw) <= _lib.MAX_TEMPLATES, "too many templates"
        templ = []
        for wavelengths, flux in raw:
            flux = np.where(flux < 0, 1e-19, flux)                 # Fe:88 / HG:102
            templ.append(Spectrum(spectral_axis=wavelengths, flux=flux, flux_error=flux))
        return templ

    def _log_rebin(self, templates):
        """Uniform ln-lambda grid with as many points as the template (Fe:173-182 / HG:173-181)."""
        out = []
        for template in templates:
            ln = np.log(template.spectral_axis)
            log_wl = np.linspace(min(ln), max(ln), num=len(template.spectral_axis))
            if not self.fast_interp:
                # rebin_spec: True (Fe:184-186 / HG:183-185): flux-conserving rebin, utils/rebin_spec.py
                from ..utils.rebin_spec import rebin_spec
                log_flux = rebin_spec(ln, template.flux, log_wl)
            elif self._clamp_edges:
                log_flux = np.interp(log_wl, ln, template.flux)
            else:
                log_flux = np.interp(log_wl, ln, template.flux, left=0, right=0)
            out.append(Spectrum(spectral_axis=log_wl, flux=log_flux, flux_error=log_flux))
        return out

    def _resolve_norm_max(self, spectrum):
        if self.norm_max == "max_flux":
            self.norm_max = max(runningMeanFast(spectrum.flux, self.inputpars["boxcar_width"]))
        elif self.norm_max == "fnw":
            self.norm_max = spectrum.norm_wavelength_flux

    def _template_set(self, templates, log_templates, template_width, ksize, keep):
        ts = _lib.SpammTemplateSet()
        ts.n_templates = len(templates)
        xs, fs = [], []
        for i, (t, lt) in enumerate(zip(templates, log_templates)):
            ts.n_points[i] = len(lt.spectral_axis)
            ts.norm_wl[i] = float(np.median(t.spectral_axis))        # Fe:325 / HG:330
            xs.append(np.asarray(lt.spectral_axis, dtype=np.float64))
            fs.append(np.asarray(lt.flux, dtype=np.float64))
        x, xp = _lib.as_c(np.concatenate(xs))
        f, fp = _lib.as_c(np.concatenate(fs))
        keep += [x, f]
        ts.log_wl, ts.log_flux = xp, fp
        ts.template_width = float(template_width)
        ts.sigma_denom = float(constants.FWHM_TO_SIGMA_DENOM)
        ts.sqrt_2pi = constants.SQRT_2PI
        ts.kernel_size_sigma = int(ksize)
        ts.clamp_edges = 1 if self._clamp_edges else 0
        return ts

    def _rebin_operators(self, ts, log_templates, spectrum, keep):
        """rebin_spec: True -- the per-call rebin of the convolved row onto ln(data wavelengths)
        (Fe:320-323 / HG:326-329) as one CSR matrix per template, applied on the device."""
        if self.fast_interp:
            return
        from ..utils.rebin_spec import rebin_matrix
        ln_data = np.log(np.asarray(spectrum.spectral_axis, dtype=np.float64))
        ptrs, idxs, ws, base = [], [], [], 0
        for lt in log_templates:
            ip, ix, w = rebin_matrix(np.asarray(lt.spectral_axis, dtype=np.float64), ln_data)
            ptrs.append(ip + base)
            idxs.append(ix)
            ws.append(w)
            base += len(ix)
        ip, ipp = _lib.as_c_i32(np.concatenate(ptrs))
        ix, ixp = _lib.as_c_i32(np.concatenate(idxs))
        w, wp = _lib.as_c(np.concatenate(ws))
        keep += [ip, ix, w]
        ts.rebin_indptr, ts.rebin_index, ts.rebin_weight = ipp, ixp, wp
