"""ModelPlan: the frozen, device-resident snapshot of a model (grids, data,
templates on their ln-lambda grids, SH95 table, prior boxes, flags) behind the
C ABI of include/spamm_b200.h.  PyTorch owns every params/output buffer; the
plan owns only its immutable device constants."""
import ctypes

import numpy as np

from . import _lib


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.SpammError("spamm_b200 needs a CUDA device (no CPU fallback)")
    return torch


class ModelPlan(object):
    def __init__(self, components, spectrum, with_data=True, with_priors=True, max_walkers=8192,
                 template_mode=_lib.TEMPLATES_AUTO, bounds=None):
        """components: list in Model.components order; spectrum: the data Spectrum."""
        self.lib = _lib.load()
        _torch()
        keep = []
        desc = _lib.SpammPlanDesc()
        desc.abi_version = _lib.ABI_VERSION
        wl, wlp = _lib.as_c(spectrum.spectral_axis)
        lnwl, lnwlp = _lib.as_c(np.log(wl))
        keep += [wl, lnwl]
        self.n_pix = len(wl)
        desc.n_pix = self.n_pix
        desc.wl, desc.ln_wl = wlp, lnwlp
        if with_data:
            fx, fxp = _lib.as_c(spectrum.flux)
            er, erp = _lib.as_c(spectrum.flux_error)
            keep += [fx, er]
            desc.flux, desc.flux_error = fxp, erp
            with np.errstate(divide="ignore", invalid="ignore"):
                # walker-independent half of Model.likelihood (Model.py:440)
                desc.sum_ln_2pi_err2 = float(np.sum(np.log(2 * np.pi * np.power(er, 2))))
        if len(components) == 0:
            raise Exception("Components must be added before defining the data spectrum.")
        if len(components) > _lib.MAX_COMPONENTS:
            raise ValueError("too many components")
        desc.n_components = len(components)
        lo, hi = [], []
        self.offsets = []
        for i, c in enumerate(components):
            if c.kind is None:
                raise ValueError("component %r has no GPU kind" % (c,))
            desc.component_kind[i] = c.kind
            self.offsets.append(len(lo))
            l, h = c.prior_bounds(spectrum if with_priors else None)
            lo += list(l)
            hi += list(h)
            c.plan_fill(desc, keep, spectrum)
        self.n_dim = len(lo)
        desc.n_dim = self.n_dim
        if bounds is not None:
            lo, hi = list(bounds[0]), list(bounds[1])
            assert len(lo) == self.n_dim and len(hi) == self.n_dim
        if with_priors or bounds is not None:
            for v in lo + hi:
                if isinstance(v, str):
                    raise TypeError("prior bound %r is unresolved; call initial_values(spectrum) "
                                    "first or pass the data spectrum" % (v,))
            self.lo, lop = _lib.as_c(lo)
            self.hi, hip = _lib.as_c(hi)
            keep += [self.lo, self.hi]
            desc.prior_lo, desc.prior_hi = lop, hip
        else:
            self.lo = self.hi = None
        desc.template_mode = template_mode
        desc.max_walkers = int(max_walkers)
        self.max_walkers = int(max_walkers)
        self.n_components = len(components)
        handle = ctypes.c_void_p()
        _lib.check(self.lib.spamm_plan_create(ctypes.byref(desc), ctypes.byref(handle)))
        self._h = handle
        self.template_mode = self.lib.spamm_plan_template_mode(self._h)
        self.bank_rows = self.lib.spamm_plan_bank_rows(self._h)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self.lib.spamm_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- helpers ---------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream().cuda_stream)

    def _check_params(self, params):
        torch = _torch()
        if not (isinstance(params, torch.Tensor) and params.is_cuda and params.dtype == torch.float64
                and params.dim() == 2 and params.shape[1] == self.n_dim and params.is_contiguous()):
            raise ValueError("params must be a contiguous CUDA float64 tensor of shape [W, %d]" % self.n_dim)

    @property
    def launch_count(self):
        return int(self.lib.spamm_plan_launch_count(self._h))

    @property
    def specialised(self):
        """True if ln-posterior calls run the specialised (FULL) instantiation of the fused kernel."""
        return bool(self.lib.spamm_plan_specialised(self._h))

    def set_walker_split(self, mode):
        """0: one CTA per walker always; 1: automatic; 2/4/8: forced cluster size; 3: forced two-part split."""
        _lib.check(self.lib.spamm_plan_set_walker_split(self._h, int(mode)))

    def status(self):
        return int(self.lib.spamm_plan_status(self._h, self._stream()))

    # -- device entry points ------------------------------------------------------
    def lnposterior(self, params, out=None):
        """[W, D] CUDA float64 tensor -> [W] tensor, asynchronous on the current stream."""
        torch = _torch()
        self._check_params(params)
        W = params.shape[0]
        if out is None:
            out = torch.empty(W, dtype=torch.float64, device=params.device)
        _lib.check(self.lib.spamm_lnpost_batch(self._h, ctypes.c_void_p(params.data_ptr()), W,
                                               ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out

    def check(self):
        """Synchronise the current stream and raise SpammError if any walker so far needed a table entry
        the plan does not hold (its ln-posterior would be wrong, not -inf); clears the condition."""
        _lib.check(self.lib.spamm_plan_check(self._h, self._stream()))

    def lnposterior_sharded(self, params, exchange, out=None, copy=True):
        """Multi-GPU form of `lnposterior`: params is the full replicated [n, D] tensor; this rank
        evaluates its slice; the gather kernel behind it pulls the other ranks' slices over NVLink.
        Returns the gathered [n] vector (a copy in `out`), complete in stream order.  copy=False leaves
        the vector in the exchange region only (what the library's own sampler loop consumes) and
        returns None."""
        torch = _torch()
        self._check_params(params)
        n = params.shape[0]
        if copy and out is None:
            out = torch.empty(n, dtype=torch.float64, device=params.device)
        _lib.check(self.lib.spamm_lnpost_batch_sharded(self._h, exchange._h, ctypes.c_void_p(params.data_ptr()), n,
                                                       ctypes.c_void_p(out.data_ptr()) if copy else None, None,
                                                       self._stream()))
        return out if copy else None

    def lnposterior_sharded_host(self, params, exchange, out=None):
        """numpy [n, D] (the same array on every rank) -> numpy [n]: H2D of this rank's rows, kernel +
        exchange, D2H of the gathered vector, inside the library."""
        p = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
        if p.shape[1] != self.n_dim:
            raise ValueError("expected %d parameters per walker, got %d" % (self.n_dim, p.shape[1]))
        if out is None:
            out = np.empty(p.shape[0], dtype=np.float64)
        assert out.dtype == np.float64 and out.shape == (p.shape[0],) and out.flags.c_contiguous
        _lib.check(self.lib.spamm_lnpost_batch_sharded_host(self._h, exchange._h, p.ctypes.data_as(ctypes.c_void_p),
                                                            p.shape[0], out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def model_flux(self, params, mask=None, out=None):
        """[W, D] -> model flux [W, P] of the components selected by `mask` (all by default)."""
        torch = _torch()
        self._check_params(params)
        W = params.shape[0]
        if mask is None:
            mask = (1 << self.n_components) - 1
        if out is None:
            out = torch.empty((W, self.n_pix), dtype=torch.float64, device=params.device)
        _lib.check(self.lib.spamm_model_flux_batch(self._h, int(mask), ctypes.c_void_p(params.data_ptr()),
                                                   W, ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out

    def likelihood(self, model_flux, out=None):
        """[W, P] model flux -> ln-likelihood [W] (Model.likelihood)."""
        torch = _torch()
        W = model_flux.shape[0]
        assert model_flux.is_cuda and model_flux.dtype == torch.float64 and model_flux.is_contiguous()
        assert model_flux.shape[1] == self.n_pix
        if out is None:
            out = torch.empty(W, dtype=torch.float64, device=model_flux.device)
        _lib.check(self.lib.spamm_likelihood_batch(self._h, ctypes.c_void_p(model_flux.data_ptr()), W,
                                                   ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out

    # -- host-buffer entry points ----------------------------------------------------
    def lnposterior_host(self, params, out=None):
        """numpy [W, D] -> numpy [W]; H2D / kernel / D2H inside the library.  Page-locked
        buffers (e.g. views of torch pin_memory tensors) are used in place, pageable ones are
        staged through the plan's pinned buffers."""
        p = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
        if p.shape[1] != self.n_dim:
            raise ValueError("expected %d parameters per walker, got %d" % (self.n_dim, p.shape[1]))
        if out is None:
            out = np.empty(p.shape[0], dtype=np.float64)
        assert out.dtype == np.float64 and out.shape == (p.shape[0],) and out.flags.c_contiguous
        _lib.check(self.lib.spamm_lnpost_batch_host(self._h, p.ctypes.data_as(ctypes.c_void_p), p.shape[0],
                                                    out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def model_flux_host(self, params, mask=None):
        torch = _torch()
        p = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
        if p.shape[1] != self.n_dim:
            raise ValueError("expected %d parameters per walker, got %d" % (self.n_dim, p.shape[1]))
        out = np.empty((p.shape[0], self.n_pix), dtype=np.float64)
        step = max(1, min(p.shape[0], (256 << 20) // (8 * self.n_pix)))
        for s in range(0, p.shape[0], step):
            d = torch.from_numpy(p[s:s + step]).cuda()
            out[s:s + step] = self.model_flux(d, mask=mask).cpu().numpy()
        return out
