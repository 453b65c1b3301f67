"""Balmer continuum (lambda <= 3646 A, Grandi 1982) + pseudo-continuum of
high-order Balmer lines (lambda > 3646 A, Storey & Hummer 1995 case-B ratios,
Kovacevic et al. 2014).  Reference: spamm/components/BalmerContinuumCombined.py."""
import numpy as np

from .ComponentBase import Component
from .. import _lib, constants
from ..utils.find_nearest_index import find_nearest_index
from ..utils.parse_pars import parse_pars

balmer_edge = constants.BALMER_EDGE


class BalmerCombined(Component):
    """Parameters: bc_norm, bc_Te, bc_tauBE (continuum) and bc_loffset, bc_lwidth,
    bc_logNe (lines)."""
    kind = _lib.COMP_BALMER

    def __init__(self, pars=None, BalmerContinuum=False, BalmerPseudocContinuum=False):
        super().__init__()
        self.inputpars = parse_pars()["balmer_continuum"] if pars is None else pars
        self.name = "Balmer"
        self.model_parameter_names = ["bc_norm", "bc_Te", "bc_tauBE",
                                      "bc_loffset", "bc_lwidth", "bc_logNe"]
        self._norm_wavelength = None
        self.normalization_min = self.inputpars["bc_norm_min"]
        self.normalization_max = self.inputpars["bc_norm_max"]
        self.Te_min = self.inputpars["bc_Te_min"]
        self.Te_max = self.inputpars["bc_Te_max"]
        self.tauBE_min = self.inputpars["bc_tauBE_min"]
        self.tauBE_max = self.inputpars["bc_tauBE_max"]
        self.loffset_min = self.inputpars["bc_loffset_min"]
        self.loffset_max = self.inputpars["bc_loffset_max"]
        self.lwidth_min = self.inputpars["bc_lwidth_min"]
        self.lwidth_max = self.inputpars["bc_lwidth_max"]
        self.logNe_min = self.inputpars["bc_logNe_min"]
        self.logNe_max = self.inputpars["bc_logNe_max"]
        self.BC = BalmerContinuum
        self.BpC = BalmerPseudocContinuum
        # GPU line-sum strategy: False = far-field series for distant line clusters (default,
        # <= 1e-15 relative), True = every one of the 397 lines summed directly (BC:315-333)
        self.exact_line_sum = bool(self.inputpars.get("bc_exact_line_sum", False))

    @property
    def is_analytic(self):
        return True

    def _resolve(self, spectrum):
        """bcmax_flux -> max flux within +-10 pixels of the edge (BC:104-107)."""
        if self.normalization_max == "bcmax_flux":
            if spectrum is None:
                raise Exception("Need a data spectrum from which to estimate maximum flux at 3646 A")
            be_index = find_nearest_index(spectrum.spectral_axis, balmer_edge)
            self.normalization_max = np.max(spectrum.flux[be_index - 10:be_index + 10])

    def initial_values(self, spectrum=None):
        """Six uniform draws in the reference's order (BC:94-133)."""
        if spectrum is None:
            raise Exception("Need a data spectrum from which to estimate maximum flux at 3646 A")
        self._resolve(spectrum)
        lo, hi = self.prior_bounds()
        return [np.random.uniform(low=l, high=h) for l, h in zip(lo, hi)]

    def ln_priors(self, params):
        """BC:137-196."""
        lo, hi = self.prior_bounds()
        return [0 if l < params[self.parameter_index(n)] < h else -np.inf
                for n, l, h in zip(self.model_parameter_names, lo, hi)]

    def prior_bounds(self, spectrum=None):
        if spectrum is not None:
            self._resolve(spectrum)
        return ([self.normalization_min, self.Te_min, self.tauBE_min, self.loffset_min,
                 self.lwidth_min, self.logNe_min],
                [self.normalization_max, self.Te_max, self.tauBE_max, self.loffset_max,
                 self.lwidth_max, self.logNe_max])

    def balmerseries(self, line_orders):
        """Wavelength [A] of the n -> 2 transition (BC:247-259)."""
        ilambda = constants.RYD_AA * (0.25 - 1. / line_orders ** 2)
        return 1. / ilambda

    def plan_fill(self, desc, keep, spectrum):
        if not (self.BC or self.BpC):
            # the reference dies with UnboundLocalError in flux() (BC:462-473)
            raise ValueError("BalmerCombined needs BalmerContinuum and/or BalmerPseudocContinuum")
        ltype = self.inputpars["bc_line_type"]
        if ltype not in ("gaussian", "lorentzian"):
            raise ValueError("Variable 'ltype' ({0}) must be 'gaussian' or 'lorentzian'".format(ltype))
        desc.balmer_bc = 1 if self.BC else 0
        desc.balmer_bpc = 1 if self.BpC else 0
        desc.balmer_line_type = _lib.LINE_GAUSSIAN if ltype == "gaussian" else _lib.LINE_LORENTZIAN
        line_orders = np.arange(self.inputpars["bc_lines_min"], self.inputpars["bc_lines_max"])  # BC:311
        minline = int(line_orders.min())
        desc.balmer_line_min = minline
        desc.balmer_n_lines = len(line_orders)
        desc.balmer_exact_line_sum = 1 if self.exact_line_sum else 0
        centers, cp = _lib.as_c(self.balmerseries(line_orders))
        # E0*(1./i**2-1./(i-1)**2) for integer i (BC:296)
        ns = [minline + k for k in range(len(line_orders))]
        exparg, ep = _lib.as_c(np.array(
            [constants.E0 * (1. / i ** 2 - 1. / (i - 1) ** 2) for i in ns], dtype=np.float64))
        desc.balmer_line_center, desc.balmer_line_exparg = cp, ep
        keep += [centers, exparg]
        from ..data import load_sh95
        ne, T, z = load_sh95(self.inputpars.get("sh95_coefficients"))
        assert z.shape == (_lib.SH95_LINES, _lib.SH95_NE, _lib.SH95_T)
        for i in range(_lib.SH95_NE):
            desc.sh95_ne[i] = ne[i]
        for i in range(_lib.SH95_T):
            desc.sh95_T[i] = T[i]
        zc, zp = _lib.as_c(z)
        desc.sh95_z = zp
        keep.append(zc)
        # log_conv's resampling grids (BC:218-221, 239)
        wl = np.asarray(spectrum.spectral_axis, dtype=np.float64)
        ln_wave = np.log(wl)
        ln_wavenew = np.r_[ln_wave.min():ln_wave.max():1j * ln_wave.size]
        ln_wavenew[0] = ln_wave[0]
        ln_wavenew[-1] = ln_wave[-1]
        a, ap = _lib.as_c(ln_wavenew)
        b, bp = _lib.as_c(np.exp(ln_wavenew))
        desc.bc_ln_wl_new, desc.bc_exp_ln_wl_new = ap, bp
        keep += [a, b]
        if self.BC and not self.fast_interp:
            # rebin_spec: True (BC:227,241): both resamplings of log_conv are flux-conserving rebins; with the
            # identity kernel between them (App. B: c in cm/s) they compose into one sparse operator on f0
            from ..utils.rebin_spec import rebin_matrix, compose_csr
            to_log = rebin_matrix(ln_wave, ln_wavenew)
            back = rebin_matrix(np.exp(ln_wavenew), wl)
            ip, ix, w = compose_csr(back, to_log, len(wl))
            ipc, ipp = _lib.as_c_i32(ip)
            ixc, ixp = _lib.as_c_i32(ix)
            wc, wp = _lib.as_c(w)
            keep += [ipc, ixc, wc]
            desc.bc_rebin_indptr, desc.bc_rebin_index, desc.bc_rebin_weight = ipp, ixp, wp
        desc.c_kms, desc.c_cgs, desc.c_aa = constants.C_KMS, constants.C_CGS, constants.C_AA
        desc.h_cgs, desc.kb_cgs = constants.H_CGS, constants.KB_CGS
        desc.balmer_edge = float(balmer_edge)
