"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  Nothing under
``spamm_b200/`` imports it; the product path fails loudly without its CUDA
library.

Parity status: PINNED.  The restatement is checked (tests/test_oracle_*.py)
against (a) the unmodified reference imported under ``oracle/refshim.py`` in
the build container, through committed fixtures in ``tests/golden/`` made by
``tests/golden/make_golden.py``; (b) the only known-answer artefact the
reference ships: the 150 000 (params, lnprob) pairs stored in
``examples/model.pickle.gz``.

Every function cites the reference lines it follows (paths relative to the
reference root).  The arithmetic is kept in the reference's operation order,
one walker at a time, *including* the 397 x P line matrix of ``makelines`` --
this is also the "port" CPU baseline that bench.py times.
"""
import os

import numpy as np

# --- constants: astropy 3.x (CODATA-2014), converted the way oracle/refshim.py does
C_MS = 299792458.0
C_KMS = C_MS / 1000.0                 # c.to("km/s")        FeComponent.py:270, BC:373
C_CGS = C_MS * 100.0                  # c.cgs               BalmerContinuumCombined.py:21
C_AA = C_MS * 1e10                    # Angstrom / s (blackbody_lambda)
H_CGS = 6.626070040e-34 * 1e7         # h.cgs               BC:22
KB_CGS = 1.38064852e-23 * 1e7         # k_B.cgs             BC:23
RYD_AA = 10973731.568508 * 1e-10      # Ryd.to("1/Angstrom") BC:24
E0 = 2.179e-11                        # BC:25
BALMER_EDGE = 3646                    # BC:26

DEFAULT_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                            "spamm_b200", "data", "spamm_data.npz")

# utils/parameters.yaml defaults (values only; the product has its own parser)
DEFAULT_PARS = {
    "host_galaxy": dict(boxcar_width=5, hg_norm_min=0., hg_norm_max="max_flux",
                        hg_stellar_disp_min=30., hg_stellar_disp_max=1000.,
                        hg_template_stellar_disp=0.0, hg_kernel_size_sigma=10),
    "nuclear_continuum": dict(boxcar_width=5, broken_pl=False, pl_slope_min=-3.,
                              pl_slope_max=3., pl_norm_min=0., pl_norm_max="fnw",
                              pl_wave_break_min="min_wl", pl_wave_break_max="max_wl"),
    "balmer_continuum": dict(bc_line_type="lorentzian", bc_lines_min=3., bc_lines_max=400.,
                             bc_norm_min=0., bc_norm_max="bcmax_flux", bc_Te_min=500.,
                             bc_Te_max=100000., bc_tauBE_min=0., bc_tauBE_max=2.,
                             bc_loffset_min=-10., bc_loffset_max=10., bc_lwidth_min=100.,
                             bc_lwidth_max=10000., bc_logNe_min=2, bc_logNe_max=9),
    "fe_forest": dict(boxcar_width=10, fe_template_width=900., fe_norm_min=0.,
                      fe_norm_max="fnw", fe_width_min=901., fe_width_max=10000.,
                      fe_kernel_size_sigma=3),
}


# ----------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------
def find_nearest_index(arr, value):
    """utils/find_nearest_index.py:21-34 (first minimum on ties)."""
    return int(np.abs(np.asarray(arr, dtype=float) - value).argmin())


def running_mean_fast(x, n):
    """utils/runningmeanfast.py:19."""
    return np.convolve(x, np.ones((n,)) / n)[(n - 1):]


def gaussian_window(m, std):
    """scipy.signal.gaussian(M, std) (FeComponent.py:292): symmetric window
    exp(-0.5 (n/std)^2), n = arange(M) - (M-1)/2."""
    m = int(m)
    n = np.arange(0, m) - (m - 1.0) / 2.0
    return np.exp(-n ** 2 / (2 * std * std))


def blackbody_lambda(lam, temperature):
    """astropy 3.x modeling.blackbody.blackbody_lambda for lam in Angstrom
    (call site BalmerContinuumCombined.py:435)."""
    lam = np.asarray(lam, dtype=np.float64)
    freq = C_AA / lam
    log_boltz = H_CGS * freq / (KB_CGS * temperature)
    boltzm1 = np.expm1(log_boltz)
    bnu = 2.0 * H_CGS * freq ** 3 / (C_CGS ** 2 * boltzm1)
    return bnu * C_AA / lam ** 2


def sh95_bilinear(table, ne_knots, t_knots, n_e, temp):
    """Value of the 48 pickled kx=ky=1 interp2d objects at (n_e, T)
    (BalmerContinuumCombined.py:288-290).  FITPACK evaluation with both
    arguments clamped into the knot range; linear in n_e and in T."""
    x = min(max(n_e, ne_knots[0]), ne_knots[-1])
    y = min(max(temp, t_knots[0]), t_knots[-1])
    i = int(np.searchsorted(ne_knots, x, side="right") - 1)
    i = min(max(i, 0), len(ne_knots) - 2)
    j = int(np.searchsorted(t_knots, y, side="right") - 1)
    j = min(max(j, 0), len(t_knots) - 2)
    tx = (x - ne_knots[i]) / (ne_knots[i + 1] - ne_knots[i])
    ty = (y - t_knots[j]) / (t_knots[j + 1] - t_knots[j])
    z = table
    return ((1 - tx) * (1 - ty) * z[:, i, j] + tx * (1 - ty) * z[:, i + 1, j]
            + (1 - tx) * ty * z[:, i, j + 1] + tx * ty * z[:, i + 1, j + 1])


class Templates(object):
    """Raw template arrays + SH95 table from the packed npz
    (tools/pack_reference_data.py)."""

    def __init__(self, path=None):
        z = np.load(path or DEFAULT_DATA)
        self.fe = [(z["fe_wl_%d" % i], z["fe_flux_%d" % i]) for i in range(len(z["fe_names"]))]
        self.host = [(z["host_wl_%d" % i], z["host_flux_%d" % i])
                     for i in range(len(z["host_names"]))]
        self.sh95_ne, self.sh95_T, self.sh95_z = z["sh95_ne"], z["sh95_T"], z["sh95_z"]


# ----------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------
def extinction_k(wl, law):
    """k(lambda) of the five curves of spamm/components/ReddeningLaw.py, so that the extinction is
    pow(10, -0.4 * E(B-V) * k): Calzetti_ext :7-20 (k = 0, i.e. ext = 1, outside 0.12-2.2 um),
    LMC_Fitzpatrick_ext :22-46, MW_Seaton_ext :48-72, SMC_Gordon_ext :74-98 (its far-UV C4 is
    0.26, not the 0.46 declared above the loop), AGN_Gaskell_ext :100-123 (all wavelengths: the
    range test is commented out).  Scalar arithmetic in the reference's operation order."""
    wl = np.asarray(wl, dtype=np.float64)
    if law == "AGN":                       # the one curve the reference evaluates on whole arrays (:118-122)
        A, B, C, D, E, F, Rv = 0.000843, -0.02496, 0.2919, -1.815, 6.83, -7.92, 5.0
        x = pow(np.array(wl / 10000.), -1)
        return A * x ** 5 + B * x ** 4 + C * x ** 3 + D * x ** 2 + E * x + F + Rv
    k = np.zeros(len(wl))
    for j in range(len(wl)):
        um = wl[j] / 10000.
        if law == "Calzetti":
            if (um >= 0.12) & (um < 0.63):
                k[j] = 2.659 * (-1.857 + 1.040 / um) + 4.05
            if (um >= 0.63) & (um <= 2.2):
                k[j] = 2.659 * (-2.156 + 1.509 / um - 0.198 / pow(um, 2) + 0.11 / pow(um, 3)) + 4.05
        elif law in ("LMC", "MW", "SMC"):
            C1, C2, C3, C4uv, x_0, gamma, Rv = {
                "LMC": (-0.69, 0.89, 2.55, 0.50, 4.608, 0.994, 3.1),
                "MW": (-0.38, 0.74, 3.96, 0.26, 4.595, 1.051, 3.1),
                "SMC": (-4.96, 2.26, 0.39, 0.26, 4.6, 1.0, 2.74)}[law]
            x = pow(um, -1)
            x2 = pow(x, 2)
            D = x2 / (pow(x2 - pow(x_0, 2), 2) + x2 * pow(gamma, 2))
            F = 0.5392 * pow((x - 5.9), 2) + 0.05644 * pow((x - 5.9), 3)
            C4 = C4uv if x >= 5.9 else 0.0
            k[j] = C1 + Rv + C2 * x + C3 * x * D + C4 * F
        else:
            raise KeyError(law)
    return k


class OracleModel(object):
    """Restates Model + the components for one data spectrum.

    comps: ordered list drawn from "PL", "FE", "HOST", "BC", "EXT" (the order fixes the
    parameter layout exactly like Model.components, spamm/Model.py:353-364;
    run_spamm.py:100-125 appends PL, FE, HOST, BC, Extinction).  "EXT" multiplies the flux
    accumulated so far by the `ext_law` curve (Model.py:358-361, 384-395).
    """

    def __init__(self, wl, flux, err, comps=("PL", "FE", "HOST", "BC"), bc=True, bpc=True,
                 broken_pl=False, line_type=None, templates=None, pars=None, ext_law=None, rebin_spec=False):
        self.wl = np.asarray(wl, dtype=np.float64)
        self.flux = np.asarray(flux, dtype=np.float64)
        self.err = np.asarray(err, dtype=np.float64)
        self.comps = [c.upper() for c in comps]
        self.bc_on, self.bpc_on = bool(bc), bool(bpc)
        self.broken_pl = bool(broken_pl)
        self.pars = {k: dict(v) for k, v in (pars or DEFAULT_PARS).items()}
        self.line_type = line_type or self.pars["balmer_continuum"]["bc_line_type"]
        self.tpl = templates if templates is not None else Templates()
        self.norm_wl = np.median(self.wl)                     # Spectrum.py:84-87
        self.ext_law = ext_law
        # rebin_spec: True (parameters.yaml): every resampling goes through utils/rebin_spec.py, restated in
        # extended precision by oracle/rebin_longdouble.py (parity unpinned, see there)
        self.rebin_spec = bool(rebin_spec)
        # Extinction.extinction :188-205: no law selected => [1.] * P
        self.ext_k = extinction_k(self.wl, ext_law) if ext_law else np.zeros(len(self.wl))
        self._bounds = None
        self._init_templates()

    # -- initialise ----------------------------------------------------------
    def _init_templates(self):
        """FeComponent.initialize :156-191 / HostGalaxyComponent.initialize :154-188."""
        self.log_fe = []
        for wl_t, fx_t in self.tpl.fe:
            fx_t = np.where(fx_t < 0, 1e-19, fx_t)               # FeComponent.py:88
            lw = np.log(wl_t)
            x = np.linspace(min(lw), max(lw), num=len(wl_t))     # :173-175
            g = np.interp(x, lw, fx_t, left=0, right=0)          # :181-182
            if self.rebin_spec:
                g = self._rebin(lw, fx_t, x)                     # :184-186
            self.log_fe.append((x, g, np.median(wl_t)))          # median: :325
        self.log_host = []
        for wl_t, fx_t in self.tpl.host:
            fx_t = np.where(fx_t < 0, 1e-19, fx_t)               # HostGalaxyComponent.py:102
            lw = np.log(wl_t)
            x = np.linspace(min(lw), max(lw), num=len(wl_t))     # :173-175
            g = np.interp(x, lw, fx_t)                           # :179-181 (edge clamp)
            if self.rebin_spec:
                g = self._rebin(lw, fx_t, x)                     # :183-185
            self.log_host.append((x, g, np.median(wl_t)))        # median: :330

    @staticmethod
    def _rebin(wave, specin, wavnew):
        from oracle.rebin_longdouble import rebin_spec_longdouble
        return np.asarray(rebin_spec_longdouble(wave, specin, wavnew), dtype=np.float64)

    # -- parameter bookkeeping ----------------------------------------------
    def names(self):
        out = []
        for c in self.comps:
            if c == "PL":
                out += (["wave_break", "norm_PL", "slope1", "slope2"] if self.broken_pl
                        else ["norm_PL", "slope1"])               # NC:51-58
            elif c == "FE":
                out += ["fe_norm_%d" % i for i in range(1, len(self.log_fe) + 1)] + ["fe_width"]
            elif c == "HOST":
                out += ["hg_norm_%d" % i for i in range(1, len(self.log_host) + 1)] \
                    + ["hg_stellar_disp"]
            elif c == "BC":
                out += ["bc_norm", "bc_Te", "bc_tauBE", "bc_loffset", "bc_lwidth", "bc_logNe"]
            elif c == "EXT":
                out += ["E(B-V)"]                                  # ReddeningLaw.py:133
        return out

    def count(self, c):
        return {"PL": 4 if self.broken_pl else 2, "FE": len(self.log_fe) + 1,
                "HOST": len(self.log_host) + 1, "BC": 6, "EXT": 1}[c]

    @property
    def ndim(self):
        return sum(self.count(c) for c in self.comps)

    # -- priors ---------------------------------------------------------------
    def norm_wavelength_flux(self):
        """Spectrum.norm_wavelength_flux :90-95 (scipy interp1d, linear)."""
        return float(np.interp(self.norm_wl, self.wl, self.flux))

    def bounds(self):
        """(lo[D], hi[D]) after the string sentinels were resolved the way
        initial_values does (NC:98-109, Fe:139-144, HG:135-140, BC:104-107)."""
        if self._bounds is not None:
            return self._bounds
        lo, hi = [], []
        for c in self.comps:
            if c == "PL":
                p = self.pars["nuclear_continuum"]
                nmax = p["pl_norm_max"]
                if nmax == "max_flux":
                    nmax = max(running_mean_fast(self.flux, p["boxcar_width"]))
                elif nmax == "fnw":
                    nmax = self.norm_wavelength_flux()
                if self.broken_pl:
                    bmin, bmax = p["pl_wave_break_min"], p["pl_wave_break_max"]
                    bmin = min(self.wl) if bmin == "min_wl" else bmin
                    bmax = max(self.wl) if bmax == "max_wl" else bmax
                    lo += [bmin, p["pl_norm_min"], p["pl_slope_min"], p["pl_slope_min"]]
                    hi += [bmax, nmax, p["pl_slope_max"], p["pl_slope_max"]]
                else:
                    lo += [p["pl_norm_min"], p["pl_slope_min"]]
                    hi += [nmax, p["pl_slope_max"]]
            elif c == "FE":
                p = self.pars["fe_forest"]
                nmax = p["fe_norm_max"]
                if nmax == "max_flux":
                    nmax = max(running_mean_fast(self.flux, p["boxcar_width"]))
                elif nmax == "fnw":
                    nmax = self.norm_wavelength_flux()
                wmin = max(p["fe_width_min"], p["fe_template_width"])   # Fe:69-71
                nt = len(self.log_fe)
                lo += [p["fe_norm_min"]] * nt + [wmin]
                hi += [nmax] * nt + [p["fe_width_max"]]
            elif c == "HOST":
                p = self.pars["host_galaxy"]
                nmax = p["hg_norm_max"]
                if nmax == "max_flux":
                    nmax = max(running_mean_fast(self.flux, p["boxcar_width"]))
                elif nmax == "fnw":
                    nmax = self.norm_wavelength_flux()
                dmin = max(p["hg_stellar_disp_min"], p["hg_template_stellar_disp"])  # HG:71-73
                nh = len(self.log_host)
                lo += [p["hg_norm_min"]] * nh + [dmin]
                hi += [nmax] * nh + [p["hg_stellar_disp_max"]]
            elif c == "BC":
                p = self.pars["balmer_continuum"]
                nmax = p["bc_norm_max"]
                if nmax == "bcmax_flux":
                    be = find_nearest_index(self.wl, BALMER_EDGE)        # BC:105
                    nmax = np.max(self.flux[be - 10:be + 10])            # BC:106
                lo += [p["bc_norm_min"], p["bc_Te_min"], p["bc_tauBE_min"], p["bc_loffset_min"],
                       p["bc_lwidth_min"], p["bc_logNe_min"]]
                hi += [nmax, p["bc_Te_max"], p["bc_tauBE_max"], p["bc_loffset_max"],
                       p["bc_lwidth_max"], p["bc_logNe_max"]]
            elif c == "EXT":
                lo += [0]                                            # ReddeningLaw.py:160-162
                hi += [2.]
        self._bounds = (np.array(lo, dtype=np.float64), np.array(hi, dtype=np.float64))
        return self._bounds

    def set_bounds(self, lo, hi):
        self._bounds = (np.array(lo, dtype=np.float64), np.array(hi, dtype=np.float64))

    def initial_values(self, rng=np.random):
        """One walker, drawn in the reference's RNG call order
        (Model.run_mcmc :255-259; NC:116-122, Fe:146-152, HG:144-150, BC:108-133)."""
        lo, hi = self.bounds()
        out, k = [], 0
        for c in self.comps:
            n = self.count(c)
            l, h = lo[k:k + n], hi[k:k + n]
            if c == "PL":
                if self.broken_pl:
                    out.append(rng.uniform(low=l[0], high=h[0]))
                    out.append(rng.uniform(l[1], high=h[1]))
                    out += list(rng.uniform(low=l[2], high=h[2], size=2))
                else:
                    out.append(rng.uniform(l[0], high=h[0]))
                    out += list(rng.uniform(low=l[1], high=h[1], size=1))
            elif c in ("FE", "HOST"):
                out += rng.uniform(low=l[0], high=h[0], size=n - 1).tolist()
                out.append(rng.uniform(low=l[-1], high=h[-1]))
            elif c == "EXT":
                out.append(rng.uniform(low=l[0], high=h[0]))          # ReddeningLaw.py:163-165
            else:
                for i in range(6):
                    out.append(rng.uniform(low=l[i], high=h[i]))
            k += n
        return out

    def prior(self, params):
        """Model.prior :449-470 with the flat-box ln_priors of every component
        (strict inequalities: NC:148-170, Fe:241-248, HG:232-248, BC:161-189)."""
        lo, hi = self.bounds()
        p = np.asarray(params, dtype=np.float64)
        ok = (lo < p) & (p < hi)
        return 0.0 if bool(np.all(ok)) else -np.inf

    # -- components -----------------------------------------------------------
    def nc_flux(self, p):
        """NuclearContinuumComponent.flux :176-201 (astropy PowerLaw1D /
        BrokenPowerLaw1D semantics)."""
        if not self.broken_pl:
            norm, slope1 = p[0], p[1]
            xx = self.wl / self.norm_wl
            return norm * xx ** (-slope1)
        x_break, norm, slope1, slope2 = p
        alpha = np.where(self.wl < x_break, slope1, slope2)
        xx = self.wl / x_break
        return norm * xx ** (-alpha)

    def _template_flux(self, p, logt, templ_width, ksize, clamp_edges):
        """FeComponent.flux :254-335 / HostGalaxyComponent.flux :254-341."""
        width = p[-1]
        out = np.zeros(len(self.wl))
        lnwl = np.log(self.wl)
        for i, (x, g, med) in enumerate(logt):
            sigma_conv = np.sqrt(width ** 2 - templ_width ** 2) / \
                (C_KMS * 2. * np.sqrt(2. * np.log(2.)))           # Fe:285-286
            bin_size = x[2] - x[1]                                # :288
            sigma_norm = np.ceil(sigma_conv / bin_size)           # :290
            sigma_size = ksize * sigma_norm                       # :291
            kernel = gaussian_window(sigma_size, sigma_norm) / \
                (np.sqrt(2 * np.pi) * sigma_norm)                 # :292-293
            conv = np.convolve(g, kernel, mode="same")            # :304
            if self.rebin_spec:
                f = self._rebin(x, conv, lnwl)                    # Fe:320-323 / HG:326-329
            elif clamp_edges:
                f = np.interp(lnwl, x, conv)                      # HG:323-325
            else:
                f = np.interp(lnwl, x, conv, left=0, right=0)     # Fe:315-319
            nf = np.interp(med, self.wl, f)                       # :325-326
            nf = np.nan_to_num(nf)                                # :331
            with np.errstate(divide="ignore", invalid="ignore"):
                out += (p[i] / nf) * f                            # :334
        return out

    def fe_flux(self, p):
        pr = self.pars["fe_forest"]
        return self._template_flux(p, self.log_fe, pr["fe_template_width"],
                                   pr["fe_kernel_size_sigma"], clamp_edges=False)

    def host_flux(self, p):
        pr = self.pars["host_galaxy"]
        return self._template_flux(p, self.log_host, pr["hg_template_stellar_disp"],
                                   pr["hg_kernel_size_sigma"], clamp_edges=True)

    def log_conv(self, orig_flux, width_lines):
        """BalmerCombined.log_conv :202-243 (fast_interp branch)."""
        ln_wave = np.log(self.wl)
        ln_wavenew = np.r_[ln_wave.min():ln_wave.max():1j * ln_wave.size]
        ln_wavenew[0] = ln_wave[0]
        ln_wavenew[-1] = ln_wave[-1]
        flux_rebin = np.interp(ln_wavenew, ln_wave, orig_flux)
        if self.rebin_spec:
            flux_rebin = self._rebin(ln_wave, orig_flux, ln_wavenew)              # :227
        dpix = width_lines / (ln_wavenew[1] - ln_wavenew[0])
        kernel_width = round(5 * dpix)
        kernel_x = np.r_[-kernel_width:kernel_width + 1]
        kernel = np.exp(- (kernel_x) ** 2 / (dpix) ** 2)
        kernel /= abs(np.sum(kernel))
        flux_conv = np.convolve(flux_rebin, kernel, mode="same")
        if self.rebin_spec:
            return self._rebin(np.exp(ln_wavenew), flux_conv, self.wl)            # :241
        return np.interp(self.wl, np.exp(ln_wavenew), flux_conv)

    def bc_flux(self, p):
        """BalmerCombined.BC_flux :394-448 (note c in cm/s, :21,432,441)."""
        normalization, Te, tauBE, loffset, lwidth, logNe = p
        edge_wl = BALMER_EDGE * (1 - loffset / C_CGS)
        bb = blackbody_lambda(self.wl, Te)
        tau = tauBE * (self.wl / BALMER_EDGE) ** 3
        absorption = 1 - np.exp(-tau)
        f = absorption * bb
        f = self.log_conv(f, lwidth / C_CGS)
        ni = find_nearest_index(self.wl, edge_wl)
        fnorm = f[ni]
        f[self.wl > self.wl[ni]] = 0.
        with np.errstate(divide="ignore", invalid="ignore"):
            f *= normalization / fnorm
        return f

    def balmer_ratio(self, n_e, line_orders, T):
        """BalmerCombined.balmer_ratio :264-297."""
        maxline = int(line_orders.max())
        minline = int(line_orders.min())
        flux_ratios = np.zeros(maxline - minline + 1)
        n = np.arange(minline, maxline + 1)
        coef_use = sh95_bilinear(self.tpl.sh95_z, self.tpl.sh95_ne, self.tpl.sh95_T, n_e, T)
        for i in n:
            if i <= 50:
                flux_ratios[i - minline] = coef_use[-i + 2]
            else:
                flux_ratios[i - minline] = flux_ratios[i - minline - 1] * \
                    np.exp(E0 * (1. / i ** 2 - 1. / (i - 1) ** 2) / (KB_CGS * T))
        return flux_ratios

    def makelines(self, T, n_e, shift, width):
        """BalmerCombined.makelines :301-335 -- the 397 x P matrix."""
        pr = self.pars["balmer_continuum"]
        line_orders = np.arange(pr["bc_lines_min"], pr["bc_lines_max"])
        lcenter = 1. / (RYD_AA * (0.25 - 1. / line_orders ** 2))   # balmerseries :258-259
        lcenter -= shift * lcenter
        LL = self.wl - lcenter.reshape(lcenter.size, 1)
        lwidth = width * lcenter.reshape(lcenter.size, 1)
        if self.line_type == "gaussian":
            lines = np.exp(- LL ** 2 / lwidth ** 2)
        elif self.line_type == "lorentzian":
            lines = lwidth / (LL ** 2 + lwidth ** 2)
        else:
            raise ValueError("line type")
        lflux = self.balmer_ratio(n_e, line_orders, T)
        lines *= lflux.reshape(lflux.size, 1)
        return np.sum(lines, axis=0)

    def bpc_flux(self, p):
        """BalmerCombined.BpC_flux :341-390."""
        normalization, Te, tauBE, loffset, lwidth, logNe = p
        edge_wl = BALMER_EDGE * (1 - loffset / C_KMS)
        n_e = 10. ** logNe
        f = self.makelines(Te, n_e, loffset / C_KMS, lwidth / C_KMS)
        ni = find_nearest_index(self.wl, edge_wl)
        fnorm = f[ni]
        f[self.wl <= self.wl[ni]] = 0
        with np.errstate(divide="ignore", invalid="ignore"):
            f *= normalization / fnorm
        return f

    def balmer_flux(self, p):
        """BalmerCombined.flux :455-473."""
        if self.bc_on and self.bpc_on:
            return self.bc_flux(p) + self.bpc_flux(p)
        if self.bc_on:
            return self.bc_flux(p)
        if self.bpc_on:
            return self.bpc_flux(p)
        raise ValueError("BalmerCombined needs BC and/or BpC")   # reference: UnboundLocalError

    def extinction(self, p):
        """Extinction.extinction :188-205: pow(10, -0.4 * E(B-V) * k) per pixel (AGN: on the array)."""
        if self.ext_law == "AGN":
            return pow(10, -0.4 * p[0] * self.ext_k)
        return np.array([pow(10, -0.4 * p[0] * kj) for kj in self.ext_k])

    def component_flux(self, c, p):
        if c == "EXT":
            return np.zeros(len(self.wl))                          # Extinction.flux :180-185
        return {"PL": self.nc_flux, "FE": self.fe_flux, "HOST": self.host_flux,
                "BC": self.balmer_flux}[c](np.asarray(p, dtype=np.float64))

    # -- model ------------------------------------------------------------------
    def model_flux(self, params):
        """Model.model_flux :328-366."""
        params = np.asarray(params, dtype=np.float64)
        f = np.zeros(len(self.wl))
        k = 0
        for c in self.comps:
            n = self.count(c)
            if c == "EXT":
                f = np.array(f) * self.extinction(params[k:k + n])   # Model.reddening :384-395
            else:
                f += self.component_flux(c, params[k:k + n])
            k += n
        return f

    def likelihood(self, model_flux):
        """Model.likelihood :418-445.  interp1d(model grid -> data grid) is an
        identity here because the grids are the same array (Model.py:189)."""
        m = np.interp(self.wl, self.wl, model_flux)
        with np.errstate(divide="ignore", invalid="ignore"):
            ln_l = np.power((self.flux - m) / self.err, 2) + \
                np.log(2 * np.pi * np.power(self.err, 2))
            ln_l = -0.5 * np.sum(ln_l)
        return float(np.nan_to_num(ln_l))

    def ln_posterior(self, params):
        """ln_posterior, spamm/Model.py:42-79."""
        lp = self.prior(params)
        if not np.isfinite(lp):
            return -np.inf
        return self.likelihood(self.model_flux(params)) + lp

    def ln_posterior_batch(self, params):
        params = np.atleast_2d(np.asarray(params, dtype=np.float64))
        return np.array([self.ln_posterior(p) for p in params])


# ----------------------------------------------------------------------------
# emcee 2.2.1 stretch move -- used by the sampler tests
# ----------------------------------------------------------------------------
def stretch_move_emcee(lnpostfn, p0, niter, a=2.0, rng=None):
    """Restates emcee 2.2.1 EnsembleSampler.sample/_propose_stretch: two half
    ensembles per iteration, z ~ g(z), q = c - z (c - s), accept if
    (D-1) ln z + lnp(q) - lnp(s) > ln U.  lnpostfn maps [n, D] -> [n]."""
    rng = rng or np.random.RandomState()
    p = np.array(p0, dtype=np.float64)
    k, dim = p.shape
    assert k % 2 == 0 and k >= 2 * dim
    half = k // 2
    lnprob = np.asarray(lnpostfn(p), dtype=np.float64)
    chain = np.empty((k, niter, dim))
    lnps = np.empty((k, niter))
    naccepted = np.zeros(k)
    first, second = slice(half), slice(half, k)
    for it in range(niter):
        for S0, S1 in [(first, second), (second, first)]:
            s = p[S0]
            c = p[S1]
            Ns, Nc = len(s), len(c)
            zz = ((a - 1.) * rng.rand(Ns) + 1) ** 2. / a
            rint = rng.randint(Nc, size=(Ns,))
            q = c[rint] - zz[:, np.newaxis] * (c[rint] - s)
            newlnprob = np.asarray(lnpostfn(q), dtype=np.float64)
            lnpdiff = (dim - 1.) * np.log(zz) + newlnprob - lnprob[S0]
            accept = (lnpdiff > np.log(rng.rand(len(lnpdiff))))
            idx = np.arange(k)[S0][accept]
            lnprob[idx] = newlnprob[accept]
            p[idx] = q[accept]
            naccepted[idx] += 1
        chain[:, it, :] = p
        lnps[:, it] = lnprob
    return chain, lnps, naccepted / max(niter, 1)
