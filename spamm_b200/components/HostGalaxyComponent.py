"""Host-galaxy starlight: N stellar-population templates broadened by the
stellar velocity dispersion and scaled.
Reference: spamm/components/HostGalaxyComponent.py."""
import numpy as np

from ._templates import TemplateComponent
from .. import _lib
from ..utils.parse_pars import parse_pars


class HostGalaxyComponent(TemplateComponent):
    """Parameters: hg_norm_1..N, hg_stellar_disp (km/s)."""
    kind = _lib.COMP_HOST
    _clamp_edges = True        # np.interp without left/right, HG:179-181,323-325
    _sort_templates = False    # plain glob order, HG:93

    def __init__(self, pars=None):
        super(HostGalaxyComponent, self).__init__()
        self.inputpars = parse_pars()["host_galaxy"] if pars is None else pars
        self.load_templates()
        self.model_parameter_names = ["hg_norm_{}".format(x) for x in range(1, len(self.host_gal) + 1)]
        self.model_parameter_names.append("hg_stellar_disp")
        self.name = "HostGalaxy"
        self.norm_min = self.inputpars["hg_norm_min"]
        self.norm_max = self.inputpars["hg_norm_max"]
        self.stellar_disp_min = self.inputpars["hg_stellar_disp_min"]
        self.stellar_disp_max = self.inputpars["hg_stellar_disp_max"]
        self.templ_stellar_disp = self.inputpars["hg_template_stellar_disp"]
        if self.stellar_disp_min < self.templ_stellar_disp:      # HG:71-73
            print("Specified minimum stellar dispersion too small, setting to intrinsic template "
                  "stellar dispersion = {}".format(self.templ_stellar_disp))
            self.stellar_disp_min = self.templ_stellar_disp
        self.log_host = None

    def load_templates(self):
        self.host_gal = self._load(self.inputpars.get("hg_models"), "host")

    def is_analytic(self):
        return False

    @property
    def parameter_count(self):
        return len(self.host_gal) + 1

    def initialize(self, data_spectrum):
        """HG:154-188."""
        self.log_host = self._log_rebin(self.host_gal)

    def initial_values(self, spectrum):
        """HG:124-150."""
        self._resolve_norm_max(spectrum)
        norm_init = np.random.uniform(low=self.norm_min, high=self.norm_max, size=len(self.host_gal))
        stellar_disp_init = np.random.uniform(low=self.stellar_disp_min, high=self.stellar_disp_max)
        return norm_init.tolist() + [stellar_disp_init]

    def ln_priors(self, params):
        """HG:210-250."""
        out = []
        for i in range(len(self.host_gal)):
            norm = params[self.parameter_index("hg_norm_{}".format(i + 1))]
            out.append(0.0 if self.norm_min < norm < self.norm_max else -np.inf)
        disp = params[self.parameter_index("hg_stellar_disp")]
        out.append(0.0 if self.stellar_disp_min < disp < self.stellar_disp_max else -np.inf)
        return out

    def prior_bounds(self, spectrum=None):
        if spectrum is not None:
            self._resolve_norm_max(spectrum)
        n = len(self.host_gal)
        return ([self.norm_min] * n + [self.stellar_disp_min],
                [self.norm_max] * n + [self.stellar_disp_max])

    def plan_fill(self, desc, keep, spectrum):
        if self.log_host is None:
            self.initialize(spectrum)
        ts = self._template_set(self.host_gal, self.log_host, self.templ_stellar_disp,
                                self.inputpars["hg_kernel_size_sigma"], keep)
        self._rebin_operators(ts, self.log_host, spectrum, keep)
        desc.host = ts
