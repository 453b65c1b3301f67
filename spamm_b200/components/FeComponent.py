"""Fe II / Fe III pseudo-continuum: N broadened, scaled templates.
Reference: spamm/components/FeComponent.py."""
import numpy as np

from ._templates import TemplateComponent
from .. import _lib
from ..utils.parse_pars import parse_pars


class FeComponent(TemplateComponent):
    """Parameters: fe_norm_1..N (template normalisations), fe_width (FWHM, km/s)."""
    kind = _lib.COMP_FE
    _clamp_edges = False       # np.interp(..., left=0, right=0), Fe:315-319
    _sort_templates = True     # sorted(glob(...)), Fe:79

    def __init__(self, pars=None):
        super(FeComponent, self).__init__()
        self.inputpars = parse_pars()["fe_forest"] if pars is None else pars
        self.load_templates()
        self.model_parameter_names = ["fe_norm_{0}".format(x) for x in range(1, len(self.fe_templ) + 1)]
        self.model_parameter_names.append("fe_width")
        self.name = "FeForest"
        self.norm_min = self.inputpars["fe_norm_min"]
        self.norm_max = self.inputpars["fe_norm_max"]
        self.width_min = self.inputpars["fe_width_min"]
        self.width_max = self.inputpars["fe_width_max"]
        self.templ_width = self.inputpars["fe_template_width"]
        if self.width_min < self.templ_width:                     # Fe:69-71
            print("Specified minimum Fe width too small, setting Fe width minimum to intrinsic "
                  "template width = {0}".format(self.templ_width))
            self.width_min = self.templ_width
        self.log_fe = None

    def load_templates(self):
        self.fe_templ = self._load(self.inputpars.get("fe_templates"), "fe")

    def is_analytic(self):
        # NB: a *method* in the reference too (Fe:95), hence always truthy there
        return False

    @property
    def parameter_count(self):
        return len(self.fe_templ) + 1

    def initialize(self, data_spectrum):
        """Rebin every template to equal ln-lambda bins, once (Fe:156-191)."""
        self.log_fe = self._log_rebin(self.fe_templ)

    def initial_values(self, spectrum):
        """Fe:122-152."""
        self._resolve_norm_max(spectrum)
        norm_init = np.random.uniform(low=self.norm_min, high=self.norm_max, size=len(self.fe_templ))
        width_init = np.random.uniform(low=self.width_min, high=self.width_max)
        return norm_init.tolist() + [width_init]

    def ln_priors(self, params):
        """Fe:218-250."""
        out = []
        for i in range(len(self.fe_templ)):
            norm = params[self.parameter_index("fe_norm_{0}".format(i + 1))]
            out.append(0.0 if self.norm_min < norm < self.norm_max else -np.inf)
        width = params[self.parameter_index("fe_width")]
        out.append(0. if self.width_min < width < self.width_max else -np.inf)
        return out

    def prior_bounds(self, spectrum=None):
        if spectrum is not None:
            self._resolve_norm_max(spectrum)
        n = len(self.fe_templ)
        return [self.norm_min] * n + [self.width_min], [self.norm_max] * n + [self.width_max]

    def plan_fill(self, desc, keep, spectrum):
        if self.log_fe is None:
            self.initialize(spectrum)
        ts = self._template_set(self.fe_templ, self.log_fe, self.templ_width,
                                self.inputpars["fe_kernel_size_sigma"], keep)
        self._rebin_operators(ts, self.log_fe, spectrum, keep)
        desc.fe = ts
