"""CPU: the numpy oracle against the reference-made fixtures for BASELINE config 1 on all six
``fakepowlaw{1..6}_werr.dat`` spectra (tests/golden/fakepowlaw.npz, made by make_golden_fake.py)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, rel_err
from oracle.spamm_numpy import OracleModel, DEFAULT_PARS


def _pars(norm_max):
    pars = {k: dict(v) for k, v in DEFAULT_PARS.items()}
    pars["nuclear_continuum"]["pl_norm_max"] = norm_max
    return pars


@pytest.mark.parametrize("i", range(1, 7))
def test_oracle_config1_on_every_fake_spectrum(i):
    z = np.load(os.path.join(GOLDEN, "fakepowlaw.npz"))
    g = lambda k: z["f%d/%s" % (i, k)]  # noqa: E731
    for box, key in (("fnw", "lnpost"), ("max_flux", "lnpost_maxflux")):
        o = OracleModel(g("wl"), g("flux"), g("err"), comps=["PL"], pars=_pars(box))
        lo, hi = o.bounds()
        want_hi = g("hi")[0] if box == "fnw" else float(g("hi_maxflux"))
        assert lo[0] == g("lo")[0] and hi[0] == want_hi
        got = np.array([o.ln_posterior(p) for p in g("params")])
        assert rel_err(got, g(key)).max() < 1e-12
        assert np.array_equal(np.isneginf(got), np.isneginf(g(key)))
    # the spreadsheet truth lies outside the default ("fnw") prior box on all six spectra (App. B Q12)
    assert np.isneginf(g("lnpost")[-1]) and np.isfinite(g("lnpost_maxflux")[-1])
