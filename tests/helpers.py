"""Shared helpers for the parity tests (fixtures -> models)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Case(object):
    """One ln-posterior fixture case from tests/golden/lnpost.npz."""

    def __init__(self, z, name):
        g = lambda k: z["%s/%s" % (name, k)]  # noqa: E731
        self.name = name
        self.wl, self.flux, self.err = g("wl"), g("flux"), g("err")
        self.params, self.lnpost = g("params"), g("lnpost")
        self.lo, self.hi = g("lo"), g("hi")
        self.comps = [str(c) for c in g("comps")]
        self.bc, self.bpc = bool(g("bc")), bool(g("bpc"))


def lnpost_cases():
    z = np.load(os.path.join(GOLDEN, "lnpost.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return z, names


def rel_err(a, b):
    """Relative error with -inf == -inf and 0 == 0 counting as exact."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    out = np.zeros(a.shape)
    same = (a == b)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.abs(a - b) / np.abs(b)
    out[~same] = r[~same]
    out[np.isnan(out)] = np.inf
    return out


def grid(P):
    return 1000.0 + np.arange(P) * (9000.0 / P)


NC_TRUTH = [5e-15, 2.3]
FE_TRUTH = [1.07988504e-14, 6.91877436e-15, 5e-15, 5450.]
HG_TRUTH = [5e-17, 5.5e-17, 5e-16, 1e-16, 2e-16, 3e-16, 4e-16, 515.]
BC_TRUTH = [3e-14, 50250., 1.0, 0.0, 5050., 5.5]
TRUTH = {"PL": NC_TRUTH, "FE": FE_TRUTH, "HOST": HG_TRUTH, "BC": BC_TRUTH}
