"""Template / atomic-data loading.

The reference reads ASCII templates from a directory given in the parameter
file (FeComponent.py:75-91, HostGalaxyComponent.py:90-104) and un-pickles 48
scipy interp2d objects on every flux call (BalmerContinuumCombined.py:288).
Here both a directory (same file format) and the bundled, losslessly re-packed
``data/spamm_data.npz`` (tools/pack_reference_data.py) are accepted; the SH95
table is loaded once."""
import glob
import os

import numpy as np

from .utils.parse_pars import BUNDLED

_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "spamm_data.npz")
_cache = {}


def _bundle():
    if "npz" not in _cache:
        if not os.path.exists(_NPZ):
            raise IOError("bundled template data %s not found" % _NPZ)
        _cache["npz"] = np.load(_NPZ)
    return _cache["npz"]


def load_template_dir(path, prefix, sort):
    """List of (wavelengths, flux) float64 arrays.

    prefix: "fe" | "host" selects the bundled family when path is the BUNDLED
    sentinel.  sort=True reproduces FeComponent's sorted(glob(...)); the host
    component does not sort (HostGalaxyComponent.py:93), so for a directory the
    order is whatever glob yields -- exactly the reference's behaviour -- and
    for the bundle it is the order recorded when the bundle was packed."""
    out = []
    if path == BUNDLED or path is None:
        z = _bundle()
        for i in range(len(z["%s_names" % prefix])):
            out.append((np.array(z["%s_wl_%d" % (prefix, i)], dtype=np.float64),
                        np.array(z["%s_flux_%d" % (prefix, i)], dtype=np.float64)))
        return out
    files = glob.glob(os.path.join(path, "*"))
    if sort:
        files = sorted(files)
    for f in files:
        wl, fx = np.loadtxt(f, unpack=True)
        out.append((np.asarray(wl, dtype=np.float64), np.asarray(fx, dtype=np.float64)))
    return out


def template_names(prefix):
    return [str(s) for s in _bundle()["%s_names" % prefix]]


def load_sh95(path=BUNDLED):
    """(ne[7], T[10], z[48,7,10]); list index k is the reference's coeff[k]."""
    if path == BUNDLED or path is None:
        z = _bundle()
        return (np.array(z["sh95_ne"], dtype=np.float64), np.array(z["sh95_T"], dtype=np.float64),
                np.array(z["sh95_z"], dtype=np.float64))
    z = np.load(path)
    return (np.array(z["sh95_ne"], dtype=np.float64), np.array(z["sh95_T"], dtype=np.float64),
            np.array(z["sh95_z"], dtype=np.float64))
