#!/usr/bin/env python
"""Pack the reference's *data inputs* (not source) into one npz the GPU box can read.

The reference loads its inputs at run time from paths relative to
``examples/`` (utils/parameters.yaml:20,66; BalmerContinuumCombined.py:288):
  Data/FeModels/*              3 Fe II/III templates  (FeComponent.py:75-91)
  Data/HostModels/*            7 host templates       (HostGalaxyComponent.py:90-104)
  Data/SH95recombcoeff/coeff.interpers.pickle
                               48 pickled scipy interp2d (kx=ky=1) objects,
                               Storey & Hummer 1995 case-B line emissivities
                               (BalmerContinuumCombined.py:288-290)
``/root/reference`` does not exist on the GPU box, so the raw numbers are
re-packed here, losslessly, as float64 arrays:

  fe_names, fe_wl_<i>, fe_flux_<i>       in the reference's sorted() order
  host_names, host_wl_<i>, host_flux_<i> in the glob order observed in the
                                          build container
  sh95_ne[7], sh95_T[10], sh95_z[48,7,10] list index k <-> line n = 50-k ...
                                          (coef_use[-i+2], BC:294: index 47 = n=3)

Run once in the build container:  python tools/pack_reference_data.py
"""
import glob
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402


def main():
    ref_root = refshim.REFERENCE_ROOT
    out = {}
    fe_files = sorted(glob.glob(os.path.join(ref_root, "Data/FeModels", "*")))
    out["fe_names"] = np.array([os.path.basename(f) for f in fe_files])
    for i, f in enumerate(fe_files):
        wl, fx = np.loadtxt(f, unpack=True)
        out["fe_wl_%d" % i] = wl
        out["fe_flux_%d" % i] = fx
    # the reference does NOT sort host templates (HostGalaxyComponent.py:93)
    host_files = glob.glob(os.path.join(ref_root, "Data/HostModels", "*"))
    out["host_names"] = np.array([os.path.basename(f) for f in host_files])
    for i, f in enumerate(host_files):
        wl, fx = np.loadtxt(f, unpack=True)
        out["host_wl_%d" % i] = wl
        out["host_flux_%d" % i] = fx
    refshim.install()
    with open(os.path.join(ref_root, "Data/SH95recombcoeff/coeff.interpers.pickle"), "rb") as fh:
        coeff = pickle.load(fh, encoding="latin1")
    ne = np.asarray(coeff[0].x, dtype=np.float64)
    T = np.asarray(coeff[0].y, dtype=np.float64)
    z = np.zeros((len(coeff), len(ne), len(T)))
    for k, ci in enumerate(coeff):
        assert np.array_equal(ci.x, ne) and np.array_equal(ci.y, T)
        assert ci.tck[3] == 1 and ci.tck[4] == 1
        zz = np.asarray(ci.tck[2], dtype=np.float64).reshape(len(ne), len(T))
        # kx=ky=1 with knots on the data grid: spline coefficients == samples
        for a in range(len(ne)):
            for b in range(len(T)):
                assert abs(ci(ne[a], T[b])[0] - zz[a, b]) <= 1e-14 * abs(zz[a, b])
        z[k] = zz
    out["sh95_ne"], out["sh95_T"], out["sh95_z"] = ne, T, z
    dst = os.path.join(ROOT, "spamm_b200", "data", "spamm_data.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")
    print("fe:", list(out["fe_names"]), [len(out["fe_wl_%d" % i]) for i in range(len(fe_files))])
    print("host:", list(out["host_names"]))


if __name__ == "__main__":
    main()
