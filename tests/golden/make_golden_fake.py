#!/usr/bin/env python
"""Generate tests/golden/fakepowlaw.npz from the UNMODIFIED reference (build container only).

BASELINE config 1 runs on ``Data/FakeData/PLcompOnly/fakepowlaw{1..6}_werr.dat``; the truths the fake spectra
were made with are in ``Data/FakeData/fakedata_properties_v3.xlsx`` (sheet 1, rows 6-11: power-law index alpha,
F_lambda at lambda_0, lambda_0).  For every file this stores

  wl, flux, err        the three columns of the *_werr.dat file
  alpha, f0, lam0      the spreadsheet truths (the component's slope1 is -alpha)
  params, lnpost       16 walkers drawn like Model.run_mcmc (seed 42) + the truth vector, and the reference's
                       ln_posterior for each (default parameters.yaml: pl_norm_max = "fnw")
  lo, hi               the prior box the reference resolved
  hi_maxflux           norm_max the reference resolves for pl_norm_max = "max_flux" (the chain test's box;
                       with the default "fnw" -- the noisy flux at lambda_0 -- all six truths fall OUTSIDE the
                       prior box)
  lnpost_maxflux       the reference's ln_posterior of the same 17 vectors in that wider box

    python tests/golden/make_golden_fake.py
"""
import os
import sys
import warnings
import zipfile
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import refshim  # noqa: E402

ref = refshim.load()
DATA = os.path.join(refshim.REFERENCE_ROOT, "Data", "FakeData")


def xlsx_truths():
    z = zipfile.ZipFile(os.path.join(DATA, "fakedata_properties_v3.xlsx"))
    ns = {"m": "http://schemas.openxmlformats.org/spreadsheetml/2006/main"}
    shared = ["".join(t.itertext()) for t in ET.fromstring(z.read("xl/sharedStrings.xml")).findall("m:si", ns)]
    out = {}
    for row in ET.fromstring(z.read("xl/worksheets/sheet1.xml")).find("m:sheetData", ns).findall("m:row", ns):
        if not 6 <= int(row.get("r")) <= 11:
            continue
        cells = {}
        for c in row.findall("m:c", ns):
            v = c.find("m:v", ns)
            if v is not None:
                cells[c.get("r")[0]] = shared[int(v.text)] if c.get("t") == "s" else v.text
        idx = int(cells["A"].replace("fakepowlaw", "").replace("_deg.dat", ""))
        out[idx] = (float(cells["F"]), float(cells["G"]), float(cells["H"]))
    return out


def nc_model(wl, flux, err, norm_max=None):
    with refshim.chdir_examples():
        from utils.parse_pars import parse_pars
        pars = parse_pars()["nuclear_continuum"]
        if norm_max is not None:
            pars["pl_norm_max"] = norm_max
        m = ref.Model()
        m.components.append(ref.NuclearContinuumComponent(broken=False, pars=pars))
        m.data_spectrum = ref.Spectrum(wl, flux, err)
    return m


def main():
    truths = xlsx_truths()
    out = {}
    for i in range(1, 7):
        wl, flux, err = np.loadtxt(os.path.join(DATA, "PLcompOnly", "fakepowlaw%d_werr.dat" % i), unpack=True)
        alpha, f0, lam0 = truths[i]
        m = nc_model(wl, flux, err)
        np.random.seed(42)
        walkers = [list(m.components[0].initial_values(m.data_spectrum)) for _ in range(16)]
        params = np.array(walkers + [[f0, -alpha]], dtype=np.float64)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            lnpost = np.array([ref.ln_posterior(p, m) for p in params])
        c = m.components[0]
        m2 = nc_model(wl, flux, err, norm_max="max_flux")
        m2.components[0].initial_values(m2.data_spectrum)
        with contextlib.redirect_stdout(io.StringIO()):
            lnpost2 = np.array([ref.ln_posterior(p, m2) for p in params])
        k = "f%d/" % i
        out[k + "wl"], out[k + "flux"], out[k + "err"] = wl, flux, err
        out[k + "alpha"], out[k + "f0"], out[k + "lam0"] = alpha, f0, lam0
        out[k + "params"], out[k + "lnpost"] = params, lnpost
        out[k + "lo"] = np.array([c.norm_min, c.slope_min], dtype=np.float64)
        out[k + "hi"] = np.array([c.norm_max, c.slope_max], dtype=np.float64)
        out[k + "hi_maxflux"] = float(m2.components[0].norm_max)
        out[k + "lnpost_maxflux"] = lnpost2
        print("file %d: %d px, truth (%.3g, %.2f) at %.2f A (median wl %.2f), norm_max fnw %.4g / max_flux %.4g, "
              "lnpost(truth) %.6g / %.10g" % (i, len(wl), f0, -alpha, lam0, np.median(wl), c.norm_max,
                                              m2.components[0].norm_max, lnpost[-1], lnpost2[-1]))
    np.savez_compressed(os.path.join(HERE, "fakepowlaw.npz"), **out)


if __name__ == "__main__":
    main()
