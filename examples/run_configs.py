#!/usr/bin/env python
"""The five BASELINE.json configurations end to end, the way the reference's examples/run_nc.py,
run_fe.py, run_hg.py, run_bc.py and test_spamm.py drive them: synthetic spectrum from the known
parameters (examples/test_spamm.py:53-109), spamm() (spamm/run_spamm.py:26-167), posterior medians with
the 16/84 percentiles (spamm/Samples.py:14-88) next to the truths.  Needs a CUDA device.

    python examples/run_configs.py --config 1            # NC on a FakeData-sized grid, 32 walkers
    python examples/run_configs.py --config 5 --iterations 300
    python examples/run_configs.py --all --quick
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200.run_spamm import spamm  # noqa: E402
from spamm_b200.Samples import Samples, median_values  # noqa: E402
from spamm_b200.synth import synthetic_spectrum, config_grid, NC_PARAMS, FE_PARAMS, HG_PARAMS, BC_PARAMS  # noqa: E402

# (components, pixels, walkers) as BASELINE.json lists them; config 1's grid has the size and range of
# Data/FakeData/PLcompOnly/fakepowlaw1_werr.dat (845 px, 1159-2003 A, 1 A steps)
CONFIGS = {
    1: (("PL",), None, 32),
    2: (("PL", "FE"), 4096, 1024),
    3: (("PL", "HOST"), 8192, 2048),
    4: (("PL", "BC"), 8192, 4096),
    5: (("PL", "FE", "HOST", "BC"), 8192, 8192),
}


def truths(components):
    out = []
    for c in components:
        if c == "PL":
            out += [NC_PARAMS["norm_PL"], NC_PARAMS["slope1"]]
        elif c == "FE":
            out += [FE_PARAMS["fe_norm_1"], FE_PARAMS["fe_norm_2"], FE_PARAMS["fe_norm_3"], FE_PARAMS["fe_width"]]
        elif c == "HOST":
            out += [HG_PARAMS["hg_norm_%d" % (i + 1)] for i in range(HG_PARAMS["no_templates"])]
            out += [HG_PARAMS["hg_stellar_disp"]]
        elif c == "BC":
            out += [BC_PARAMS[k] for k in ("bc_norm", "bc_Te", "bc_tauBE", "bc_loffset", "bc_lwidth", "bc_logNe")]
    return np.array(out, dtype=np.float64)


def run(config, n_iterations, n_walkers=None, seed=1):
    components, n_pix, walkers = CONFIGS[config]
    walkers = n_walkers or walkers
    wl = 1159.0 + np.arange(845.0) if n_pix is None else config_grid(n_pix)
    wl, flux, err = synthetic_spectrum(components, wl)
    rng = np.random.RandomState(seed)
    flux = flux + err * rng.standard_normal(len(wl))          # one noise realisation of the 5 % errors
    outdir = tempfile.mkdtemp(prefix="spamm_b200_cfg%d_" % config)
    t0 = time.time()
    # "BC" alone switches on the Balmer continuum only (run_spamm.py:118-121); the BASELINE configurations
    # fit the continuum together with the high-order lines, like create_bc generates them
    complist = [x for c in components for x in (("BC", "BPC") if c == "BC" else (c,))]
    out = spamm(complist, (wl, flux, err), n_walkers=walkers, n_iterations=n_iterations,
                outdir=outdir, picklefile="cfg%d" % config, seed=seed)
    dt = time.time() - t0
    model = out["model"]
    print("config %d: %s, %d px, %d walkers x %d iterations in %.2f s (%.3g walker-steps/s), acceptance %.2f"
          % (config, "+".join(components), len(wl), walkers, n_iterations, dt, walkers * n_iterations / dt,
             float(np.mean(model.sampler.acceptance_fraction))))
    s = Samples(os.path.join(outdir, "cfg%d.pickle.gz" % config), burn=n_iterations // 2)
    tv = truths(components)
    stats = median_values(s.samples)                          # median, -err, +err (analysis.py:514-540)
    for i, name in enumerate(s.model_parameter_names):
        print("  %-16s truth %12.5g   median %12.5g  (-%.3g +%.3g)" % (name, tv[i], stats[i, 0], stats[i, 1], stats[i, 2]))
    return s, stats, tv


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", type=int, choices=sorted(CONFIGS), default=1)
    ap.add_argument("--all", action="store_true", help="run configs 1-5 in turn")
    ap.add_argument("--iterations", type=int, default=500)
    ap.add_argument("--walkers", type=int, default=None, help="override the config's walker count")
    ap.add_argument("--quick", action="store_true", help="50 iterations, at most 256 walkers")
    args = ap.parse_args()
    for cfg in (sorted(CONFIGS) if args.all else [args.config]):
        n_it, n_w = args.iterations, args.walkers
        if args.quick:
            n_it, n_w = 50, min(CONFIGS[cfg][2], 256)
        run(cfg, n_it, n_w)


if __name__ == "__main__":
    main()
