#!/usr/bin/env python
"""Secondary metric: full-sampler throughput (walkers x steps per second) of the device
stretch-move sampler on BASELINE config 5."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200.synth import full_model_workload  # noqa: E402
from spamm_b200.sampler import EnsembleSampler  # noqa: E402
from spamm_b200.Model import ln_posterior  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
model, params = full_model_workload(n_pix=8192, n_walkers=W, seed=42)
for native in (False, True):
    s = EnsembleSampler(nwalkers=W, dim=20, lnpostfn=ln_posterior, args=[model], seed=7, store_chain=True)
    s.native_loop = native
    s.run_mcmc(params, 5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run_mcmc(s._walkers.cpu().numpy(), N)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("native_loop=%s  W=%d N=%d: %.3f s, %.3f ms/iteration, %.3g walker-steps/s (= ln-posterior evals/s), acceptance %.3f"
          % (native, W, N, dt, 1e3 * dt / N, W * N / dt, s.acceptance_fraction.mean()))
