#!/usr/bin/env python
"""Three launches of the fused kernel on config 5 (for ncu -k regex:k_walkers -s 2 -c 1)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200.synth import full_model_workload  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
model, params = full_model_workload(n_pix=8192, n_walkers=W, seed=42)
p = torch.from_numpy(np.asarray(params)).cuda()
out = torch.empty(W, dtype=torch.float64, device="cuda")
for _ in range(3):
    model.plan.lnposterior(p, out=out)
torch.cuda.synchronize()
print("ok", float(out.sum().cpu()))
