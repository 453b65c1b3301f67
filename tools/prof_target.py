#!/usr/bin/env python
"""Small ncu target: a few launches of the config-5 walker kernel at the bench's launch size (4096 walkers =
one half-ensemble), one launch of its 512-walker two-part form, and the broadening kernels (3072 rows)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200 import _lib  # noqa: E402
from spamm_b200.synth import full_model_workload  # noqa: E402

model, params = full_model_workload(n_pix=8192, n_walkers=4096, seed=42)
p = torch.from_numpy(np.asarray(params)).cuda()
out = torch.empty(4096, dtype=torch.float64, device="cuda")
for _ in range(4):
    model.plan.lnposterior(p, out=out)
p512 = p[:512].contiguous()
model.plan.lnposterior(p512, out=out[:512])
torch.cuda.synchronize()
# ncu --profile-from-start off: only what follows is captured (the model set-up launches many small kernels)
torch.cuda.cudart().cudaProfilerStart()
model.plan.lnposterior(p, out=out)
model.plan.lnposterior(p512, out=out[:512])
torch.cuda.synchronize()
lib = _lib.load()
c, i, f = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
_lib.check(lib.spamm_broadening_probe(model.plan._h, 0, 30, 3072, 1, ctypes.byref(c), ctypes.byref(i), ctypes.byref(f)))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", float(out[:8].sum().cpu()), c.value, i.value)
