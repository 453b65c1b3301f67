#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/exchange_bench.py
Cost of the device-side exchange on one half-ensemble (4096 walkers of config 5 sharded N ways): the plain
local call on this rank's slice vs the sharded call (results to every rank + flag wait), back to back."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200.distributed import init_from_env, Exchange, shard_bounds  # noqa: E402
from spamm_b200.synth import full_model_workload  # noqa: E402

rank, world, local_rank = init_from_env("nccl")
torch.cuda.set_device(local_rank)
model, params = full_model_workload(n_pix=8192, n_walkers=8192, seed=42)
plan = model.plan
NH = int(os.environ.get("HALF", "4096"))
half = torch.from_numpy(params[:NH]).cuda()
lo, hi, per = shard_bounds(NH, world, rank)
mine = half[lo:hi].contiguous()
out = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
ex = Exchange(8192)


def timed(fn, reps=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


t_local = timed(lambda: plan.lnposterior(mine, out=out))
t_shard = timed(lambda: plan.lnposterior_sharded(half, ex, copy=False))
if rank == 0:
    print("world %d, %d walkers per GPU (tail wait=%s): local call %.1f us, sharded call with exchange %.1f us "
          "(+%.1f us)" % (world, hi - lo, os.environ.get("SPAMM_EXCHANGE_TAILWAIT", "1"), t_local, t_shard,
                          t_shard - t_local))
ex.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
