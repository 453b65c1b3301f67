#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_multi_gpu.py
Sharded ln-posterior + device sampler over N GPUs (CUDA IPC exchange regions, flags raised by
the walker kernel): every rank must hold the WHOLE gathered vector, bit-identical to the single-rank
evaluation; every rank must end with the identical chain, and it must equal the chain one GPU produces
with the same seed -- also when each rank was handed a different seed and start state (rank 0's win)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spamm_b200.distributed import init_from_env, ShardedLnPost, ExchangeLnPost  # noqa: E402
from spamm_b200.synth import full_model_workload  # noqa: E402
from spamm_b200.sampler import EnsembleSampler  # noqa: E402
from spamm_b200.Model import ln_posterior  # noqa: E402

rank, world, local_rank = init_from_env("nccl")
torch.cuda.set_device(local_rank)
model, params = full_model_workload(n_pix=2048, n_walkers=256, seed=42)
plan = model.plan
d = torch.from_numpy(params).cuda()
full = plan.lnposterior(d).cpu().numpy()
sharded = ShardedLnPost(plan.lnposterior)(d).cpu().numpy()
assert np.array_equal(full, sharded), "all-gather sharded evaluation differs from the single-rank one"
ev = ExchangeLnPost(plan, max_n=256)
for n in (256, 250, 3):
    for _ in range(3):
        g = ev(d[:n].contiguous()).cpu().numpy()
        assert np.array_equal(g, full[:n]), "exchanged vector differs on rank %d (n=%d)" % (rank, n)
gh = plan.lnposterior_sharded_host(params, ev.exchange)
assert np.array_equal(gh, full), "host-buffer sharded call differs"
ev.exchange.close()

# each rank draws its own seed and start state: rank 0's must win
rng = np.random.RandomState(100 + rank)
p0 = params * (1.0 + 1e-3 * rng.randn(*params.shape)) if rank else params
s = EnsembleSampler(nwalkers=256, dim=20, lnpostfn=ln_posterior, args=[model], seed=7 if rank == 0 else None)
if rank == 0:
    print("exchange:", s.exchange)
s.run_mcmc(p0, 30)
chain = torch.from_numpy(s.chain).cuda()
ref = chain.clone()
if world > 1:
    dist.broadcast(ref, src=0)
assert torch.equal(ref, chain), "ranks disagree on the chain"
s.close()

# single-rank evaluation of the same chain (no sharding) for comparison
s1 = EnsembleSampler(nwalkers=256, dim=20, lnpostfn=ln_posterior, args=[model], seed=7)
s1.exchange, s1.world = "single", 1
s1._evaluate = plan.lnposterior
s1.run_mcmc(params, 30)
assert np.array_equal(s1.chain, s.chain), "sharded chain differs from the unsharded chain"
assert np.array_equal(s1.lnprobability, s.lnprobability)
if rank == 0:
    print("multi-gpu check ok: world=%d acceptance=%.3f" % (world, s.acceptance_fraction.mean()))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
