"""world_size-2 gloo test of the walker sharding + ln-prob all-gather used by the
sampler (spamm_b200/distributed.py).  CPU only; the per-rank evaluator is a plain
closed-form function standing in for ModelPlan.lnposterior."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spamm_b200.distributed import ShardedLnPost, shard_bounds, replicate_object, replicate_tensor


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fn(p):
    return -0.5 * (p ** 2).sum(dim=1) + p[:, 0]


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def local_eval(p):
        calls.append(p.shape[0])
        return _fn(p)

    ev = ShardedLnPost(local_eval)
    g = torch.Generator().manual_seed(0)            # replicated state: same on every rank
    params = torch.randn(n, 5, dtype=torch.float64, generator=g)
    res = ev(params)
    res2 = ev(params[: n - 3])                        # ragged second call
    np.save(os.path.join(out_dir, "r%d.npy" % rank), res.numpy())
    np.save(os.path.join(out_dir, "s%d.npy" % rank), res2.numpy())
    np.save(os.path.join(out_dir, "c%d.npy" % rank), np.array(calls))
    # the sampler's replicated state: every rank draws its own seed and start positions, rank 0's must win
    seed = replicate_object(1000 + rank)
    start = replicate_tensor(torch.full((4, 3), float(rank + 1), dtype=torch.float64))
    np.save(os.path.join(out_dir, "seed%d.npy" % rank), np.array([seed]))
    np.save(os.path.join(out_dir, "start%d.npy" % rank), start.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 37])
def test_sharded_lnpost_world2(tmp_path, n):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    params = torch.randn(n, 5, dtype=torch.float64, generator=g)
    want = _fn(params).numpy()
    for r in range(world):
        got = np.load(tmp_path / ("r%d.npy" % r))
        assert np.array_equal(got, want)               # identical on every rank
        got2 = np.load(tmp_path / ("s%d.npy" % r))
        assert np.array_equal(got2, want[: n - 3])
        calls = np.load(tmp_path / ("c%d.npy" % r))
        lo, hi, _ = shard_bounds(n, world, r)
        assert calls[0] == hi - lo                     # each rank evaluated only its slice
        assert np.load(tmp_path / ("seed%d.npy" % r))[0] == 1000
        assert np.array_equal(np.load(tmp_path / ("start%d.npy" % r)), np.full((4, 3), 1.0))


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi, per = shard_bounds(n, world, r)
                assert hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(n))
