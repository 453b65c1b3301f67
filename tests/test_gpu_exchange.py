"""Multi-GPU exchange through the C ABI, exercised on ONE device: the library's `SpammExchange` is
plain device memory + flags, so two ranks can live in one process as two regions on the same GPU,
each driven on its own stream (`spamm_exchange_connect_ptrs`).  The walker kernel of each "rank"
stores its slice into both regions and raises its flag; the consumers (k_wait_flags, the accept
kernel) spin on the device until both flags are up.  Everything must be bit-identical to the
unsharded path.  (The multi-process form over CUDA IPC is checked by tools/check_multi_gpu.py and by
bench.py itself whenever it runs on more than one GPU.)"""
import ctypes

import numpy as np
import pytest

from helpers import Case, lnpost_cases

pytestmark = pytest.mark.gpu
vp = ctypes.c_void_p

# Several ranks in ONE process need the kernels of different streams to run at the same time (a consumer spins
# until another stream's producer has raised its flag).  One context and a handful of CTAs: that holds on every
# B200 these tests have run on, and a device-side time-out (5 s) turns a violation into a failure instead of a
# hang -- but nothing in CUDA guarantees it, so they only run when asked for:
#     SPAMM_TEST_INPROCESS_RANKS=1 python -m pytest tests/test_gpu_exchange.py -m gpu
# The multi-process form (one rank per GPU over CUDA IPC) is asserted by bench.py itself whenever it runs on more
# than one GPU (whole gathered vector on every rank, identical chains) and by tools/check_multi_gpu.py.
import os  # noqa: E402
inprocess_ranks = pytest.mark.skipif(os.environ.get("SPAMM_TEST_INPROCESS_RANKS") != "1",
                                     reason="set SPAMM_TEST_INPROCESS_RANKS=1 (needs concurrent kernels of one process)")


def _model(n_walkers=256, n_pix=2048):
    from spamm_b200.synth import full_model_workload
    return full_model_workload(n_pix=n_pix, n_walkers=n_walkers, seed=42)


class _Rank(object):
    """One in-process rank: its exchange handle and stream."""

    def __init__(self, lib, n_ranks, rank, max_n):
        import torch
        from spamm_b200 import _lib
        self.lib, self.rank = lib, rank
        self.h = vp()
        _lib.check(lib.spamm_exchange_create(n_ranks, rank, max_n, ctypes.byref(self.h), None))
        self.stream = torch.cuda.Stream()

    def connect(self, ptrs):
        from spamm_b200 import _lib
        arr = (ctypes.c_uint64 * len(ptrs))(*ptrs)
        _lib.check(self.lib.spamm_exchange_connect_ptrs(self.h, arr))

    def sptr(self):
        return vp(self.stream.cuda_stream)


def _ranks(n_ranks, max_n):
    from spamm_b200 import _lib
    lib = _lib.load()
    rs = [_Rank(lib, n_ranks, r, max_n) for r in range(n_ranks)]
    ptrs = [int(lib.spamm_exchange_region_ptr(r.h)) for r in rs]
    if n_ranks > 1:
        for r in rs:
            r.connect(ptrs)
    return lib, rs


def test_shard_bounds_match_python():
    from spamm_b200 import _lib
    from spamm_b200.distributed import shard_bounds
    lib = _lib.load()
    for n in (1, 7, 37, 4096, 4097):
        for world in (1, 2, 3, 8):
            for r in range(world):
                lo, hi = ctypes.c_int64(), ctypes.c_int64()
                lib.spamm_shard_bounds(n, world, r, ctypes.byref(lo), ctypes.byref(hi))
                assert (lo.value, hi.value) == shard_bounds(n, world, r)[:2]


def test_single_rank_exchange_equals_plain_call():
    import torch
    from spamm_b200.distributed import Exchange
    model, params = _model()
    d = torch.from_numpy(params).cuda()
    want = model.plan.lnposterior(d).cpu().numpy()
    ex = Exchange(max_walkers=len(params))
    for _ in range(3):                          # both buffers, several epochs
        got = model.plan.lnposterior_sharded(d, ex).cpu().numpy()
        assert np.array_equal(got, want)
    got_h = model.plan.lnposterior_sharded_host(params, ex)
    assert np.array_equal(got_h, want)
    assert ex.epoch == 4
    ex.close()


@inprocess_ranks
@pytest.mark.parametrize("n_ranks,n", [(2, 256), (3, 250), (4, 3)])
def test_in_process_ranks_exchange_full_vector(n_ranks, n):
    """Every rank ends up with the WHOLE gathered vector (not just its slice), ragged and empty shards
    included (n = 3 on 4 ranks leaves rank 3 without walkers: it only raises its flag)."""
    import torch
    from spamm_b200 import _lib
    model, params = _model()
    params = params[:n]
    d = torch.from_numpy(params).cuda()
    plan = model.plan
    want = plan.lnposterior(d).cpu().numpy()
    torch.cuda.synchronize()
    lib, rs = _ranks(n_ranks, max_n=256)
    outs = [torch.full((n,), float("nan"), dtype=torch.float64, device="cuda") for _ in rs]
    for epoch in range(3):
        for r, o in zip(rs, outs):
            _lib.check(lib.spamm_lnpost_batch_sharded(plan._h, r.h, vp(d.data_ptr()), n, vp(o.data_ptr()), None,
                                                      r.sptr()))
        torch.cuda.synchronize()
        for o in outs:
            assert np.array_equal(o.cpu().numpy(), want)
            o.fill_(float("nan"))
    assert plan.status() == 0
    for r in rs:
        lib.spamm_exchange_destroy(r.h)


@pytest.mark.parametrize("n_ranks", [1, pytest.param(2, marks=inprocess_ranks)])
def test_sharded_sampler_equals_single_device(n_ranks):
    """spamm_stretch_run_sharded (one rank; two in-process ranks) == spamm_stretch_run on one device, bit for
    bit (chain, ln-probabilities, acceptance counts), and all ranks hold the same chain."""
    import torch
    from spamm_b200 import _lib
    K, n_iter, seed = 128, 12, 11
    model, params = _model(n_walkers=K)
    plan = model.plan
    D = plan.n_dim
    dev = "cuda"

    def state():
        w = torch.from_numpy(params).to(dev)
        return dict(w=w, lnp=plan.lnposterior(w).clone(), q=torch.empty((K // 2, D), dtype=torch.float64, device=dev),
                    z=torch.empty(K // 2, dtype=torch.float64, device=dev),
                    nacc=torch.zeros(K, dtype=torch.int64, device=dev),
                    chain=torch.zeros((K, n_iter, D), dtype=torch.float64, device=dev),
                    lnpc=torch.zeros((K, n_iter), dtype=torch.float64, device=dev))

    lib = _lib.load()
    ref = state()
    lnp_q = torch.empty(K // 2, dtype=torch.float64, device=dev)
    st = vp(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.spamm_stretch_run(plan._h, vp(ref["w"].data_ptr()), vp(ref["lnp"].data_ptr()), K, 2.0, seed, 0,
                                     n_iter, vp(ref["q"].data_ptr()), vp(ref["z"].data_ptr()), vp(lnp_q.data_ptr()),
                                     vp(ref["nacc"].data_ptr()), vp(ref["chain"].data_ptr()),
                                     vp(ref["lnpc"].data_ptr()), n_iter, st))
    torch.cuda.synchronize()
    lib, rs = _ranks(n_ranks, max_n=K)
    states = [state() for _ in rs]
    torch.cuda.synchronize()
    for r, s in zip(rs, states):
        _lib.check(lib.spamm_stretch_run_sharded(plan._h, r.h, vp(s["w"].data_ptr()), vp(s["lnp"].data_ptr()), K, 2.0,
                                                 seed, 0, n_iter, vp(s["q"].data_ptr()), vp(s["z"].data_ptr()),
                                                 vp(s["nacc"].data_ptr()), vp(s["chain"].data_ptr()),
                                                 vp(s["lnpc"].data_ptr()), n_iter, r.sptr()))
    torch.cuda.synchronize()
    assert plan.status() == 0
    for s in states:
        for k in ("w", "lnp", "nacc", "chain", "lnpc"):
            assert torch.equal(s[k], ref[k]), k
    assert int(ref["nacc"].sum()) > 0
    for r in rs:
        lib.spamm_exchange_destroy(r.h)


def test_lost_peer_times_out_with_an_error():
    """A rank whose peer never shows up must get an error after the device-side time-out, not a hang
    and not a vector with holes."""
    import torch
    from spamm_b200 import _lib
    model, params = _model(n_walkers=64)
    plan = model.plan
    d = torch.from_numpy(params).cuda()
    lib, rs = _ranks(2, max_n=64)
    out = torch.empty(64, dtype=torch.float64, device="cuda")
    _lib.check(lib.spamm_lnpost_batch_sharded(plan._h, rs[0].h, vp(d.data_ptr()), 64, vp(out.data_ptr()), None,
                                              rs[0].sptr()))          # rank 1 never calls
    torch.cuda.synchronize()
    with pytest.raises(_lib.SpammError, match="peer"):
        plan.check()
    assert plan.status() == 0                     # the condition was cleared
    for r in rs:
        lib.spamm_exchange_destroy(r.h)
