"""Walker sharding across the GPUs of one box.

The reference's only parallelism is emcee's task farm (``pool.map`` over
pickled (Model, params) tasks: multiprocessing or MPIPool, spamm/Model.py:267-286).
Here every rank holds the full replicated sampler state and an identical
counter-based RNG, so proposals are generated redundantly and *no parameters
travel*; each rank evaluates its contiguous slice of the active half-ensemble
and the only exchange is the per-walker ln-probabilities of each stretch half-step.
On GPUs the walker kernel leaves them in its rank's exchange region and raises a flag
in every rank's region over NVLink; consumers wait on the flags on the device and pull
the other ranks' entries (`Exchange`, `ExchangeLnPost`; include/spamm_b200.h
"multi-GPU"); `ShardedLnPost` is the collective form of the same contract
(all-gather: NCCL, or gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK,
    WORLD_SIZE, MASTER_ADDR, MASTER_PORT).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, world, rank):
    """Contiguous slice [lo, hi) of n walkers for `rank`; every rank gets the
    same padded length ceil(n / world) so one fixed-size all-gather suffices."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    hi = min(lo + per, n)
    return lo, hi, per


def replicate_object(obj, group=None):
    """Rank 0's picklable object on every rank (the sampler's seed: each rank's own numpy RNG would draw a
    different one and the ranks would silently propose different moves)."""
    if world_info()[1] == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def replicate_tensor(t, group=None):
    """Overwrite `t` on every rank with rank 0's values (the sampler's start state); returns t."""
    if world_info()[1] > 1:
        dist.broadcast(t, src=0, group=group)
    return t


class ShardedLnPost(object):
    """ln-posterior of [n, D] parameters evaluated as one slice per rank.

    local_eval: callable mapping a [m, D] tensor to a [m] tensor (on a GPU rank
    this is ModelPlan.lnposterior).  The result, identical on every rank, is the
    full [n] vector."""

    def __init__(self, local_eval, group=None):
        self.local_eval = local_eval
        self.group = group
        self.rank, self.world = world_info()
        self._buf = None

    def __call__(self, params):
        n = params.shape[0]
        if self.world == 1:
            return self.local_eval(params)
        lo, hi, per = shard_bounds(n, self.world, self.rank)
        local = torch.full((per,), float("nan"), dtype=torch.float64, device=params.device)
        if hi > lo:
            local[:hi - lo] = self.local_eval(params[lo:hi].contiguous())
        if self._buf is None or self._buf.numel() != per * self.world or self._buf.device != params.device:
            self._buf = torch.empty(per * self.world, dtype=torch.float64, device=params.device)
        dist.all_gather_into_tensor(self._buf, local, group=self.group)
        return self._buf[:n] if per * self.world == n else self._gather_unpadded(n, per)

    def _gather_unpadded(self, n, per):
        parts = []
        for r in range(self.world):
            lo, hi, _ = shard_bounds(n, self.world, r)
            parts.append(self._buf[r * per:r * per + (hi - lo)])
        return torch.cat(parts)


class Exchange(object):
    """One rank's end of the library's device-side result exchange (`SpammExchange`, include/spamm_b200.h).

    The library allocates the region (flag block + two result buffers), this wrapper all-gathers the
    64-byte CUDA IPC handles over torch.distributed and lets the library map every peer's region; from
    then on the data path has no collective and no host barrier: the walker kernel raises a flag in every
    rank's region over NVLink, consumers wait on the device and pull the other ranks' results.
    All ranks must make the same sequence of sharded calls."""

    def __init__(self, max_walkers, group=None):
        import ctypes
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world = world_info()
        self.group = group
        self.max_walkers = int(max_walkers)
        handle = (ctypes.c_ubyte * _lib.IPC_HANDLE_BYTES)()
        h = ctypes.c_void_p()
        _lib.check(self.lib.spamm_exchange_create(self.world, self.rank, self.max_walkers, ctypes.byref(h),
                                                  ctypes.cast(handle, ctypes.c_void_p)))
        self._h = h
        if self.world > 1:
            mine = torch.tensor(list(handle), dtype=torch.uint8)
            if dist.get_backend(group) == "nccl":
                mine = mine.cuda()
            every = torch.empty(self.world * _lib.IPC_HANDLE_BYTES, dtype=torch.uint8, device=mine.device)
            dist.all_gather_into_tensor(every, mine, group=group)
            blob = every.cpu().numpy().tobytes()
            _lib.check(self.lib.spamm_exchange_connect(self._h, ctypes.c_char_p(blob)))
            torch.cuda.synchronize()
            dist.barrier(group=group)            # every region is mapped everywhere before anyone stores

    @property
    def epoch(self):
        return int(self.lib.spamm_exchange_epoch(self._h))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            if self.world > 1 and dist.is_initialized():
                torch.cuda.synchronize()
                dist.barrier(group=self.group)   # no peer still stores into a region that is about to go
            self.lib.spamm_exchange_destroy(h)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self.lib.spamm_exchange_destroy(h)
            except Exception:
                pass


class ExchangeLnPost(object):
    """ShardedLnPost with the exchange fused into the compute kernel (`spamm_lnpost_batch_sharded`):
    [n, D] replicated parameters -> the full [n] vector on every rank."""

    def __init__(self, plan, max_n, group=None, exchange=None):
        self.plan = plan
        self.exchange = exchange if exchange is not None else Exchange(max_n, group=group)
        self.rank, self.world = self.exchange.rank, self.exchange.world

    @staticmethod
    def available():
        return torch.cuda.is_available() and (not dist.is_initialized() or dist.get_backend() == "nccl")

    def __call__(self, params, out=None):
        return self.plan.lnposterior_sharded(params, self.exchange, out=out)
