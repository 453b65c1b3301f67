"""Walker sharding across the GPUs of one box.

The reference's only parallelism is emcee's task farm (``pool.map`` over
pickled (Model, params) tasks: multiprocessing or MPIPool, spamm/Model.py:267-286).
Here every rank holds the full replicated sampler state and an identical
counter-based RNG, so proposals are generated redundantly and *no parameters
travel*; each rank evaluates its contiguous slice of the active half-ensemble
and the only exchange is one all-gather of per-walker ln-probabilities per
stretch half-step (NCCL over NVLink on GPUs; gloo in the CPU tests).
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


class P2PLnPost(object):
    """ShardedLnPost with the exchange fused into the compute kernel.

    Every rank owns a symmetric (peer-mapped) result buffer; the fused walker kernel stores each
    ln-posterior directly into all peers' buffers over NVLink (`spamm_lnpost_batch_p2p`), so the
    only thing left after the kernel is a cross-rank barrier on the symmetric-memory signal pads --
    no all-gather copy.  Two buffers alternate so a fast rank's next half-step never overwrites
    values a slower rank is still reading (each step ends in a barrier, so ranks are never more
    than one step apart)."""

    def __init__(self, plan, max_n, group=None):
        import torch.distributed._symmetric_memory as symm
        self.plan = plan
        self.rank, self.world = world_info()
        self.group = group if group is not None else dist.group.WORLD
        self.max_n = int(max_n)
        self.bufs, self.hdls = [], []
        for _ in range(2):
            t = symm.empty(self.max_n, dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
            t.fill_(float("nan"))
            h = symm.rendezvous(t, self.group)
            self.bufs.append(t)
            self.hdls.append(h)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.step = 0

    @staticmethod
    def available():
        try:
            import torch.distributed._symmetric_memory  # noqa: F401
            return torch.cuda.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"
        except Exception:
            return False

    def __call__(self, params):
        n = params.shape[0]
        if n > self.max_n:
            raise ValueError("P2PLnPost buffer holds %d walkers, got %d" % (self.max_n, n))
        k = self.step & 1
        self.step += 1
        buf, hdl = self.bufs[k], self.hdls[k]
        lo, hi, _ = shard_bounds(n, self.world, self.rank)
        if hi > lo:
            self.plan.lnposterior_p2p(params[lo:hi].contiguous(), hdl.buffer_ptrs_dev, self.world, lo)
        hdl.barrier(channel=0)
        return buf[:n]
