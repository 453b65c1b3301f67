"""Ensemble sampler with emcee 2.2.1's stretch-move semantics, resident on the
device, plus a `pool` adapter that lets an *unmodified* emcee evaluate a whole
half-ensemble in one batched GPU call.

Reference call site: spamm/Model.py:283-289 builds
``emcee.EnsembleSampler(nwalkers, dim, lnpostfn=ln_posterior, args=[model], threads=1)``
and calls ``run_mcmc(walkers_matrix, n_iterations)``; afterwards the code
reads ``sampler.chain`` [k, N, dim], ``sampler.lnprobability`` [k, N] and
``sampler.acceptance_fraction`` (run_spamm.py:133, Samples.py:75).  Those
attributes are provided here with the same shapes.

Per iteration, for (S0, S1) in [(first half, second half), (second, first)]:
    z = ((a-1) U + 1)^2 / a;  j ~ randint(|S1|);  q = c_j - z (c_j - s)
    accept iff (dim-1) ln z + lnp(q) - lnp(s) > ln U'
(emcee 2.2.1 EnsembleSampler._propose_stretch).  The RNG is a counter-based
Philox keyed by (seed, iteration, half, walker): stream-exact parity with
numpy's MT19937 is neither possible nor required; parity is statistical.
"""
import ctypes

import numpy as np

from . import _lib


class GpuPool(object):
    """`pool=` object for emcee: emcee calls ``pool.map(lnprobfn, [p_i ...])``
    once per half-step; this stacks the parameter vectors and makes ONE batched
    ln-posterior call instead of len(seq) Python calls."""

    def __init__(self, model):
        self.model = model
        self.calls = 0
        self.evaluations = 0

    def map(self, fn, seq):
        params = np.array([np.asarray(p, dtype=np.float64) for p in seq])
        if params.size == 0:
            return []
        self.calls += 1
        self.evaluations += len(params)
        return [float(v) for v in self.model.lnposterior_batch(params)]

    def close(self):
        pass

    def is_master(self):
        return True

    def wait(self):
        pass


class EnsembleSampler(object):
    """Device-resident stretch-move sampler (emcee 2.2.1 call signature).

    args[0] must be a spamm_b200 Model (or pass plan=); lnpostfn is accepted for
    signature compatibility and must be spamm_b200.Model.ln_posterior.
    """

    def __init__(self, nwalkers, dim, lnpostfn=None, a=2.0, args=(), kwargs=None, threads=1,
                 pool=None, seed=None, plan=None, store_chain=True):
        assert nwalkers % 2 == 0, "The number of walkers must be even."
        assert nwalkers >= 2 * dim, \
            "The number of walkers needs to be more than twice the dimension of your parameter space."
        self.k = nwalkers
        self.dim = dim
        self.a = float(a)
        self.args = list(args)
        if plan is None:
            if not self.args or not hasattr(self.args[0], "plan"):
                raise ValueError("EnsembleSampler needs args=[model] (a spamm_b200 Model) or plan=")
            plan = self.args[0].plan
        assert plan.n_dim == dim, "dim does not match the model's parameter count"
        self.plan = plan
        self.lib = _lib.load()
        self.seed = int(np.random.randint(0, 2 ** 31 - 1) if seed is None else seed)
        self.store_chain = store_chain
        self.native_loop = True          # single device: iterate inside the library (spamm_stretch_run)
        self.iterations = 0
        self._chain = None
        self._lnprob = None
        self._walkers = None
        self._lnp = None
        self._nacc = None
        # multi-GPU: each rank evaluates its slice of every half-step.  Default: the library's device-side
        # exchange (walker kernel stores into every rank's region over NVLink + flags; the whole loop runs
        # inside spamm_stretch_run_sharded).  SPAMM_EXCHANGE=nccl: all-gather per half-step from Python.
        import os
        from .distributed import ShardedLnPost, ExchangeLnPost, Exchange, world_info, replicate_object
        self.exchange = "single"
        self._xch = None
        self.rank, self.world = world_info()
        if self.world > 1:
            # every rank proposes the whole half-ensemble from the same counter-based RNG: the seed (and, in
            # run_mcmc, the start state) must be rank 0's, whatever each rank's own numpy RNG drew
            self.seed = int(replicate_object(self.seed))
            if os.environ.get("SPAMM_EXCHANGE", "p2p") == "p2p" and ExchangeLnPost.available():
                self._xch = Exchange(nwalkers)
                self._evaluate = ExchangeLnPost(self.plan, nwalkers, exchange=self._xch)
                self.exchange = "p2p"
            else:
                self._evaluate = ShardedLnPost(self.plan.lnposterior)
                self.exchange = "nccl"
        else:
            self._evaluate = ShardedLnPost(self.plan.lnposterior)
        self._run = 0                    # bumped by reset(): a new run draws a fresh random stream

    # -- emcee-style accessors --------------------------------------------------
    @property
    def chain(self):
        if "_chain_np" in self.__dict__:
            return self._chain_np
        return None if self._chain is None else self._chain[:, :self.iterations, :].cpu().numpy()

    @property
    def lnprobability(self):
        if "_lnprob_np" in self.__dict__:
            return self._lnprob_np
        return None if self._lnprob is None else self._lnprob[:, :self.iterations].cpu().numpy()

    @property
    def flatchain(self):
        c = self.chain
        return c.reshape(-1, self.dim)

    @property
    def naccepted(self):
        if "_nacc_np" in self.__dict__:
            return self._nacc_np
        return self._nacc.cpu().numpy().astype(np.float64)

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(self.iterations, 1)

    def __getstate__(self):
        """Pickle the results (chain, lnprobability, naccepted) as numpy arrays; device
        buffers, the plan and the library handle stay behind."""
        return {"k": self.k, "dim": self.dim, "a": self.a, "seed": self.seed,
                "iterations": self.iterations, "_chain_np": self.chain,
                "_lnprob_np": self.lnprobability,
                "_nacc_np": None if self._nacc is None else self.naccepted}

    def __setstate__(self, d):
        self.__dict__.update(d)

    def _rng_seed(self):
        # the Philox counter is keyed by (seed, iteration, half, walker); reset() restarts the iteration
        # count, so the run index is folded into the key -- a production run after a burn-in must not
        # replay the burn-in's random numbers
        return (self.seed + 0x9E3779B97F4A7C15 * self._run) & 0xFFFFFFFFFFFFFFFF

    def close(self):
        """Release the multi-GPU exchange region (collective: every rank must call it)."""
        if self._xch is not None:
            self._xch.close()
            self._xch = None

    def reset(self):
        self._run += 1
        self.iterations = 0
        self._chain = self._lnprob = None
        if self._nacc is not None:
            self._nacc.zero_()

    # -- the move ------------------------------------------------------------------
    def _stream(self):
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def half_step(self, half, it):
        """One stretch half-step on the device: propose, evaluate (sharded), accept."""
        vp = ctypes.c_void_p
        st = self._stream()
        _lib.check(self.lib.spamm_stretch_propose(vp(self._walkers.data_ptr()), self.k, self.dim, half,
                                                  self.a, self._rng_seed(), it, vp(self._q.data_ptr()),
                                                  vp(self._z.data_ptr()), st))
        lnp_q = self._evaluate(self._q)
        if not lnp_q.is_contiguous():
            lnp_q = lnp_q.contiguous()
        _lib.check(self.lib.spamm_stretch_accept(vp(self._walkers.data_ptr()), vp(self._lnp.data_ptr()),
                                                 self.k, self.dim, half, vp(self._q.data_ptr()),
                                                 vp(self._z.data_ptr()), vp(lnp_q.data_ptr()), self._rng_seed(),
                                                 it, vp(self._nacc.data_ptr()), st))

    def run_mcmc(self, pos0, N, lnprob0=None):
        """Advance the ensemble N iterations from pos0 ([k, dim] list/array)."""
        import torch
        p0 = np.array(pos0, dtype=np.float64)
        if p0.shape != (self.k, self.dim):
            raise ValueError("pos0 must have shape (nwalkers, dim)")
        if np.any(np.isinf(p0)):
            raise ValueError("At least one parameter value was infinite.")
        if np.any(np.isnan(p0)):
            raise ValueError("At least one parameter value was NaN.")
        dev = torch.device("cuda", torch.cuda.current_device())
        self._walkers = torch.from_numpy(p0).to(dev)
        from .distributed import replicate_tensor
        replicate_tensor(self._walkers)                  # the replicated state is rank 0's
        self._q = torch.empty((self.k // 2, self.dim), dtype=torch.float64, device=dev)
        self._z = torch.empty(self.k // 2, dtype=torch.float64, device=dev)
        if self._nacc is None:
            self._nacc = torch.zeros(self.k, dtype=torch.int64, device=dev)
        if lnprob0 is None:
            self._lnp = self._evaluate(self._walkers).clone()
        else:
            self._lnp = torch.from_numpy(np.array(lnprob0, dtype=np.float64)).to(dev)
            replicate_tensor(self._lnp)
        if bool(torch.isnan(self._lnp).any()):
            raise ValueError("The initial lnprob was NaN.")
        start = self.iterations
        total = start + N
        if self.store_chain:
            chain = torch.empty((self.k, total, self.dim), dtype=torch.float64, device=dev)
            lnprob = torch.empty((self.k, total), dtype=torch.float64, device=dev)
            if self._chain is not None and start:
                chain[:, :start] = self._chain[:, :start]
                lnprob[:, :start] = self._lnprob[:, :start]
            self._chain, self._lnprob = chain, lnprob
        vp = ctypes.c_void_p
        if self.exchange == "single" and self.native_loop:
            # one device: the whole loop runs inside the library (spamm_stretch_run), a few hundred
            # iterations per call so a long run does not queue unboundedly far ahead of the device
            lnp_q = torch.empty(self.k // 2, dtype=torch.float64, device=dev)
            it = start
            while it < total:
                n = min(256, total - it)
                _lib.check(self.lib.spamm_stretch_run(
                    self.plan._h, vp(self._walkers.data_ptr()), vp(self._lnp.data_ptr()), self.k, self.a,
                    self._rng_seed(), it, n, vp(self._q.data_ptr()), vp(self._z.data_ptr()), vp(lnp_q.data_ptr()),
                    vp(self._nacc.data_ptr()),
                    vp(self._chain.data_ptr()) if self.store_chain else None,
                    vp(self._lnprob.data_ptr()) if self.store_chain else None, total, self._stream()))
                it += n
                self.iterations += n
        elif self.exchange == "p2p" and self.native_loop:
            # several devices: the same loop inside the library, every half-step's proposals sharded over
            # the ranks and exchanged by the walker kernel itself (spamm_stretch_run_sharded)
            it = start
            while it < total:
                n = min(256, total - it)
                _lib.check(self.lib.spamm_stretch_run_sharded(
                    self.plan._h, self._xch._h, vp(self._walkers.data_ptr()), vp(self._lnp.data_ptr()), self.k,
                    self.a, self._rng_seed(), it, n, vp(self._q.data_ptr()), vp(self._z.data_ptr()),
                    vp(self._nacc.data_ptr()),
                    vp(self._chain.data_ptr()) if self.store_chain else None,
                    vp(self._lnprob.data_ptr()) if self.store_chain else None, total, self._stream()))
                it += n
                self.iterations += n
        else:
            for it in range(start, total):
                self.half_step(0, it)
                self.half_step(1, it)
                if self.store_chain:
                    _lib.check(self.lib.spamm_chain_record(vp(self._walkers.data_ptr()), vp(self._lnp.data_ptr()),
                                                           self.k, self.dim, it, total,
                                                           vp(self._chain.data_ptr()), vp(self._lnprob.data_ptr()),
                                                           self._stream()))
                self.iterations += 1
        torch.cuda.synchronize()
        self.plan.check()                # a walker outside the plan's tables is an error, not a number
        pos = self._walkers.cpu().numpy()
        lnp = self._lnp.cpu().numpy()
        self._last_run_mcmc_result = (pos, lnp, None)
        return pos, lnp, None
