#!/usr/bin/env python
"""bench.py -- ln-posterior evaluations per second of the full NC+Fe+HG+BC model (BASELINE config 5).

    python bench.py --gpus N --steps K --warmup W            (our CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   (the reference's own ln_posterior on the
                                                               host cores: oracle/_ref under oracle/refshim)

Workload: ONE ensemble of 8192 walkers x 8192 pixels, D = 20, FP64.  A "step" is the ln-posterior work of
one stretch-move iteration of that ensemble: two batched calls of one half-ensemble (4096 walkers) each --
the stretch move updates the halves alternately, so a half is the largest batch a sampler can hand over --
with the parameters already resident in HBM.  On N > 1 GPUs every half is sharded N ways (strong scaling:
the ensemble does not grow) and each call ends with the per-walker results on EVERY rank: the walker kernel
leaves them in its rank's exchange region and raises a flag in all ranks' regions over NVLink, and the
gather kernel queued behind it waits for all flags and pulls the other ranks' slices (include/spamm_b200.h,
"multi-GPU").  `value` = 8192 * K / time, timed with CUDA events,
max over ranks.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ln-posterior evals/sec (walkers x steps), full NC+Fe+HG+BC model, 8k px"
UNIT = "evals/s"
N_PIX = 8192
N_WALKERS = 8192
COMPONENTS = ("PL", "FE", "HOST", "BC")
WORKLOAD = "config5: one ensemble of 8192 walkers x 8192 px, full NC+Fe+HG+BC, D=20, walkers from the priors (seed 42)"


# ---------------------------------------------------------------------------
# algorithmic work per evaluation
# ---------------------------------------------------------------------------
def algorithmic_flops(wl, fe_ranges, n_host=7, n_host_pts=4771):
    """F = 3P [NC] + sum_t(2 N_t Mbar_t + 6 P_t) [Fe] + 7 (2 N_h Mbar_h + 6 P) [host]
         + 22 P [BC] + 6*397*(P_red+1) [BpC] + 5 P [chi^2]."""
    P = len(wl)
    n_t = (4181, 584, 4000)
    mbar_t = (91.3, 100.0, 121.3)
    p_t = [int(((wl >= lo) & (wl <= hi)).sum()) for lo, hi in fe_ranges]
    p_red = int((wl > 3646.0).sum())
    f_nc = 3 * P
    f_fe = sum(2 * n * m + 6 * p for n, m, p in zip(n_t, mbar_t, p_t))
    f_host = n_host * (2 * n_host_pts * 16.5 + 6 * P)
    f_bc = 22 * P
    f_bpc = 6 * 397 * (p_red + 1)
    f_chi = 5 * P
    return float(f_nc + f_fe + f_host + f_bc + f_bpc + f_chi)


FE_RANGES = ((1074.987, 3089.96), (3090.0, 3534.0), (3535.0, 7535.0))


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock, power and throttle reasons polled through NVML every 5 ms while a timed region runs
    (`nvidia-smi -lms` cannot sample a region of tens of milliseconds); falls back to one nvidia-smi
    query if NVML is unavailable."""

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.nv = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None
        return self

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) \
                    if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.time(), clk, pw, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self, t0, t1):
        if self.nv is None:
            try:
                out = subprocess.check_output(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                     "--format=csv,noheader,nounits"], text=True).strip().split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]),
                        "reasons": ["NVML unavailable: single nvidia-smi sample after the run"]}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source available"]}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nv
        smax = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        if not inside:
            inside = self.samples[-3:]
        reasons = set()
        for _, _, _, rs in inside:
            for n, bit in names.items():
                if rs & bit:
                    reasons.add(n)
        return {"sm_mhz": float(np.median([s[1] for s in inside])), "sm_max_mhz": float(smax),
                "sm_mhz_min": float(np.min([s[1] for s in inside])),
                "power_w_median": float(np.median([s[2] for s in inside])),
                "power_w_max": float(np.max([s[2] for s in inside])), "samples": len(inside),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# CPU legs.  kind "reference": the UNMODIFIED reference (spamm.Model.ln_posterior, Model.py:42-79) from
# oracle/_ref, one Model per worker process, multiprocessing.Pool(cores).map -- exactly what emcee's
# threads=n does with it (Model.py:283-286).  kind "port": the numpy restatement (oracle/spamm_numpy.py),
# a faster CPU program, reported beside it.
# ---------------------------------------------------------------------------
_CPU_MODEL = None
_CPU_KIND = None


def _cpu_init(kind, wl, flux, err):
    global _CPU_MODEL, _CPU_KIND
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import warnings
    warnings.filterwarnings("ignore")
    _CPU_KIND = kind
    if kind == "reference":
        from oracle import ref_runner
        _CPU_MODEL = ref_runner.build_model(wl, flux, err, COMPONENTS)
    else:
        from oracle.spamm_numpy import OracleModel
        _CPU_MODEL = OracleModel(wl, flux, err, comps=list(COMPONENTS))
        _CPU_MODEL.bounds()


def _cpu_eval(p):
    if _CPU_KIND == "reference":
        from oracle import ref_runner
        return ref_runner.ln_posterior(_CPU_MODEL, p)
    return _CPU_MODEL.ln_posterior(p)


def reference_available():
    try:
        from oracle import ref_runner
        return ref_runner.available()
    except Exception:
        return False


def oracle_workload(n_walkers, seed=42):
    """The config-5 workload built without touching the GPU path (reference arm)."""
    from oracle.spamm_numpy import OracleModel
    wl = 1000.0 + np.arange(N_PIX) * (9000.0 / N_PIX)
    truth = {"PL": [5e-15, 2.3], "FE": [1.07988504e-14, 6.91877436e-15, 5e-15, 5450.],
             "HOST": [5e-17, 5.5e-17, 5e-16, 1e-16, 2e-16, 3e-16, 4e-16, 515.],
             "BC": [3e-14, 50250., 1.0, 0.0, 5050., 5.5]}
    o0 = OracleModel(wl, wl, wl, comps=list(COMPONENTS))
    fl = [o0.component_flux(c, truth[c]) for c in COMPONENTS]
    flux = np.sum(fl, axis=0)
    err = np.sqrt(np.sum([(0.05 * f) ** 2 for f in fl], axis=0))
    o = OracleModel(wl, flux, err, comps=list(COMPONENTS))
    np.random.seed(seed)
    params = np.array([o.initial_values() for _ in range(n_walkers)])
    return wl, flux, err, params


def cpu_pool_throughput(kind, wl, flux, err, params, cores):
    """evals/s of multiprocessing.Pool(cores).map(ln_posterior, walkers)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(kind, wl, flux, err)) as pool:
        pool.map(_cpu_eval, list(params[:cores]))       # warm the workers
        t0 = time.perf_counter()
        vals = pool.map(_cpu_eval, list(params), chunksize=max(1, len(params) // (4 * cores)))
        dt = time.perf_counter() - t0
    return len(params) / dt, np.array(vals)


def cpu_one_core(kind, wl, flux, err, params):
    _cpu_init(kind, wl, flux, err)
    _cpu_eval(params[0])
    t0 = time.perf_counter()
    for p1 in params:
        _cpu_eval(p1)
    return len(params) / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    kind = "reference" if reference_available() else "port"
    per_step = max(8 * cores, 64)
    wl, flux, err, params = oracle_workload(min(per_step * (args.steps + args.warmup), 4096))
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(kind, wl, flux, err)) as pool:
        for s in range(args.warmup + args.steps):
            chunk = params[(s * per_step) % len(params):][:per_step]
            if len(chunk) < per_step:
                chunk = params[:per_step]
            t0 = time.perf_counter()
            pool.map(_cpu_eval, list(chunk), chunksize=max(1, per_step // (4 * cores)))
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
    total = float(np.sum(times))
    value = per_step * args.steps / total
    what = "the reference's own ln_posterior (spamm/Model.py:42-79, unmodified, oracle/_ref under oracle/refshim)" \
        if kind == "reference" else "numpy port of the reference (oracle/spamm_numpy.py; oracle/_ref absent)"
    sample = "%d walkers per step x %d steps of the config-5 ensemble, %s, multiprocessing.Pool(%d).map" \
             % (per_step, args.steps, what, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "walkers_total": N_WALKERS, "walkers_per_step": per_step, "n_pix": N_PIX},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def _events(n):
    import torch
    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from spamm_b200 import _lib
    from spamm_b200.distributed import init_from_env, Exchange
    from spamm_b200.synth import full_model_workload

    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"                # stdout carries one JSON line, not NCCL's version banner
    rank, world, local_rank = init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()

    W = args.walkers
    Wh = W // 2
    model, params_host = full_model_workload(N_PIX, W, seed=42)          # the same ensemble on every rank
    plan = model.plan
    params = torch.from_numpy(params_host).to(dev)
    halves = [params[:Wh].contiguous(), params[Wh:].contiguous()]
    halves_host = [np.ascontiguousarray(params_host[:Wh]), np.ascontiguousarray(params_host[Wh:])]
    lnp = [torch.empty(Wh, dtype=torch.float64, device=dev) for _ in range(2)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    ex = Exchange(W) if world > 1 else None
    per_gpu = -(-Wh // world)

    def step():
        for h in range(2):
            if ex is None:
                plan.lnposterior(halves[h], out=lnp[h])
            else:
                plan.lnposterior_sharded(halves[h], ex, copy=False)     # results + flag wait, in the region

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # FP64 peak for the roofline denominator (MEASURED_PEAKS.json has no FP64 entry)
    peak = ctypes.c_double(0.0)
    _lib.check(lib.spamm_fp64_peak_probe(20000, 5, ctypes.byref(peak)))

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.05)
    launches0 = plan.launch_count
    evs = _events(args.steps)
    barrier()
    t_wall0 = time.time()
    for e0, e1 in evs:
        flush.zero_()                    # L2 flush between timed iterations (untimed)
        e0.record()
        step()
        e1.record()
    barrier()
    t_wall1 = time.time()
    launches = plan.launch_count - launches0
    ms = np.array([e0.elapsed_time(e1) for e0, e1 in evs])
    total_ms = max_over_ranks(float(ms.sum()))
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
    value = W * args.steps / (total_ms * 1e-3)

    # ---- correctness of what was timed: every rank must hold the WHOLE vector of each half, equal bit for
    # bit to the unsharded evaluation of the same parameters on this rank
    local_full = [plan.lnposterior(halves[h]).cpu().numpy() for h in range(2)]
    verified = True
    if ex is not None:
        for h in range(2):
            got = plan.lnposterior_sharded(halves[h], ex).cpu().numpy()
            verified = verified and np.array_equal(got, local_full[h])
        flag = torch.tensor([0.0 if verified else 1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        assert float(flag.item()) == 0.0, "the exchanged ln-posterior vector differs from the unsharded one on some rank"
    plan.check()

    # ---- end to end through the public API with HOST buffers: per half, H2D of (this rank's share of) the
    # parameters from pinned memory, kernel (+ exchange), D2H of the whole half's results, synchronise
    e2e_steps = max(3, min(args.steps, 10))
    pin_in = [torch.empty(h.shape, dtype=torch.float64, pin_memory=True) for h in halves_host]
    pin_out = [torch.empty(Wh, dtype=torch.float64, pin_memory=True) for _ in range(2)]
    for t, h in zip(pin_in, halves_host):
        t.numpy()[...] = h
    h_in, h_out = [t.numpy() for t in pin_in], [t.numpy() for t in pin_out]
    call = model.lnposterior_batch_sharded if world > 1 else model.lnposterior_batch
    for h in range(2):
        call(h_in[h], out=h_out[h])
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        for h in range(2):
            call(h_in[h], out=h_out[h])
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = W * e2e_steps / t_e2e
    for h in range(2):
        assert np.array_equal(h_out[h], local_full[h]), "host-buffer and device-buffer paths disagree"
    lo_r = min(rank * per_gpu, Wh)
    hi_r = min(lo_r + per_gpu, Wh)
    h2d = 2 * (hi_r - lo_r) * plan.n_dim * 8
    d2h = 2 * Wh * 8

    secondary = {}
    # ---- the same ensemble in ONE call of 8192 walkers (round 1's step), device-timed, one GPU
    if world == 1:
        lnp_all = torch.empty(W, dtype=torch.float64, device=dev)
        for _ in range(3):
            plan.lnposterior(params, out=lnp_all)
        ev1 = _events(5)
        for e0, e1 in ev1:
            flush.zero_()
            e0.record()
            plan.lnposterior(params, out=lnp_all)
            e1.record()
        torch.cuda.synchronize()
        one_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev1]))
        secondary["single_call_8192"] = {"ms": one_ms, "evals_per_s": W / (one_ms * 1e-3)}

    # ---- the kernel that executes every algorithmic flop (bc_exact_line_sum): its FP64 fraction is the one
    # comparable with the SURVEY 8(d) count
    exact = None
    if rank == 0 and world == 1 and not args.no_exact:
        from spamm_b200.synth import build_model
        ds = model.data_spectrum
        m2 = build_model(list(COMPONENTS), np.asarray(ds.spectral_axis), np.asarray(ds.flux),
                         np.asarray(ds.flux_error), max_walkers=Wh)
        m2.components[-1].exact_line_sum = True
        lnp2 = torch.empty(Wh, dtype=torch.float64, device=dev)
        for _ in range(2):
            m2.plan.lnposterior(halves[0], out=lnp2)
        torch.cuda.synchronize()
        ev2 = _events(3)
        for e0, e1 in ev2:
            flush.zero_()
            e0.record()
            m2.plan.lnposterior(halves[0], out=lnp2)
            e1.record()
        torch.cuda.synchronize()
        exact_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev2]))
        b = lnp2.cpu().numpy()
        exact = {"ms_per_half": exact_ms, "evals_per_s": Wh / (exact_ms * 1e-3),
                 "max_rel_diff_vs_default": float(np.max(np.abs(local_full[0] - b) / np.abs(b)))}
        del m2

    # ---- the whole device sampler on this ensemble (emcee 2.2.1 stretch move inside the library):
    # walkers x iterations per second, every half-step sharded over the N GPUs
    if not args.no_sampler:
        from spamm_b200.sampler import EnsembleSampler
        from spamm_b200.Model import ln_posterior
        n_it = args.sampler_iterations
        smp = EnsembleSampler(nwalkers=W, dim=plan.n_dim, lnpostfn=ln_posterior, args=[model], seed=7,
                              store_chain=True)
        smp.run_mcmc(params_host, 5)         # burn the start-up transient; one-off allocations
        pos = smp._walkers.cpu().numpy()
        cs = ClockSampler(local_rank).start() if rank == 0 else None
        for _ in range(2):                   # the first pass pays the allocation of the longer chain buffers
            smp.reset()
            barrier()
            tw0 = time.time()
            t0 = time.perf_counter()
            smp.run_mcmc(pos, n_it)
            dt = max_over_ranks(time.perf_counter() - t0)
            tw1 = time.time()
        sclk = cs.stop(tw0, tw1) if cs is not None else None
        chain_sum = float(np.asarray(smp.chain[:, -1, :]).sum())
        if world > 1:
            # the sharded chain must be the single-GPU chain: every rank re-runs a few iterations on its own device
            # alone (same seed, same start) and compares bit for bit
            n_chk = 6
            sm = EnsembleSampler(nwalkers=W, dim=plan.n_dim, lnpostfn=ln_posterior, args=[model], seed=11)
            sm.run_mcmc(params_host, n_chk)
            sharded_chain, sharded_lnp = sm.chain, sm.lnprobability
            sm.close()
            s1 = EnsembleSampler(nwalkers=W, dim=plan.n_dim, plan=plan, seed=11)
            s1.close()                                     # (its exchange region: not used, released collectively)
            s1.exchange, s1.world, s1._evaluate = "single", 1, plan.lnposterior
            s1.run_mcmc(params_host, n_chk)
            same = np.array_equal(s1.chain, sharded_chain) and np.array_equal(s1.lnprobability, sharded_lnp)
            t = torch.tensor([0.0 if same else 1.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            assert float(t.item()) == 0.0, "the sharded sampler's chain differs from the single-GPU chain"
            del s1, sm
        if world > 1:                        # every rank must hold the same chain
            t = torch.tensor([chain_sum, -chain_sum], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            assert float(t[0].item()) == chain_sum and float(t[1].item()) == -chain_sum, "ranks disagree on the chain"
        secondary["sampler"] = {"walkers": int(W), "iterations": n_it, "ms_per_iteration": 1e3 * dt / n_it,
                                "walker_steps_per_s": W * n_it / dt, "exchange": smp.exchange,
                                "acceptance_fraction": float(smp.acceptance_fraction.mean()), "clocks": sclk,
                                "verified": ("6 iterations of the sharded run equal the single-GPU chain bit for bit on every "
                                             "rank; all ranks hold the same chain") if world > 1 else "n/a",
                                "what": "emcee 2.2.1 stretch move inside the library (spamm_stretch_run%s): propose, "
                                        "ln-posterior, accept for both halves, chain record; wall clock, max over ranks"
                                        % ("_sharded" if world > 1 else "")}
        smp.close()

    # ---- sustained: >= args.sustained seconds of back-to-back steps with the clock record
    if world == 1 and args.sustained > 0:
        n_guess = int(args.sustained / (total_ms / args.steps * 1e-3)) + 1
        cs = ClockSampler(local_rank).start()
        time.sleep(0.02)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        tw0 = time.time()
        e0.record()
        done = 0
        while done < n_guess:
            for _ in range(256):
                step()
            done += 256
            if done % 2048 == 0:                      # keep the launch queue bounded
                torch.cuda.current_stream().synchronize()
        e1.record()
        torch.cuda.synchronize()
        tw1 = time.time()
        s_ms = e0.elapsed_time(e1)
        secondary["sustained"] = {"seconds": s_ms * 1e-3, "steps": done, "ms_per_step": s_ms / done,
                                  "evals_per_s": W * done / (s_ms * 1e-3),
                                  "vs_burst": (W * done / (s_ms * 1e-3)) / value,
                                  "l2": "no flush (back-to-back steps; the 13 MB template bank stays in L2)",
                                  "clocks": cs.stop(tw0, tw1)}
        plan.check()

    # ---- the broadening kernels on the record (north_star: "a Fe-template Gaussian broadening kernel ... applies a
    # per-walker kernel width"): BASELINE config 2 (NC + Fe, 1024 walkers x 4096 px) with the templates broadened
    # PER WALKER inside the call (template_mode DIRECT) beside the default pre-broadened bank, and k_conv_rows /
    # k_interp_rows timed alone against their rooflines
    if world == 1 and not args.no_broadening:
        from spamm_b200.synth import build_model
        m_b, p_b = full_model_workload(4096, 1024, seed=42, components=("PL", "FE"))
        ds = m_b.data_spectrum
        m_d = build_model(["PL", "FE"], np.asarray(ds.spectral_axis), np.asarray(ds.flux), np.asarray(ds.flux_error),
                          max_walkers=1024, template_mode=_lib.TEMPLATES_DIRECT)
        pb = torch.from_numpy(p_b).to(dev)
        ob = torch.empty(1024, dtype=torch.float64, device=dev)
        res = {}
        for name, mm in (("bank", m_b), ("direct", m_d)):
            for _ in range(3):
                mm.plan.lnposterior(pb, out=ob)
            evb = _events(5)
            for e0, e1 in evb:
                flush.zero_()
                e0.record()
                mm.plan.lnposterior(pb, out=ob)
                e1.record()
            torch.cuda.synchronize()
            res[name] = float(np.mean([e0.elapsed_time(e1) for e0, e1 in evb]))
            res[name + "_values"] = ob.cpu().numpy().copy()
        # (the bank runs the NC+Fe specialised instantiation of the walker kernel, per-walker rows the generic one:
        #  same rows bit for bit -- tests/test_gpu_parity.py::test_bank_equals_direct_broadening -- a few ulp apart
        #  in the final sum)
        bd_rel = float(np.max(np.abs(res["bank_values"] - res["direct_values"]) / np.abs(res["bank_values"])))
        assert bd_rel < 1e-12, "bank and direct broadening disagree"
        hbm = 6465.2
        mp_ = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp_):
            hbm = json.load(open(mp_)).get("hbm_gbs", hbm)
        probes = []
        for sig in (8, 30, 60):
            cms, ims, fma = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
            _lib.check(lib.spamm_broadening_probe(m_d.plan._h, 0, sig, 3072, 5, ctypes.byref(cms), ctypes.byref(ims),
                                                  ctypes.byref(fma)))
            tfl = 2.0 * fma.value / (cms.value * 1e-3) / 1e12
            n_pts = (4181 + 584 + 4000) / 3.0
            byt = 3072 * (n_pts * 8 + 4096 * 8)
            probes.append({"sigma_norm": sig, "taps": 3 * sig, "rows": 3072,
                           "k_conv_rows": {"ms": cms.value, "tflops": tfl, "frac_of_dfma_peak": tfl / float(peak.value)},
                           "k_interp_rows+k_norm_rows": {"ms": ims.value, "gbs": byt / (ims.value * 1e-3) / 1e9,
                                                         "frac_of_hbm_peak": byt / (ims.value * 1e-3) / 1e9 / hbm}})
        secondary["broadening"] = {
            "config2": {"workload": "NC + Fe, 1024 walkers x 4096 px (BASELINE config 2)",
                        "bank_ms": res["bank"], "direct_ms": res["direct"],
                        "direct_evals_per_s": 1024 / (res["direct"] * 1e-3),
                        "max_rel_diff_bank_vs_direct": bd_rel,
                        "what": "direct = k_rows_from_params + k_conv_rows + k_interp_rows + k_norm_rows per call "
                                "(3072 rows), then k_walkers; asynchronous (kernel widths bounded by the width prior)"},
            "kernels": probes,
            "rooflines": "k_conv_rows: FP64 (2 flop per tap and point, direct convolution on the ln-lambda grid) "
                         "against the DFMA peak measured in this run; k_interp_rows: HBM (reads the convolved row, "
                         "writes P doubles) against MEASURED_PEAKS.json hbm_gbs"}
        del m_b, m_d

    # ---- weak scaling (secondary): 8192 walkers per GPU in one sharded call
    if world > 1 and not args.no_weak:
        exw = Exchange(W * world)
        pw = params.repeat(world, 1).contiguous()
        for _ in range(3):
            plan.lnposterior_sharded(pw, exw, copy=False)
        barrier()
        evw = _events(5)
        for e0, e1 in evw:
            flush.zero_()
            e0.record()
            plan.lnposterior_sharded(pw, exw, copy=False)
            e1.record()
        barrier()
        w_ms = max_over_ranks(float(np.sum([e0.elapsed_time(e1) for e0, e1 in evw])))
        secondary["weak"] = {"walkers_per_gpu": W, "walkers_total": W * world, "ms_per_call": w_ms / 5,
                             "evals_per_s": W * world * 5 / (w_ms * 1e-3)}
        exw.close()

    if rank == 0:
        wl = np.asarray(model.data_spectrum.spectral_axis)
        flops = algorithmic_flops(wl, FE_RANGES)
        prof = {}
        tpath = os.path.join(ROOT, "profiles", "k_walkers_traffic.json")
        if os.path.exists(tpath):
            prof = json.load(open(tpath))
        launch_ms = total_ms / args.steps / 2.0                 # one launch per half
        evals_per_s_gpu = per_gpu / (launch_ms * 1e-3)          # this GPU's kernel rate
        exec_flops = prof.get("executed_flops_per_eval")
        achieved = exec_flops * evals_per_s_gpu / 1e12 if exec_flops else None
        roofline = {
            "bound": "fp64", "achieved": achieved, "peak": float(peak.value), "unit": "TFLOP/s",
            "frac": achieved / float(peak.value) if achieved and peak.value else None,
            "traffic": prof.get("dram_bytes_per_launch"),
            "kernel": "k_walkers<0, true, false, 15%s> (full-model instantiation), %d walkers per launch"
                      % (" | 128: walkers of the partial wave as two CTAs" if plan_uses_parts(per_gpu) else "", per_gpu),
            "launch_ms": launch_ms,
            "what": "frac = FP64 flops the kernel EXECUTES per evaluation (ncu: 2 x DFMA + DADD + DMUL thread "
                    "instructions, profiles/k_walkers_traffic.json) x evaluations/s of the launch / measured DFMA peak",
            "executed_flops_per_eval": exec_flops,
            "algorithmic_flops_per_eval": flops,
            "algorithmic_ratio": flops * evals_per_s_gpu / 1e12 / float(peak.value) if peak.value else None,
            "algorithmic_ratio_note": "SURVEY 8(d) counts every one of the 397 x P_red Lorentzians; the kernel sums "
                                      "distant line clusters by a far-field series and reads pre-broadened template "
                                      "rows, so it executes ~12x fewer flops than that count -- this ratio is a "
                                      "speed-up over the naive algorithm, not a utilisation",
            "fp64_pipe_busy_ncu": prof.get("fp64_pipe_busy_default"),
            "peak_source": "DFMA microbenchmark measured in this run (spamm_fp64_peak_probe); MEASURED_PEAKS.json "
                           "has no FP64 entry; nominal 148*64*2*1.965 GHz = 37.2",
            "traffic_note": "DRAM bytes per launch from one ncu --set full capture: the 13.4 MB pre-broadened "
                            "template bank refilled into L2 after the deliberate flush between timed steps (2.5 us "
                            "at the measured HBM rate) + parameters in / results out; nothing is re-read"}
        hbm_peak = None
        mp_ = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp_):
            hbm_peak = json.load(open(mp_)).get("hbm_gbs")
        alg_bytes = 8.0 * (plan.n_dim + 1) * per_gpu
        gbs = alg_bytes / (launch_ms * 1e-3) / 1e9
        roofline["hbm_view"] = {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": gbs,
                                "peak_gbs": hbm_peak if hbm_peak else 6650.0,
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "fallback 6.65 TB/s",
                                "frac": gbs / (hbm_peak if hbm_peak else 6650.0),
                                "note": "arithmetic intensity ~1e4 executed flop/B: HBM is not the bound"}
        if exact is not None:
            ach2 = flops * exact["evals_per_s"] / 1e12
            exact.update({"achieved": ach2, "frac": ach2 / float(peak.value),
                          "what": "bc_exact_line_sum kernel: executes the SURVEY 8(d) flops (4096 walkers)"})
            roofline["exact_line_sum"] = exact
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            flux, err = np.asarray(model.data_spectrum.flux), np.asarray(model.data_spectrum.flux_error)
            gpu_vals = np.concatenate(local_full)
            n_port = max(16 * cores, 128)
            v_port, vals_port = cpu_pool_throughput("port", wl, flux, err, params_host[:n_port], cores)
            rel_port = float(np.max(np.abs(vals_port - gpu_vals[:n_port]) / np.abs(vals_port)))
            port = {"value": v_port, "cores": cores, "walkers": n_port, "max_rel_diff_vs_gpu": rel_port,
                    "value_one_core": cpu_one_core("port", wl, flux, err, params_host[:16])}
            if reference_available():
                n_ref = max(12 * cores, 96)
                v_ref, vals_ref = cpu_pool_throughput("reference", wl, flux, err, params_host[:n_ref], cores)
                rel_ref = float(np.max(np.abs(vals_ref - gpu_vals[:n_ref]) / np.abs(vals_ref)))
                cpu = {"value": v_ref, "unit": UNIT, "cores": cores, "kind": "reference",
                       "value_one_core": cpu_one_core("reference", wl, flux, err, params_host[:8]),
                       "sample": "first %d walkers of the same ensemble through the reference's own ln_posterior "
                                 "(spamm/Model.py:42-79, unmodified, oracle/_ref under oracle/refshim), "
                                 "multiprocessing.Pool(%d).map" % (n_ref, cores),
                       "max_rel_diff_vs_gpu": rel_ref, "numpy_port": port}
            else:
                cpu = {"value": v_port, "unit": UNIT, "cores": cores, "kind": "port",
                       "value_one_core": port["value_one_core"],
                       "sample": "first %d walkers of the same ensemble, numpy port of the reference (oracle/_ref "
                                 "absent), multiprocessing.Pool(%d).map" % (n_port, cores),
                       "max_rel_diff_vs_gpu": rel_port}
        exchange = "none (one GPU)" if world == 1 else \
            "walker kernel leaves its results in its rank's region (CUDA IPC mapped) and raises a flag in every " \
            "rank's region over NVLink; a gather kernel waits for all flags and pulls the peers' slices; " \
            "no collective, no host barrier"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "walkers_total": W, "n_pix": N_PIX,
                       "step": "two batched half-ensemble calls (%d walkers each), each sharded over %d GPU(s)"
                               % (Wh, world),
                       "walkers_per_gpu_per_call": per_gpu,
                       "l2": "256 MiB flush between timed steps", "parallelism": "walkers sharded x%d" % world,
                       "exchange": exchange, "verified": "every rank holds the whole vector, bit-identical to the "
                                                         "unsharded evaluation" if world > 1 else "n/a",
                       "template_mode": "bank" if plan.template_mode == _lib.TEMPLATES_BANK else "direct"},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "what": "Model.lnposterior_batch%s with pinned host buffers, two half-ensemble calls per step"
                            % ("_sharded" if world > 1 else "")},
            "gpu_launches": int(launches),
            "ms_per_step_minmax": [float(ms.min()), float(ms.max())],
            "secondary": secondary,
        }
        print(json.dumps(line))
    if ex is not None:
        ex.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def plan_uses_parts(n):
    """Mirror of the library's automatic choice (spamm_b200.cu parts_count): from a third of a wave of CTAs up to
    four waves some or all walkers of a launch run as two independent CTAs (the kSpecParts instantiation)."""
    slots = 148 * 4
    if 3 * n <= slots or n > 4 * slots:
        return False
    return 2 * n <= 3 * slots or n % slots != 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--walkers", type=int, default=N_WALKERS, help="walkers of the (one) ensemble")
    ap.add_argument("--sustained", type=float, default=5.0, help="seconds of the sustained leg (0: skip; N = 1 only)")
    ap.add_argument("--sampler-iterations", type=int, default=100,
                    help="iterations of the timed device-sampler run (its fixed costs -- state upload, broadcast, the "
                         "first evaluation, result download -- are inside the wall clock)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exact", action="store_true", help="skip the exact-line-sum comparison run")
    ap.add_argument("--no-sampler", action="store_true", help="skip the device-sampler figure")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling figure")
    ap.add_argument("--no-broadening", action="store_true", help="skip the broadening-kernel figures (N = 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
