#!/usr/bin/env python
"""bench.py -- ln-posterior evaluations per second of the full NC+Fe+HG+BC model.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   (the CPU restatement of
                                                               the reference, all host cores)

A "step" is one batched ln-posterior call over the rank's walker ensemble
(BASELINE.json config 5: 8192 walkers x 8192 pixels, D = 20, FP64), parameters
already resident in HBM, followed -- when N > 1 -- by the all-gather of the
per-walker ln-probabilities (the only exchange of the stretch move).  One JSON
line is printed by rank 0.
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
    """SM clock, power and throttle reasons polled through NVML every 5 ms while the timed
    region runs (the region lasts tens of milliseconds, too short for `nvidia-smi -lms`);
    falls back to one nvidia-smi query if NVML is unavailable."""

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
                "power_w_max": float(np.max([s[2] for s in inside])), "samples": len(inside),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# CPU baseline: the numpy restatement of the reference, one process per core
# ---------------------------------------------------------------------------
_ORACLE = None


def _oracle_init(wl, flux, err):
    global _ORACLE
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle.spamm_numpy import OracleModel
    _ORACLE = OracleModel(wl, flux, err, comps=list(COMPONENTS))
    _ORACLE.bounds()


def _oracle_eval(p):
    return _ORACLE.ln_posterior(p)


def oracle_workload(n_walkers, seed=42):
    """Same workload built without touching the GPU path (reference arm)."""
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


def cpu_pool_throughput(wl, flux, err, params, cores, repeats=1):
    """evals/s of multiprocessing.Pool(cores).map(ln_posterior, walkers) -- what
    emcee's threads=n does with the reference (spamm/Model.py:283-286)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_oracle_init, initargs=(wl, flux, err)) as pool:
        pool.map(_oracle_eval, list(params[:cores]))       # warm the workers
        best = 0.0
        vals = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            vals = pool.map(_oracle_eval, list(params), chunksize=max(1, len(params) // (4 * cores)))
            dt = time.perf_counter() - t0
            best = max(best, len(params) / dt)
    return best, np.array(vals)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_step = max(8 * cores, 64)
    n = per_step * (args.steps + args.warmup)
    wl, flux, err, params = oracle_workload(min(n, 4096))
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores, initializer=_oracle_init, initargs=(wl, flux, err)) as pool:
        for s in range(args.warmup + args.steps):
            chunk = params[(s * per_step) % len(params):][:per_step]
            if len(chunk) < per_step:
                chunk = params[:per_step]
            t0 = time.perf_counter()
            pool.map(_oracle_eval, list(chunk), chunksize=max(1, per_step // (4 * cores)))
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
    total = float(np.sum(times))
    value = per_step * args.steps / total
    sample = "%d walkers per step x %d steps of the config-5 ensemble, numpy port of the reference, " \
             "multiprocessing.Pool(%d)" % (per_step, args.steps, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "config5: full NC+Fe+HG+BC, 8192 px, D=20, walkers from priors (seed 42)",
                   "walkers_per_step": per_step, "n_pix": N_PIX},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from spamm_b200 import _lib
    from spamm_b200.distributed import init_from_env
    from spamm_b200.synth import full_model_workload

    rank, world, local_rank = init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()

    n_local = args.walkers
    if args.scaling == "strong":
        n_local = (args.walkers + world - 1) // world
    model, params_all = full_model_workload(N_PIX, args.walkers if args.scaling == "strong" else args.walkers,
                                            seed=42 + (rank if args.scaling == "weak" else 0))
    if args.scaling == "strong":
        params_host = np.ascontiguousarray(params_all[rank * n_local:(rank + 1) * n_local])
        n_local = len(params_host)
    else:
        params_host = params_all
    plan = model.plan
    params = torch.from_numpy(params_host).to(dev)
    lnp = torch.empty(n_local, dtype=torch.float64, device=dev)
    gathered = torch.empty(n_local * world, dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    exchange = "none"
    p2p = None
    if world > 1:
        from spamm_b200.distributed import P2PLnPost
        exchange = "nccl all_gather"
        if args.exchange == "p2p" and P2PLnPost.available():
            try:
                p2p = P2PLnPost(plan, max_n=n_local * world)
                exchange = "P2P stores from the compute kernel into every peer's symmetric buffer + barrier"
            except Exception as exc:                                  # symmetric memory unavailable
                p2p = None
                exchange = "nccl all_gather (p2p unavailable: %s)" % type(exc).__name__
    p2p_step = [0]

    def step():
        if p2p is not None:
            k = p2p_step[0] & 1
            p2p_step[0] += 1
            plan.lnposterior_p2p(params, p2p.hdls[k].buffer_ptrs_dev, world, rank * n_local)
            p2p.hdls[k].barrier(channel=0)
            return
        plan.lnposterior(params, out=lnp)
        if world > 1:
            dist.all_gather_into_tensor(gathered, lnp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 peak for the roofline denominator (MEASURED_PEAKS.json has no FP64 entry)
    import ctypes
    peak = ctypes.c_double(0.0)
    _lib.check(lib.spamm_fp64_peak_probe(20000, 5, ctypes.byref(peak)))

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.05)
    launches0 = plan.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
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
    total_ms = float(ms.sum())
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None

    # end to end through the public API with host buffers (H2D + kernel + D2H every step)
    e2e_steps = max(3, min(args.steps, 10))
    pin_in = torch.empty(params_host.shape, dtype=torch.float64, pin_memory=True)
    pin_out = torch.empty(n_local, dtype=torch.float64, pin_memory=True)
    pin_in.numpy()[...] = params_host                 # the step's inputs live in pinned host memory
    h_in, h_out = pin_in.numpy(), pin_out.numpy()
    model.lnposterior_batch(h_in, out=h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = model.lnposterior_batch(h_in, out=h_out)      # H2D + kernel + D2H + sync, every step
    t_e2e = time.perf_counter() - t0
    res = np.array(res)
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = n_local * world * e2e_steps / t_e2e
    if p2p is not None:
        # every rank's buffer holds the whole gathered vector; this rank's slice must equal
        # what the plain local call returns
        got = p2p.bufs[(p2p_step[0] - 1) & 1][rank * n_local:(rank + 1) * n_local].cpu().numpy()
        plan.lnposterior(params, out=lnp)
        assert np.array_equal(got, lnp.cpu().numpy()), "P2P-exchanged results differ from the local call"
    assert np.array_equal(res, lnp.cpu().numpy()), "host-buffer and device-buffer paths disagree"

    value = n_local * world * args.steps / (total_ms * 1e-3)

    # the same ensemble with every Balmer line summed directly (bc_exact_line_sum): this kernel
    # executes the algorithmic flops of SURVEY.md 8(d), so its FP64 fraction is the comparable one
    exact_ms = None
    if rank == 0 and not args.no_exact:
        from spamm_b200.synth import build_model
        ds = model.data_spectrum
        m2 = build_model(list(COMPONENTS), np.asarray(ds.spectral_axis), np.asarray(ds.flux),
                         np.asarray(ds.flux_error), max_walkers=n_local)
        m2.components[-1].exact_line_sum = True
        lnp2 = torch.empty_like(lnp)
        for _ in range(3):
            m2.plan.lnposterior(params, out=lnp2)
        torch.cuda.synchronize()
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for e0, e1 in ev2:
            flush.zero_()
            e0.record()
            m2.plan.lnposterior(params, out=lnp2)
            e1.record()
        torch.cuda.synchronize()
        exact_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev2]))
        a, b = lnp.cpu().numpy(), lnp2.cpu().numpy()
        exact_rel = float(np.max(np.abs(a - b) / np.abs(b)))

    # secondary figures of SURVEY.md 8(d): one stretch half-step (W/2 walkers per call) and the whole
    # device sampler (walkers x iterations per second, proposals + ln-posterior + accept + chain record)
    secondary = None
    if world == 1 and not args.no_sampler:
        from spamm_b200.sampler import EnsembleSampler
        from spamm_b200.Model import ln_posterior
        half = params[: n_local // 2].contiguous()
        lnp_h = torch.empty(len(half), dtype=torch.float64, device=dev)
        for _ in range(3):
            plan.lnposterior(half, out=lnp_h)
        torch.cuda.synchronize()
        evh = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for e0, e1 in evh:
            flush.zero_()
            e0.record()
            plan.lnposterior(half, out=lnp_h)
            e1.record()
        torch.cuda.synchronize()
        half_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in evh]))
        n_it = 40
        for _ in range(2):         # the first pass pays the one-off device allocations of the chain buffers
            smp = EnsembleSampler(nwalkers=n_local, dim=plan.n_dim, lnpostfn=ln_posterior, args=[model], seed=7,
                                  store_chain=True)
            smp.run_mcmc(params_host, 5)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            smp.run_mcmc(smp._walkers.cpu().numpy(), n_it)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        secondary = {"half_step": {"walkers": int(len(half)), "ms": half_ms,
                                   "evals_per_s": len(half) / (half_ms * 1e-3)},
                     "sampler": {"walkers": int(n_local), "iterations": n_it, "ms_per_iteration": 1e3 * dt / n_it,
                                 "walker_steps_per_s": n_local * n_it / dt,
                                 "acceptance_fraction": float(smp.acceptance_fraction.mean()),
                                 "what": "emcee 2.2.1 stretch move on the device (spamm_stretch_run): propose, "
                                         "ln-posterior, accept for both halves, chain record; wall clock"}}

    line = None
    if rank == 0:
        wl = np.asarray(model.data_spectrum.spectral_axis)
        flops = algorithmic_flops(wl, FE_RANGES)
        achieved = flops * (n_local * args.steps / (total_ms * 1e-3)) / 1e12    # this GPU's share
        traffic, pipe_busy = None, None
        tpath = os.path.join(ROOT, "profiles", "k_walkers_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic, pipe_busy = tj.get("dram_bytes_per_launch"), tj.get("fp64_pipe_busy_default")
        roofline = {"bound": "fp64", "achieved": achieved, "peak": float(peak.value), "unit": "TFLOP/s",
                    "frac": achieved / float(peak.value) if peak.value else None, "traffic": traffic,
                    "peak_source": "DFMA microbenchmark measured in this run (spamm_fp64_peak_probe); "
                                   "MEASURED_PEAKS.json has no FP64 entry; nominal 148*64*2*1.965 GHz = 37.2",
                    "algorithmic_flops_per_eval": flops,
                    "kernel": "k_walkers<0, true, false, 15> (full-model instantiation)",
                    "note": "achieved = SURVEY 8(d) algorithmic flops x evals/s; the default kernel sums "
                            "distant Balmer-line clusters by a far-field series and executes ~8x fewer "
                            "flops than that count, so frac can exceed 1; exact_line_sum below is the "
                            "kernel that executes them all",
                    "fp64_pipe_busy_ncu": pipe_busy}
        # HBM view, for completeness: 168 algorithmic bytes per evaluation (params in, ln-prob out)
        hbm_peak = None
        mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp):
            hbm_peak = json.load(open(mp)).get("hbm_gbs")
        alg_bytes = 8.0 * (plan.n_dim + 1) * n_local
        gbs = alg_bytes / (total_ms / args.steps * 1e-3) / 1e9
        roofline["hbm_view"] = {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": gbs,
                                "peak_gbs": hbm_peak if hbm_peak else 6650.0,
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "fallback 6.65 TB/s",
                                "frac": gbs / (hbm_peak if hbm_peak else 6650.0),
                                "note": "arithmetic intensity ~1e5 flop/B: HBM is not the bound"}
        if exact_ms is not None:
            ach2 = flops * (n_local / (exact_ms * 1e-3)) / 1e12
            roofline["exact_line_sum"] = {"ms_per_step": exact_ms, "evals_per_s": n_local / (exact_ms * 1e-3),
                                          "achieved": ach2, "frac": ach2 / float(peak.value),
                                          "max_rel_diff_vs_default": exact_rel}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_cpu = max(16 * cores, 128)
            v, vals = cpu_pool_throughput(wl, np.asarray(model.data_spectrum.flux),
                                          np.asarray(model.data_spectrum.flux_error),
                                          params_host[:n_cpu], cores)
            gpu_vals = res[:n_cpu]
            rel = float(np.max(np.abs(vals - gpu_vals) / np.abs(vals)))
            # one core, in process (SURVEY.md 8(d) asks for n = 1 and n = all cores)
            _oracle_init(wl, np.asarray(model.data_spectrum.flux), np.asarray(model.data_spectrum.flux_error))
            t0 = time.perf_counter()
            for p1 in params_host[:24]:
                _oracle_eval(p1)
            v1 = 24.0 / (time.perf_counter() - t0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "value_one_core": v1,
                   "sample": "first %d walkers of the same ensemble, numpy port of the reference, "
                             "multiprocessing.Pool(%d).map" % (n_cpu, cores),
                   "max_rel_diff_vs_gpu": rel}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config5: full NC+Fe+HG+BC, 8192 px, D=20, walkers from priors (seed 42)",
                       "walkers_per_gpu": n_local, "walkers_total": n_local * world, "n_pix": N_PIX,
                       "l2": "256 MiB flush between timed steps", "parallelism": "walkers sharded x%d" % world, "exchange": exchange,
                       "template_mode": "bank" if plan.template_mode == _lib.TEMPLATES_BANK else "direct"},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(params_host.nbytes),
                    "d2h_bytes_per_step": int(n_local * 8), "steps": e2e_steps},
            "gpu_launches": int(launches),
            "ms_per_step_minmax": [float(ms.min()), float(ms.max())],
            "secondary": secondary,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--walkers", type=int, default=N_WALKERS, help="walkers per GPU (weak) / total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exact", action="store_true", help="skip the exact-line-sum comparison run")
    ap.add_argument("--no-sampler", action="store_true", help="skip the half-step / device-sampler figures")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how per-walker ln-probs reach every rank")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
