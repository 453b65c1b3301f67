// spamm_b200 -- the fused walker kernel k_walkers and the stand-alone likelihood kernel
// (part of the single translation unit spamm_b200.cu; reference call sites are cited as
// file:line relative to the reference root)
#pragma once
#include <cooperative_groups.h>

#include "plan_dev.cuh"
#include "device_math.cuh"
#include "template_kernels.cuh"
// ---------------------------------------------------------------------------
// the fused walker kernel
// ---------------------------------------------------------------------------
struct DirectRows {          // per-walker broadened templates (direct mode); null in bank mode
    const double* fe_f;  const double* fe_nf;
    const double* host_f; const double* host_nf;
};

struct TemplCoef {
    const double* f;   // row on the data grid
    double a;          // norm_t / nf
};

// Multi-GPU result exchange.  Every rank owns an exchange REGION (plan_dev.cuh: kFlagWords flag words, then
// two result buffers) that all peers have mapped (CUDA IPC / NVLink).  A sharded call works like this:
//   producer  k_walkers stores each of its walkers' results into its OWN region (bufs != nullptr selects
//             that instead of out[w]), fences at device scope and counts the walker; the thread that retires
//             the call's last walker fences at system scope and raises this rank's flag -- word `rank` of
//             EVERY rank's flag block, value `epoch` -- with release stores over NVLink.
//   consumer  (k_gather, k_stretch_accept) waits until every rank's flag in its own region has reached the
//             epoch, then PULLS the entries other ranks own straight out of their regions (peer loads).
// No per-walker traffic or system fence on the producer side, no collective, no host barrier.
// (Measured alternatives, profiles/r02_experiments.txt: per-walker P2P stores + system fence from every
//  finishing thread cost 19 us per call on 2 GPUs -- the fence's NVLink round trip delays every CTA's exit;
//  a push of the whole slice by the last CTA 29 us.)
struct PeerOut {
    const unsigned long long* bufs;   // device array of n_peers region base pointers (incl. this rank's); null: out[w]
    int n_peers;
    long long offset;                 // element offset of this launch's walker 0 from a base pointer
    unsigned* done_cnt;               // device counter of retired walkers
    unsigned n_done;                  // walkers of the whole sharded call (may span several launches)
    int rank;                         // this rank
    unsigned long long epoch;         // value to raise
};

// what a consumer needs to find walker i of the gathered vector: rank i / per owns it
struct PeerIn {
    const unsigned long long* bufs;   // region base pointers; null: single-device call, nothing to wait for
    int n_peers, rank;
    long long boff;                   // element offset of the epoch's result buffer in a region
    int per;                          // walkers per rank (spamm_shard_bounds)
    unsigned long long epoch;
    int* status;
    int* status_host;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// raise this rank's flag in every rank's flag block: ONE system-scope fence orders every result of the call
// before the flag stores, which then go out back to back (a release store per peer would wait for the NVLink
// round trip of the one before it: measured +8 us per call on 8 GPUs)
__device__ __forceinline__ void raise_flags(const PeerOut& po) {
    __threadfence_system();
    for (int r = 0; r < po.n_peers; ++r)
        st_relaxed_sys(reinterpret_cast<unsigned long long*>(po.bufs[r]) + po.rank, po.epoch);
}

__device__ __forceinline__ void store_result(double* out, int w, double v, const PeerOut& po) {
    if (po.bufs == nullptr) { out[w] = v; return; }
    reinterpret_cast<double*>(po.bufs[po.rank])[po.offset + w] = v;
    __threadfence();                                              // the result before the count
    if (atomicAdd(po.done_cnt, 1u) + 1u == po.n_done) {           // this walker retires the call
        *po.done_cnt = 0u;
        raise_flags(po);
    }
}

// a rank without walkers in a sharded call still has to raise its flag
__global__ void k_raise_flags(PeerOut po) { raise_flags(po); }

// Consumer side, called by every thread of a block: threads 0..n_peers-1 spin until rank r's flag in this
// rank's flag block has reached the epoch; gives up after ~5 s and sets kStatPeerTimeout so a lost peer
// cannot hang the device.
__device__ __forceinline__ void wait_flags(const PeerIn& pi) {
    if ((int)threadIdx.x < pi.n_peers) {
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(pi.bufs[pi.rank]) + threadIdx.x;
        const unsigned long long t0 = global_timer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(flag) < pi.epoch) {
            if ((++spins & 1023u) == 0u && global_timer_ns() - t0 > 5000000000ull) {
                flag_status(pi.status, pi.status_host, kStatPeerTimeout);
                break;
            }
        }
    }
    __syncthreads();
}

// entry i of the gathered vector, after wait_flags: from this rank's region or pulled from the owner's
__device__ __forceinline__ double gathered_value(const PeerIn& pi, int i) {
    const int owner = i / pi.per;
    const double* src = reinterpret_cast<const double*>(pi.bufs[owner]) + pi.boff + i;
    return owner == pi.rank ? __ldcg(src) : ld_volatile_f64(src);
}

// complete the gathered vector of a sharded call in this rank's region (and in `copy`, if given)
__global__ void __launch_bounds__(256) k_gather(PeerIn pi, int n_total, double* __restrict__ copy) {
    wait_flags(pi);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const double v = gathered_value(pi, i);
    if (i / pi.per != pi.rank) reinterpret_cast<double*>(pi.bufs[pi.rank])[pi.boff + i] = v;
    if (copy != nullptr) copy[i] = v;
}

// The library's sampler loop lets the walker kernel make the stretch proposal itself (no propose kernel, no q / z
// round trip through memory): walker w of the launch is proposal i0 + w of the active half, and its parameter
// vector is c_j - z (c_j - s_i) drawn from the Philox counter exactly as k_stretch_propose / stretch_draw do.
struct Proposal {
    const double* walkers;     // [K][D] ensemble; null: parameters come from `params`
    int K, half, i0;
    double a;
    unsigned long long seed, iter;
};
__device__ __forceinline__ void stretch_draw(uint64_t seed, int i, int half, uint64_t iter, double a, int K,
                                             double& zz, int& partner);

// Walker PARTS (kSpecParts instantiations): a walker is evaluated by two independent CTAs -- the "red" part
// takes the 64-pixel chunks from PlanDev::part_chunk (just blueward of the Balmer edge) on and builds the line
// table and cluster moments, the "blue" part takes the chunks before it and stages the Balmer continuum -- so
// neither repeats the other's share of the prologue.  Each part leaves its per-chunk chi^2 partials in `chi`;
// the part that finishes second sums all of them in the unsplit kernel's fixed order (bit-identical result).
// A launch may mix the two forms: its first n_whole walkers run one CTA each (the full waves of CTAs), the
// n_parts walkers after them two CTAs each (the last, partial wave, where finer work items shorten the tail).
struct PartsWs {
    double* chi;           // [walkers][n_chunks]
    unsigned* cnt;         // [walkers], zero between launches
    int n_whole;           // blocks [0, n_whole): one whole walker each
    int n_parts;           // then n_parts red parts, then n_parts blue parts (walkers n_whole ... n_whole + n_parts - 1)
};

// Storey & Hummer table look-up: kx=ky=1 FITPACK spline == clamped bilinear
__device__ double sh95_bilinear(const PlanDev& pl, int k, double n_e, double T) {
    double x = fmin(fmax(n_e, pl.sh_ne[0]), pl.sh_ne[SPAMM_SH95_NE - 1]);
    double y = fmin(fmax(T, pl.sh_T[0]), pl.sh_T[SPAMM_SH95_T - 1]);
    int i = 0, j = 0;
    for (int q = 1; q < SPAMM_SH95_NE - 1; ++q) if (pl.sh_ne[q] <= x) i = q;
    for (int q = 1; q < SPAMM_SH95_T - 1; ++q) if (pl.sh_T[q] <= y) j = q;
    double tx = (x - pl.sh_ne[i]) / (pl.sh_ne[i + 1] - pl.sh_ne[i]);
    double ty = (y - pl.sh_T[j]) / (pl.sh_T[j + 1] - pl.sh_T[j]);
    const double* z = pl.sh_z + (size_t)k * SPAMM_SH95_NE * SPAMM_SH95_T;
    double z00 = z[i * SPAMM_SH95_T + j], z10 = z[(i + 1) * SPAMM_SH95_T + j];
    double z01 = z[i * SPAMM_SH95_T + j + 1], z11 = z[(i + 1) * SPAMM_SH95_T + j + 1];
    return ((1 - tx) * (1 - ty) * z00 + tx * (1 - ty) * z10 + (1 - tx) * ty * z01) + tx * ty * z11;
}

// unnormalised Balmer continuum sample (1 - exp(-tau)) * blackbody_lambda (BC:435-440).
// bc_K[k] = 2 h nu^3 c_AA / (c_cgs^2 lambda^2) is the walker-independent factor of B_lambda.
// SAFE (FULL kernel: the prior box bounds every exp argument, PlanDev::exp_safe): no range checks,
// and the two divisions become a multiplication by 1/kT (<= 1.5 ulp on the exponent, i.e. <= 1e-15
// relative at T >= 5000 K) and a 1-ulp reciprocal.
template <bool SAFE>
__device__ __forceinline__ double bc_f0(double hnu, double K, double t3, double kT, double inv_kT, double tauBE) {
    double log_boltz = SAFE ? hnu * inv_kT : hnu / kT;
    // expm1(x): for x >= 0.5 (T < 8e4 K at 3646 A) exp(x) - 1 loses at most one bit
    double boltzm1 = log_boltz >= 0.5 ? exp_fast<!SAFE>(log_boltz) - 1.0 : expm1(log_boltz);
    double flam = SAFE ? K * rcp_fast(boltzm1) : K / boltzm1;
    double tau = tauBE * t3;
    double absorption = 1 - exp_fast<!SAFE>(-tau);
    return absorption * flam;
}

// The same sample for the pixel pair (k, k+1) when the arguments of the two exponentials differ by little
// (PlanDev::bc_pair_ok and bc_dhnu_max / kT <= 0.01): with x = h nu / kT and t = tau_BE t3,
//   expm1(x1) = expm1(x0) + expm1(dx) (expm1(x0) + 1),   exp(-t1) = exp(-t0) + exp(-t0) expm1(-dt)
// and expm1 of the small differences by its Taylor series (degree 6: <= 2e-18 relative at |dx| <= 0.01;
// degree 5 at |dt| <= 0.004) -- one full exp pair instead of two.
__device__ __forceinline__ void bc_f0_pair(double2 hnu, double2 K, double2 t3, double inv_kT, double tauBE,
                                           double& f0, double& f1) {
    const double x0 = hnu.x * inv_kT;
    const double bm0 = x0 >= 0.5 ? exp_fast<false>(x0) - 1.0 : expm1(x0);
    const double t0 = tauBE * t3.x;
    const double e0 = exp_fast<false>(-t0);
    f0 = (1 - e0) * (K.x * rcp_fast(bm0));
    const double dx = (hnu.y - hnu.x) * inv_kT;
    double p = fma(dx, 1.0 / 720.0, 1.0 / 120.0);
    p = fma(p, dx, 1.0 / 24.0);
    p = fma(p, dx, 1.0 / 6.0);
    p = fma(p, dx, 0.5);
    p = fma(p, dx, 1.0);
    p = p * dx;                                       // expm1(dx)
    const double bm1 = fma(p, bm0 + 1.0, bm0);
    const double dt = tauBE * (t3.x - t3.y);          // -(t1 - t0)
    double q = fma(dt, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(q, dt, 1.0 / 6.0);
    q = fma(q, dt, 0.5);
    q = fma(q, dt, 1.0);
    q = q * dt;                                       // expm1(-dt)
    const double e1 = fma(e0, q, e0);
    f1 = (1 - e1) * (K.y * rcp_fast(bm1));
}

// log_conv output at data pixel i (BC:225-239 with kernel == [1.0]): the two np.interp
// resamplings collapse to a 4-tap stencil on f0 with walker-independent weights
//   fr_j = f0[k] + (f0[k+1]-f0[k]) tB_j ;  out_i = fr_j + (fr_{j+1}-fr_j) tA_i
// (rebin_spec: True, generic instantiation only: the composed rebin operator's row i instead)
template <bool FULL = false>
__device__ __forceinline__ double bc_conv_at(const PlanDev& pl, const double* __restrict__ f0s, int i) {
    if (!FULL && pl.bc_rb_ptr != nullptr) {
        double acc = 0.0;
        for (int k = pl.bc_rb_ptr[i]; k < pl.bc_rb_ptr[i + 1]; ++k) acc += pl.bc_rb_w[k] * f0s[pl.bc_rb_idx[k]];
        return acc;
    }
    const int4 kk = pl.bc_idx[i];              // k0, k0', k1, k1'
    const double tb0 = pl.bc_tb0[i], tb1 = pl.bc_tb1[i], ta = pl.bc_ta[i];
    double a0 = f0s[kk.x], a1 = f0s[kk.z];
    double fr0 = a0 + (f0s[kk.y] - a0) * tb0;
    double fr1 = a1 + (f0s[kk.w] - a1) * tb1;
    return fr0 + (fr1 - fr0) * ta;
}

struct alignas(16) ClusterW {    // per-walker constants of one cluster
    double cb, kcb2, twoR, twoRkcb, negAs, thr, As;
    double far_lam;     // tasks whose first pixel is at or beyond this wavelength are far (rho <= 0.1)
    double mom[kMom];
};

struct alignas(16) WalkerCtx {   // shared-memory block, one per CTA
    double p[kMaxDim];
    ClusterW cw[kMaxClusters];
    TemplCoef tc[2 * SPAMM_MAX_TEMPLATES];   // [0, n_fe): Fe rows, [n_fe, n_fe + n_host): host rows
    double red[kThreads / 32];
    double nc_norm, nc_slope1, nc_slope2, nc_break;
    double bc_scale;                 // normalization / fnorm
    double kT, tauBE;
    int bc_istar, bpc_istar;
    int prior_ok;
    int bc_finite;       // bc_scale is finite: chunks without Balmer-continuum pixels may skip the 0 * scale
    int task_next;       // PARTS: next task of the CTA's dynamic deal
    int pad_;
};
static_assert(sizeof(WalkerCtx) % 16 == 0, "line table must stay 16-byte aligned");

// utils/find_nearest_index.py:33 starting from the plan-time insertion point of the
// Balmer edge (the per-walker edge moves by < 0.2 A, i.e. a pixel or two)
__device__ int nearest_index_guess(const double* __restrict__ wl, int n, double v, int g) {
    if (isnan(v)) return 0;
    int lo = min(max(g, 0), n);                  // want: first index with wl[lo] >= v
    int steps = 0;
    while (lo > 0 && wl[lo - 1] >= v && steps < 32) { --lo; ++steps; }
    while (lo < n && wl[lo] < v && steps < 32) { ++lo; ++steps; }
    if (steps >= 32) return nearest_index(wl, n, v);
    int c = lo;
    if (c >= n) c = n - 1;
    if (c > 0 && fabs(wl[c - 1] - v) <= fabs(wl[c] - v)) c = c - 1;
    while (c > 0 && fabs(wl[c - 1] - v) == fabs(wl[c] - v)) --c;
    return c;
}

// NuclearContinuumComponent.flux at pixel i (NC:176-201)
template <bool SAFE>
__device__ __forceinline__ double nc_flux(const PlanDev& pl, const WalkerCtx& c, int i, double hi, double lo,
                                          bool broken) {
    if (!broken) {
        // norm * (lam/lam0)^(-slope1)  (NC:195) as exp(-slope1 * ln x) with ln x (hi, lo) held in
        // double-double, so the result is within ~1 ulp of pow()
        double s = -c.nc_slope1;
        double th = s * hi;
        double tl = fma(s, hi, -th) + s * lo;
        double v = exp_fast<!SAFE>(th);
        v = fma(v, tl, v);
        return c.nc_norm * v;
    }
    double lam = pl.wl[i];
    double alpha = lam < c.nc_break ? c.nc_slope1 : c.nc_slope2;              // NC:199
    return c.nc_norm * pow(lam / c.nc_break, -alpha);
}

// sum_t a_t f_t[i] over N template rows starting at tc[t0] for the task's kPix pixels
// (i0 + 32 q): every row load is issued before any is consumed, so a single L2 round trip is
// exposed per block; N is a compile-time count so no predicated-off slots are issued
template <int N>
__device__ __forceinline__ void templ_sum_n(const TemplCoef* __restrict__ tc, int i0, double (&out)[kPix]) {
    double v[N][kPix];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double* f = tc[j].f + i0;
#pragma unroll
        for (int q = 0; q < kPix; ++q) v[j][q] = f[q * 32];
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double a = tc[j].a;
        // one FMA per term (numpy rounds the product first: <= 0.5 ulp per term apart)
#pragma unroll
        for (int q = 0; q < kPix; ++q) out[q] = fma(a, v[j][q], out[q]);                 // Fe:334 / HG:339
    }
}

// two families with compile-time counts (FULL model): all NA + NB row loads in one batch
// (PAIR: the lane's two pixels are adjacent, one 16-byte load per row)
template <int NA, int NB, bool PAIR>
__device__ __forceinline__ void templ_sum_2(const TemplCoef* __restrict__ tc, int i0, double (&outa)[kPix],
                                            double (&outb)[kPix]) {
    double v[NA + NB][kPix];
#pragma unroll
    for (int j = 0; j < NA + NB; ++j) {
        const double* f = tc[j].f + i0;
        if (PAIR) {
            const double2 t = *reinterpret_cast<const double2*>(f);
            v[j][0] = t.x; v[j][kPix - 1] = t.y;
        } else {
#pragma unroll
            for (int q = 0; q < kPix; ++q) v[j][q] = f[q * 32];
        }
    }
#pragma unroll
    for (int q = 0; q < kPix; ++q) { outa[q] = 0.0; outb[q] = 0.0; }
#pragma unroll
    for (int j = 0; j < NA; ++j) {
        const double a = tc[j].a;
#pragma unroll
        for (int q = 0; q < kPix; ++q) outa[q] = fma(a, v[j][q], outa[q]);               // Fe:334
    }
#pragma unroll
    for (int j = NA; j < NA + NB; ++j) {
        const double a = tc[j].a;
#pragma unroll
        for (int q = 0; q < kPix; ++q) outb[q] = fma(a, v[j][q], outb[q]);               // HG:339
    }
}

__device__ __forceinline__ void templ_sum(const WalkerCtx& c, int t0, int n, int i0, double (&out)[kPix]) {
#pragma unroll
    for (int q = 0; q < kPix; ++q) out[q] = 0.0;
    const TemplCoef* tc = c.tc + t0;
    while (n >= 8) { templ_sum_n<8>(tc, i0, out); tc += 8; n -= 8; }
    switch (n) {
        case 7: templ_sum_n<7>(tc, i0, out); break;
        case 6: templ_sum_n<6>(tc, i0, out); break;
        case 5: templ_sum_n<5>(tc, i0, out); break;
        case 4: templ_sum_n<4>(tc, i0, out); break;
        case 3: templ_sum_n<3>(tc, i0, out); break;
        case 2: templ_sum_n<2>(tc, i0, out); break;
        case 1: templ_sum_n<1>(tc, i0, out); break;
        default: break;
    }
}

// sum of Lorentzians a/(d^2+w^2) over line groups [g0, g1): four lines share one reciprocal
__device__ __forceinline__ void lorentz_direct(const double2* __restrict__ sL2, int g0, int g1,
                                               const double (&lam)[kPix], double (&acc)[kPix]) {
#pragma unroll 1
    for (int g = g0; g < g1; ++g) {
        const double2 c01 = sL2[g * 6 + 0], c23 = sL2[g * 6 + 1];
        const double2 w01 = sL2[g * 6 + 2], w23 = sL2[g * 6 + 3];
        const double2 a01 = sL2[g * 6 + 4], a23 = sL2[g * 6 + 5];
#pragma unroll
        for (int q = 0; q < kPix; ++q) {
            double d0 = lam[q] - c01.x, d1 = lam[q] - c01.y;
            double d2 = lam[q] - c23.x, d3 = lam[q] - c23.y;
            double x0 = fma(d0, d0, w01.x), x1 = fma(d1, d1, w01.y);
            double x2 = fma(d2, d2, w23.x), x3 = fma(d3, d3, w23.y);
            double p01 = x0 * x1, p23 = x2 * x3;
            double n01 = fma(a01.x, x1, a01.y * x0);
            double n23 = fma(a23.x, x3, a23.y * x2);
            double num = fma(n01, p23, n23 * p01);
            double den = p01 * p23;
            acc[q] = fma(num, rcp_fast(den), acc[q]);
        }
    }
}

// min over the warp's pixels of C = u^2 + kappa c_bar^2 (squared distance to the cluster's
// pole line, up to the factor 1/(1+kappa)); warp-uniform result
__device__ __forceinline__ double cluster_cmin(const ClusterW& cw, const double (&lam)[kPix]) {
    double cmin = CUDART_INF;
#pragma unroll
    for (int q = 0; q < kPix; ++q) {
        double u = lam[q] - cw.cb;
        cmin = fmin(cmin, fma(u, u, cw.kcb2));
    }
    for (int o = 16; o > 0; o >>= 1) cmin = fmin(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
    return cmin;
}

// sum_n a_n / q_n(lam) for one cluster as sum_k t_k M_k:
//   1/q(delta) = sum_k t_k (delta/R)^k,  t_0 = 1/C, t_1 = beta t_0, t_k = beta t_{k-1} + alpha t_{k-2}
//   with q(delta) = As (delta/R)^2 - 2R(u - kappa c_bar)(delta/R) + C
__device__ __forceinline__ void cluster_series(const ClusterW& cw, int K, const double (&lam)[kPix],
                                               double (&acc)[kPix]) {
    // Clenshaw summation of sum_k M_k t_k for t_k = beta t_{k-1} + alpha t_{k-2}, t_0 = 1/C, t_1 = beta t_0:
    //   b_k = M_k + beta b_{k+1} + alpha b_{k+2},  sum = t_0 b_0     (2 FMA per term and pixel)
    double al[kPix], be[kPix], invC[kPix], bh[kPix], bl[kPix];
    const double mh = cw.mom[K - 1], ml = cw.mom[K - 2];
#pragma unroll
    for (int q = 0; q < kPix; ++q) {
        double u = lam[q] - cw.cb;
        invC[q] = rcp_fast(fma(u, u, cw.kcb2));
        be[q] = fma(u, cw.twoR, -cw.twoRkcb) * invC[q];
        al[q] = cw.negAs * invC[q];
        bh[q] = mh;                                   // b_{K-1}
        bl[q] = fma(be[q], mh, ml);                   // b_{K-2}
    }
    // two terms per trip (K is even): bh <- b_k, bl <- b_{k-1}
#pragma unroll 1
    for (int k = K - 3; k > 0; k -= 2) {
        const double mk = cw.mom[k], mk1 = cw.mom[k - 1];
#pragma unroll
        for (int q = 0; q < kPix; ++q) {
            bh[q] = fma(be[q], bl[q], fma(al[q], bh[q], mk));
            bl[q] = fma(be[q], bh[q], fma(al[q], bl[q], mk1));
        }
    }
#pragma unroll
    for (int q = 0; q < kPix; ++q) acc[q] = fma(invC[q], bl[q], acc[q]);
}

// number of series terms for rho^2 = As / Cmin  (rho^K <~ 1e-16; capped at kMom, which the
// far test rho <= 0.1 makes sufficient)
// (single precision throughout: K only has to be large enough, and the +2 absorbs the rounding)
__device__ __forceinline__ int series_terms(double As, double cmin) {
    float l = __logf((float)cmin) - __logf((float)As);          // -ln(rho^2) >= ln 100
    int K = ((int)__fdividef(74.0f, l) + 2) & ~1;                 // even
    return min(max(K, 4), kMom);
}

#ifndef SPAMM_MIN_BLOCKS
#define SPAMM_MIN_BLOCKS 4
#endif
// kernel-iteration builds (-DSPAMM_TRACE): per-CTA phase time stamps, read back by tools/trace_ctas.py
#ifdef SPAMM_TRACE
__device__ unsigned long long* g_trace = nullptr;
__device__ __forceinline__ void trace_mark(int slot) {
    if (threadIdx.x == 0 && g_trace != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_trace[(size_t)blockIdx.x * 8 + slot] = t;
        if (slot == 0) {
            unsigned sm;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            g_trace[(size_t)blockIdx.x * 8 + 7] = sm;
        }
    }
}
#define SPAMM_MARK(slot) trace_mark(slot)
#else
#define SPAMM_MARK(slot)
#endif
constexpr int kChunk = 32 * kPix;      // pixels per warp task

// MODE 0: ln-posterior, 1: model flux out.  CANON: the live components are in the reference's
// canonical order with the Balmer component last (run_spamm.py:100-122), so the model is
// ((NC + Fe) + host) + Balmer and no per-component ordering state is carried through the task.
// SPLIT: small ensembles -- one walker per thread-block CLUSTER: every CTA of the cluster builds the
// per-walker tables, the 64-pixel tasks are dealt round-robin over the CTAs, and CTA 0 gathers the
// per-chunk chi^2 partials from its peers' shared memory (DSMEM) in the same fixed chunk order.
// FULL: compile-time view of the headline model (run_spamm's NC + Fe x3 + host x7 + BalmerCombined with
// both parts on, Lorentzian lines, far-field series, bank rows): every component test, template count
// and line-type branch below folds away, which roughly halves the non-FP64 instructions of a task.
constexpr int kFullFe = 3, kFullHost = 7, kFullClusters = 4, kFullSuper = 3;
// SPEC = 0: generic.  Otherwise a bit set of the components the instantiation is compiled for (NC is
// always bit 0; kSpecFe: Fe x3, kSpecHost: host x7, kSpecBalmer: BalmerCombined with both parts on,
// Lorentzian + series, kSpecExt: a trailing Extinction); the supported sets are listed in kSpecs.
constexpr int kSpecNc = 1, kSpecFe = 2, kSpecHost = 4, kSpecBalmer = 8, kSpecExt = 16;
// modifiers of kSpecBalmer: BalmerCombined with one part switched off -- run_spamm's "BC" without "BPC" is the
// continuum alone (run_spamm.py:118-121; the reference's own examples/test_spamm.py fits that model)
constexpr int kSpecNoBpc = 32, kSpecNoBc = 64;
// launch-shape modifier: one walker = a red-part CTA + a blue-part CTA (PartsWs); full Balmer models only
constexpr int kSpecParts = 128;
// NC | NC+Fe | NC+host | NC+Fe+host | NC+Balmer | NC+Fe+Balmer | NC+host+Balmer | full | full+Extinction |
// NC+Balmer and full with the continuum only / the lines only
// (SPAMM_DEV_BUILD: kernel-iteration builds compile the config-5 instantiation only)
#ifdef SPAMM_DEV_BUILD
#define SPAMM_SPEC_LIST(X) X(15)
#define SPAMM_PARTS_LIST(X) X(15)
#else
#define SPAMM_PARTS_LIST(X) X(9) X(11) X(13) X(15) X(31)     // models with both Balmer parts
#define SPAMM_SPEC_LIST(X) X(1) X(3) X(5) X(7) X(9) X(11) X(13) X(15) X(31) X(41) X(73) X(47) X(79)
#endif
#define SPAMM_SPEC_ENTRY(S) S,
constexpr int kSpecs[] = {SPAMM_SPEC_LIST(SPAMM_SPEC_ENTRY)};
#undef SPAMM_SPEC_ENTRY
static_assert((9 | kSpecNoBpc) == 41 && (9 | kSpecNoBc) == 73 && (15 | kSpecNoBpc) == 47 && (15 | kSpecNoBc) == 79,
              "SPAMM_SPEC_LIST spells the Balmer-part modifiers as numbers");
template <int MODE, bool CANON, bool SPLIT, int SPECP>
__global__ void __launch_bounds__(kThreads, SPAMM_MIN_BLOCKS)
k_walkers(const PlanDev pl, const double* __restrict__ params, int w0, double* __restrict__ out,
          uint32_t mask, DirectRows dr, PeerOut po, PartsWs pw, Proposal pr) {
    constexpr int SPEC = SPECP & ~kSpecParts;
    constexpr bool PARTS = (SPECP & kSpecParts) != 0;
    static_assert(!PARTS || (!SPLIT && MODE == 0 && (SPEC & kSpecBalmer) && !(SPEC & (kSpecNoBpc | kSpecNoBc))),
                  "walker parts: ln-posterior of a specialised model with both Balmer parts, no cluster split");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WalkerCtx& c = *reinterpret_cast<WalkerCtx*>(smem_raw);
    double* sLine = reinterpret_cast<double*>(smem_raw + sizeof(WalkerCtx));   // [n_pad/4][12] grouped
    double* sG = sLine + 3 * pl.n_lines_pad;                                    // [n_lines_pad]
    double* sW = sG + pl.n_lines_pad;                                           // [n_lines_pad]
    double* sChi = sW + pl.n_lines_pad;                                         // [n_chunks]
    double* sF0 = sChi + pl.n_chunks;                                           // [bc_nstage]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    SPAMM_MARK(0);
    int csize = 1, crank = 0;
    if (SPLIT) {
        cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
        csize = (int)cl.num_blocks();
        crank = (int)cl.block_rank();
    }
    int wloc = SPLIT ? blockIdx.x / csize : blockIdx.x;   // walker within this launch
    // PARTS: is_red builds the line table / moments and owns chunks >= part_chunk, is_blue stages the whole
    // Balmer continuum and owns the chunks before it (both true: the whole walker in one CTA)
    bool is_red = true, is_blue = true;
    if (PARTS && (int)blockIdx.x >= pw.n_whole) {
        const int b = (int)blockIdx.x - pw.n_whole;
        is_red = b < pw.n_parts;
        is_blue = !is_red;
        wloc = pw.n_whole + (is_red ? b : b - pw.n_parts);
    }
    const bool only_red = PARTS && !is_blue, only_blue = PARTS && !is_red;
    const int w = w0 + wloc;
    const int P = pl.P;
    const double* pwalk = params + (size_t)w * pl.D;

    // The first CTAs stream the template bank into L2 (it is demand-fetched row by row otherwise:
    // ~20 us on a cold L2, i.e. after another kernel evicted it)
    if (SPEC != 0 && blockIdx.x < 128 && tid < 2) {
        const TemplDev& ts = tid == 0 ? pl.fe : pl.host;
        const int rows = tid == 0 ? pl.fe_rows : pl.host_rows;
        if (ts.bank_f != nullptr && rows > 0) {
            const size_t total = (size_t)rows * P * sizeof(double);
            const size_t slice = ((total + 127) / 128 + 127) & ~(size_t)127;      // bytes per CTA, 128-aligned
            size_t off = slice * blockIdx.x;
            const size_t end = off + slice < total ? off + slice : total;
            for (; off < end; off += 32768) {
                const unsigned n = (unsigned)((end - off < 32768 ? end - off : 32768) & ~(size_t)15);
                if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::
                                    "l"(reinterpret_cast<const char*>(ts.bank_f) + off), "r"(n) : "memory");
            }
        }
    }
    // ---- P0: parameters + flat-box prior (Model.prior, Model.py:449-470) -----------------
    if (tid == 0) { c.prior_ok = 1; c.task_next = 0; }
    __syncthreads();
    if (tid < pl.D) {
        double v;
        if (pr.walkers != nullptr) {
            const int i = pr.i0 + w, Ns = pr.K / 2;
            double zz;
            int partner;
            stretch_draw(pr.seed, i, pr.half, pr.iter, pr.a, pr.K, zz, partner);
            const double sd = pr.walkers[(size_t)(pr.half == 0 ? i : Ns + i) * pl.D + tid];
            const double cd = pr.walkers[(size_t)partner * pl.D + tid];
            v = cd - zz * (cd - sd);
        } else {
            v = pwalk[tid];
        }
        c.p[tid] = v;

        if (MODE == 0 && pl.lo != nullptr) {
            if (!(pl.lo[tid] < v && v < pl.hi[tid])) c.prior_ok = 0;   // strict; NaN fails
        }
    }
    __syncthreads();
    if (MODE == 0 && !c.prior_ok) {       // ln_posterior returns -inf before any flux (Model.py:67-69)
        if (tid == 0 && crank == 0 && is_red) store_result(out, w, -CUDART_INF, po);
        return;
    }

    SPAMM_MARK(1);
    // ---- which components are live ------------------------------------------------------
    int off_nc = -1, off_bal = -1, q_bal = pl.n_comp;
    bool has_fe = false, has_host = false;
    for (int q = 0; q < pl.n_comp; ++q) {
        if (!((mask >> q) & 1u)) continue;
        int kind = pl.comp_kind[q];
        if (kind == SPAMM_COMP_NUCLEAR) off_nc = pl.comp_off[q];
        else if (kind == SPAMM_COMP_FE) has_fe = true;
        else if (kind == SPAMM_COMP_HOST) has_host = true;
        else if (kind == SPAMM_COMP_BALMER) { off_bal = pl.comp_off[q]; q_bal = q; }
    }
    // Extinction is the last component (run_spamm.py:123-125) and multiplies the flux summed so far
    // (Model.py:358-361, 384-395); selected alone (component flux API) it returns the curve itself
    constexpr bool FULL = SPEC != 0;
    constexpr int NFE = (SPEC & kSpecFe) ? kFullFe : 0, NHOST = (SPEC & kSpecHost) ? kFullHost : 0;
    const bool has_ext = FULL ? (SPEC & kSpecExt) != 0 : pl.ext_k != nullptr && ((mask >> pl.ext_q) & 1u);
    const bool ext_alone = has_ext && (mask & ~(1u << pl.ext_q)) == 0u;
    if (FULL) { has_fe = NFE > 0; has_host = NHOST > 0; }
    const bool has_nc = FULL ? true : off_nc >= 0;
    const bool nc_broken = FULL ? false : pl.nc_broken != 0;
    const bool have_balmer = FULL ? (SPEC & kSpecBalmer) != 0 : off_bal >= 0;
    const int n_fe = FULL ? NFE : (has_fe ? pl.fe.n_templ : 0);
    const int n_host = FULL ? NHOST : (has_host ? pl.host.n_templ : 0);
    const bool do_bc = FULL ? have_balmer && (SPEC & kSpecNoBc) == 0 : have_balmer && pl.bc_on;
    const bool do_bpc = FULL ? have_balmer && (SPEC & kSpecNoBpc) == 0 : have_balmer && pl.bpc_on;
    const bool lorentz = FULL ? true : pl.line_type == SPAMM_LINE_LORENTZIAN;
    const bool series_on = FULL ? true : pl.series_on != 0;
    constexpr bool PAIR = FULL && kPix == 2;      // adjacent-pixel lanes with 16-byte loads
    constexpr int QS = PAIR ? 1 : 32;             // pixel stride between a lane's pixels
    const int n_base = FULL ? kFullClusters : pl.n_clusters, n_super = FULL ? kFullSuper : pl.n_super;
    const bool fe_rows_direct = FULL ? false : dr.fe_f != nullptr;
    const bool host_rows_direct = FULL ? false : dr.host_f != nullptr;

    double b_norm = 0, b_Te = 0, b_tau = 0, b_loff = 0, b_lwid = 0, b_logNe = 0;
    if (have_balmer) {
        b_norm = c.p[off_bal]; b_Te = c.p[off_bal + 1]; b_tau = c.p[off_bal + 2];
        b_loff = c.p[off_bal + 3]; b_lwid = c.p[off_bal + 4]; b_logNe = c.p[off_bal + 5];
    }
    const int n_lines = pl.n_lines, n_pad = pl.n_lines_pad;
    const double kT = pl.kb * b_Te;                      // k.value*T, BC:296
    // quotients used by the line table (P1) and the cluster constants (P3): divided once per thread
    const double bal_shift = b_loff / pl.c_kms;          // BC:379
    const double bal_width = b_lwid / pl.c_kms;

    // ---- P1: independent per-walker scalars, spread over the warps ------------------------
    if (tid == 0 && has_nc) {
        if (!nc_broken) {
            c.nc_norm = c.p[off_nc]; c.nc_slope1 = c.p[off_nc + 1];
            c.nc_slope2 = 0; c.nc_break = 0;
        } else {
            c.nc_break = c.p[off_nc]; c.nc_norm = c.p[off_nc + 1];
            c.nc_slope1 = c.p[off_nc + 2]; c.nc_slope2 = c.p[off_nc + 3];
        }
    }
    constexpr int kNW = kThreads / 32;                      // warps of the CTA (roles below wrap around)
    if (has_fe && warp == 1 % kNW && lane < n_fe) {
        const TemplDev& ts = pl.fe;
        int t = lane;
        double norm = c.p[ts.param_off + t];
        const double* f; double nf;
        if (fe_rows_direct) {
            size_t row = (size_t)wloc * ts.n_templ + t;
            f = dr.fe_f + row * P; nf = dr.fe_nf[row];
        } else {
            double sn = sigma_norm_of(ts, t, c.p[ts.param_off + ts.n_templ]);
            int sig = 1;
            if (sn >= 1.0 && sn <= (double)ts.bank_sig_max[t]) sig = (int)sn;
            else { flag_status(pl.status, pl.status_host, kStatSigmaRange); norm = CUDART_NAN; }
            int row = ts.bank_row0[t] + sig - 1;
            f = ts.bank_f + (size_t)row * P; nf = ts.bank_nf[row];
        }
        c.tc[t].f = f;
        c.tc[t].a = norm / nf;                                          // Fe:334
    }
    if (has_host && warp == 2 % kNW && lane < n_host) {
        const TemplDev& ts = pl.host;
        int t = lane;
        double norm = c.p[ts.param_off + t];
        const double* f; double nf;
        if (host_rows_direct) {
            size_t row = (size_t)wloc * ts.n_templ + t;
            f = dr.host_f + row * P; nf = dr.host_nf[row];
        } else {
            double sn = sigma_norm_of(ts, t, c.p[ts.param_off + ts.n_templ]);
            int sig = 1;
            if (sn >= 1.0 && sn <= (double)ts.bank_sig_max[t]) sig = (int)sn;
            else { flag_status(pl.status, pl.status_host, kStatSigmaRange); norm = CUDART_NAN; }
            int row = ts.bank_row0[t] + sig - 1;
            f = ts.bank_f + (size_t)row * P; nf = ts.bank_nf[row];
        }
        c.tc[n_fe + t].f = f;
        c.tc[n_fe + t].a = norm / nf;                                        // HG:339
    }
    if (warp == 3 % kNW && lane == 0) {
        c.bc_istar = -1;
        c.kT = kT; c.tauBE = b_tau;
        if (do_bc) {
            double edge_wl = pl.edge * (1 - b_loff / pl.c_cgs);          // BC:432 (cm/s!)
            int is = nearest_index_guess(pl.wl, P, edge_wl, pl.edge_guess);   // BC:443
            c.bc_istar = is;
            // log_conv kernel: kernel_width = round(5*dpix) must be 0          BC:229-233
            double dpix = (b_lwid / pl.c_cgs) / pl.bc_dlnew;
            if (rint(5 * dpix) != 0.0) flag_status(pl.status, pl.status_host, kStatBcKernel);
            const int4 kk = pl.bc_idx[is];
            int top = max(kk.y, kk.w);
            if (!FULL && pl.bc_rb_ptr != nullptr) top = pl.bc_rb_idx[pl.bc_rb_ptr[is + 1] - 1];   // columns ascend
            // (reported as an error by the host; no Balmer-continuum pixel is evaluated so nothing reads past the stage)
            if (top >= pl.bc_nstage) { flag_status(pl.status, pl.status_host, kStatBcStage); c.bc_istar = -1; }
        }
    }
    if (warp == 4 % kNW && lane == 0) {
        c.bpc_istar = -1;
        if (do_bpc) {
            double edge_wl = pl.edge * (1 - b_loff / pl.c_kms);          // BC:374
            c.bpc_istar = nearest_index_guess(pl.wl, P, edge_wl, pl.edge_guess);   // BC:382
        }
    }
    if (do_bpc && is_red) {
        const double shift = bal_shift, width = bal_width;
        const double inv_kT_l = 1.0 / kT;
        // BC:376; only the SH95 look-ups (n <= 50, the first threads) use it
        const double n_e = pl.line_min + tid <= 50 ? pow(10.0, b_logNe) : 0.0;
        for (int n = tid; n < n_pad; n += kThreads) {
            double lc = 0.0, wv = 0.0, rg = 0.0;
            if (n < n_lines) {
                double c0 = pl.line_c0[n];
                lc = c0 - shift * c0;                    // lcenter -= shift*lcenter, BC:314
                wv = width * lc;                         // BC:316
                int nn = pl.line_min + n;
                if (nn <= 50) rg = sh95_bilinear(pl, 50 - nn, n_e, b_Te);   // coef_use[-i+2], BC:294
                else rg = FULL ? exp_fast<false>(pl.line_cum[n] * inv_kT_l)    // |S_n / kT| <= 0.2 (PlanDev::mom_poly)
                               : exp(pl.line_cum[n] / kT);                    // BC:296, telescoped
            }
            sLine[(n >> 2) * 12 + (n & 3)] = lc;
            sW[n] = wv;
            sG[n] = rg;      // ratio (n <= 50) or r_n / r_50 (n > 50)
        }
    }
    // Balmer-continuum samples f0 for the staged pixels; each trip issues the next trip's loads before
    // its own arithmetic (no trip is issued for pixels beyond the stage, unlike a batched loop)
    if (do_bc) {
        const int nst = min(pl.bc_nstage, P);
        const double inv_kT = 1.0 / kT;
        if (FULL && pl.bc_pair_ok && pl.bc_dhnu_max * inv_kT <= 0.01) {
            // adjacent pixel pairs, one pair per trip, the next pair's loads issued before this pair's arithmetic
            // (a red part needs the samples under its own Balmer-continuum pixels and the normaliser only)
            int k = (only_red ? pl.part_stage0 : 0) + 2 * tid;
            double2 h, kk, t3;
            if (k < nst) {
                h = *reinterpret_cast<const double2*>(pl.bc_hnu + k);
                kk = *reinterpret_cast<const double2*>(pl.bc_K + k);
                t3 = *reinterpret_cast<const double2*>(pl.bc_t3 + k);
            }
#pragma unroll 1
            for (; k < nst; k += 2 * kThreads) {
                const int kn = min(k + 2 * kThreads, (nst - 1) & ~1);
                const double2 hn = *reinterpret_cast<const double2*>(pl.bc_hnu + kn);
                const double2 kkn = *reinterpret_cast<const double2*>(pl.bc_K + kn);
                const double2 t3n = *reinterpret_cast<const double2*>(pl.bc_t3 + kn);
                double f0, f1;
                bc_f0_pair(h, kk, t3, inv_kT, b_tau, f0, f1);
                sF0[k] = f0;
                if (k + 1 < nst) sF0[k + 1] = f1;
                h = hn; kk = kkn; t3 = t3n;
            }
        } else {
            // one pixel per trip, the next pixel's loads issued before this pixel's arithmetic
            int k = (only_red ? pl.part_stage0 : 0) + tid;
            double h = 0.0, kk = 0.0, t3 = 0.0;
            if (k < nst) { h = pl.bc_hnu[k]; kk = pl.bc_K[k]; t3 = pl.bc_t3[k]; }
#pragma unroll 1
            for (; k < nst; k += kThreads) {
                const int kn = min(k + kThreads, nst - 1);
                const double hn = pl.bc_hnu[kn], kkn = pl.bc_K[kn], t3n = pl.bc_t3[kn];
                sF0[k] = bc_f0<FULL>(h, kk, t3, kT, inv_kT, b_tau);
                h = hn; kk = kkn; t3 = t3n;
            }
        }
    }
    __syncthreads();

    SPAMM_MARK(2);
    // (the ratio recursion flux_ratios[i] = flux_ratios[i-1] * exp(e_i / kT), BC:292-296, is telescoped:
    //  r_n = r_50 * exp(sum_{m=51..n} e_m / kT); line_cum holds the partial sums, P3 applies r_50)
    // ---- P3: finish the line table; BC and BpC normalisers; cluster moments -------------------------
    const int bc_istar = c.bc_istar, bpc_istar = c.bpc_istar;
    double bpc_scale = 0.0;
    if (do_bc && warp == kNW - 1 && lane == 0) {
        double fnorm = bc_istar >= 0 ? bc_conv_at<FULL>(pl, sF0, bc_istar) : CUDART_NAN;     // BC:444
        c.bc_scale = b_norm / fnorm;                                     // BC:446
        c.bc_finite = isfinite(c.bc_scale) ? 1 : 0;
    }
    if (do_bpc && is_red) {
        const int n50 = min(max(51 - pl.line_min, 0), n_lines);       // first line with n > 50
        const double r50 = n50 > 0 ? sG[n50 - 1] : 0.0;               // flux_ratios[-1] of a zeroed array
        // line table entries (w^2, a = ratio * w) and, on the way, the BpC normaliser = the line sum
        // at the edge pixel (BC:385)
        const double lam_e = pl.wl[bpc_istar];
        double s = 0.0;
        for (int n = tid; n < n_pad; n += kThreads) {
            const int gq = n >> 2, r4 = n & 3;
            const double wv = sW[n];
            double rr = sG[n];
            if (n >= n50) rr = r50 * rr;
            double w2 = 1.0, a = 0.0;                                 // padding: contributes exactly 0
            if (n < n_lines) {
                w2 = wv * wv;
                a = lorentz ? rr * wv : rr;
                const double d = lam_e - sLine[gq * 12 + r4];
                if (lorentz) s += FULL ? a * rcp_fast(fma(d, d, w2)) : a / (d * d + w2);
                else s += a * exp_fast(-(d * d) / w2);
            }
            sLine[gq * 12 + 4 + r4] = w2;
            sLine[gq * 12 + 8 + r4] = a;
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) c.red[warp] = s;
        // cluster moments M_k = sum_n a_n d_n^k and per-walker cluster constants: one warp each.
        // FULL with pl.mom_poly: the n > 50 lines, whose ratios are r_50 exp(S_n / kT), enter through
        // the walker-independent power series in z = S_ref / kT (mom_G), so only the n <= 50 lines of
        // the cluster are visited.
        const int n_cl = series_on && lorentz ? n_base + n_super : 0;
        for (int kcl = warp; kcl < n_cl; kcl += kNW) {
            const ClusterDev& cd = pl.cl[kcl];
            const double* dd = pl.line_d[kcl];
            const bool poly = FULL && pl.mom_poly;
            const int n_end = poly ? min(cd.g_hi * 4, n50) : cd.g_hi * 4;
            double mk[kMom];
#pragma unroll
            for (int k = 0; k < kMom; ++k) mk[k] = 0.0;
            for (int n = cd.g_lo * 4 + lane; n < n_end; n += 32) {
                double rr = sG[n];
                if (n >= n50) rr = r50 * rr;
                double pw = rr * sW[n];                               // a_n = r_n w_n (0 for padding)
                const double dn = dd[n];
#pragma unroll
                for (int k = 0; k < kMom; ++k) { mk[k] += pw; pw *= dn; }
            }
            // lane sums of all 16 moments by a halving butterfly (16 shuffled doubles instead of 80;
            // same xor tree 16, 8, 4, 2, 1 per moment): lanes 2k, 2k+1 end up with moment k
            static_assert(kMom == 16, "butterfly below is written for 16 moments");
#pragma unroll
            for (int h = 8, bit = 16; h >= 1; h >>= 1, bit >>= 1) {
                const bool up = (lane & bit) != 0;
#pragma unroll
                for (int k = 0; k < h; ++k) {
                    const double send = up ? mk[k] : mk[k + h];
                    const double keep = up ? mk[k + h] : mk[k];
                    mk[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
                }
            }
            mk[0] += __shfl_xor_sync(0xffffffffu, mk[0], 1);
            if (poly) {
                const double* G = pl.mom_G[kcl] + (lane >> 1) * kMomPoly;
                double g[kMomPoly];
#pragma unroll
                for (int q = 0; q < kMomPoly; ++q) g[q] = G[q];
                const double z = pl.mom_Sref / kT;
                double acc = g[kMomPoly - 1];
#pragma unroll
                for (int q = kMomPoly - 2; q >= 0; --q) acc = fma(acc, z, g[q]);
                const double shift = bal_shift, width = bal_width;
                mk[0] = fma(r50 * (width * (1.0 - shift)), acc, mk[0]);
            }
            if (!(lane & 1)) c.cw[kcl].mom[lane >> 1] = mk[0];
            if (lane == 0) {
                const double shift = bal_shift, width = bal_width;
                const double kappa = width * width;
                const double cb = cd.cb0 - shift * cd.cb0, R = cd.R0 - shift * cd.R0;
                ClusterW& cw = c.cw[kcl];
                cw.cb = cb;
                cw.kcb2 = kappa * cb * cb;
                cw.twoR = 2.0 * R;
                cw.twoRkcb = 2.0 * R * kappa * cb;
                cw.As = (1.0 + kappa) * R * R;
                cw.negAs = -cw.As;
                cw.thr = 100.0 * cw.As;                              // rho <= 0.1
                // redward of c_bar the nearest pixel of a task is its first one: u^2 + kcb2 >= thr
                cw.far_lam = cb + sqrt(fmax(cw.thr - cw.kcb2, 0.0)) * (1.0 + 1e-12);
            }
        }
    }
    __syncthreads();
    if (do_bpc && is_red) {
        double tot = 0.0;
        for (int q = 0; q < kThreads / 32; ++q) tot += c.red[q];
        bpc_scale = b_norm / tot;                                        // BC:387
    }
    // (a blue part holds no pixel beyond the Balmer edge: its pseudo-continuum term is 0 * bpc_scale = 0,
    //  bpc_scale being positive and finite inside the prior box of a specialised model)
    SPAMM_MARK(3);
    const double bc_scale = c.bc_scale;
    const bool bc_finite = c.bc_finite != 0;
    double* fout = (MODE == 1) ? out + (size_t)wloc * P : nullptr;

    // ---- main phase: 64-pixel warp tasks, reddest (line-heavy) chunks first -------------------------
    const int n_chunks = pl.n_chunks;
    const int n_groups = n_pad >> 2;
    const double2* sL2 = reinterpret_cast<const double2*>(sLine);
    // (static round-robin.  A shared task counter costs a dozen instructions per task -- ptxas turns even
    // a plain atom.shared into its warp-aggregation sequence -- and measured 2.7 % slower, 1 % slower
    // when tasks are fetched in pairs, although the one task inside the line forest, ~13x the work of
    // the others, leaves its warp behind: the SM's other CTAs fill the gap.)
    const int chunk_top = only_blue ? pl.part_chunk : n_chunks;               // this CTA owns chunks
    const int chunk_bot = only_red ? pl.part_chunk : 0;                       //   [chunk_bot, chunk_top)
    // PARTS (ensembles of a few waves, where nothing fills the gap a straggling warp leaves): tasks are dealt
    // dynamically, and a red part starts with its chunks next to the Balmer edge -- inside the line forest, an
    // order of magnitude more work than a far chunk -- so the CTA's tail is made of light tasks
    for (int t = SPLIT ? warp * csize + crank : warp;; t += (kThreads / 32) * csize) {
        if (PARTS) {
            if (lane == 0) t = atomicAdd(&c.task_next, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
        }
        if (t >= chunk_top - chunk_bot) break;
        const int chunk = only_red ? chunk_bot + t : chunk_top - 1 - t;
        // pixels i0 + 32 q; per-pixel arrays carry kChunk elements of slack, so no clamping
        // PAIR (FULL kernel): lane l holds the adjacent pixels 2l, 2l+1 of the chunk, fetched with 16-byte
        // loads; otherwise pixels l and l + 32
        const int i0 = chunk * kChunk + (PAIR ? 2 * lane : lane);
        double m[kPix], post[kPix], acc[kPix];
        // every global load of the task is issued here, ahead of its first use, so one L2 round
        // trip is exposed per task instead of one per phase
        double lnhi[kPix], lnlo[kPix], lam[kPix], dat[kPix], ierr[kPix];
        if (PAIR) {
            const double2 a = *reinterpret_cast<const double2*>(pl.nc_lnx_hi + i0);
            const double2 b = *reinterpret_cast<const double2*>(pl.nc_lnx_lo + i0);
            const double2 l = *reinterpret_cast<const double2*>(pl.wl + i0);
            const double2 d = *reinterpret_cast<const double2*>(pl.flux + i0);
            const double2 e = *reinterpret_cast<const double2*>(pl.inv_err + i0);
            lnhi[0] = a.x; lnhi[kPix - 1] = a.y; lnlo[0] = b.x; lnlo[kPix - 1] = b.y;
            lam[0] = l.x; lam[kPix - 1] = l.y; dat[0] = d.x; dat[kPix - 1] = d.y;
            ierr[0] = e.x; ierr[kPix - 1] = e.y;
        }
#pragma unroll
        for (int q = 0; q < kPix; ++q) {
            acc[q] = 0.0;
            post[q] = 0.0;
            if (PAIR) continue;
            lnhi[q] = has_nc ? pl.nc_lnx_hi[i0 + 32 * q] : 0.0;
            lnlo[q] = has_nc ? pl.nc_lnx_lo[i0 + 32 * q] : 0.0;
            lam[q] = pl.wl[i0 + 32 * q];
            dat[q] = MODE == 0 ? pl.flux[i0 + 32 * q] : 0.0;
            ierr[q] = MODE == 0 ? pl.inv_err[i0 + 32 * q] : 0.0;
        }
        {
            double fe_v[kPix], host_v[kPix], nc_v[kPix];
            if (FULL) {
                if constexpr (NFE + NHOST > 0) templ_sum_2<NFE, NHOST, PAIR>(c.tc, i0, fe_v, host_v);
            } else {
                if (n_fe) templ_sum(c, 0, n_fe, i0, fe_v);
                if (n_host) templ_sum(c, n_fe, n_host, i0, host_v);
            }
            if (PAIR) {
                // second pixel of the pair from the first: x1^-s = x0^-s exp(z), z = -s (ln x1 - ln x0);
                // |z| <= 0.016 (PlanDev::nc_pair_ok), degree-6 Taylor: <= 4e-17
                nc_v[0] = nc_flux<true>(pl, c, i0, lnhi[0], lnlo[0], false);
                const double ms = -c.nc_slope1;
                const double z = fma(ms, lnlo[kPix - 1] - lnlo[0], ms * (lnhi[kPix - 1] - lnhi[0]));
                double pz = fma(z, 1.0 / 720.0, 1.0 / 120.0);
                pz = fma(pz, z, 1.0 / 24.0);
                pz = fma(pz, z, 1.0 / 6.0);
                pz = fma(pz, z, 0.5);
                pz = fma(pz, z, 1.0);
                pz = fma(pz, z, 1.0);
                nc_v[kPix - 1] = nc_v[0] * pz;
            } else {
#pragma unroll
                for (int q = 0; q < kPix; ++q)
                    nc_v[q] = has_nc ? nc_flux<FULL>(pl, c, i0 + QS * q, lnhi[q], lnlo[q], nc_broken) : 0.0;
            }
            if (CANON) {
                // model_spectrum.flux starts at zero and adds NC, Fe, host in turn (Model.py:350-380)
#pragma unroll
                for (int q = 0; q < kPix; ++q) {
                    double v = 0.0;
                    if (has_nc) v = FULL ? nc_v[q] : v + nc_v[q];      // (0.0 + x == x up to the sign of zero)
                    if (n_fe) v = v + fe_v[q];
                    if (n_host) v = v + host_v[q];
                    m[q] = v;
                }
            } else {
                // component values summed in Model.components order (Model.py:353-364):
                // m = everything listed before the Balmer component, post = everything after it
#pragma unroll
                for (int q = 0; q < kPix; ++q) m[q] = 0.0;
                for (int k = 0; k < pl.n_comp; ++k) {
                    if (!((mask >> k) & 1u) || k == q_bal) continue;
                    const int kind = pl.comp_kind[k];
                    if (kind == SPAMM_COMP_EXTINCTION) continue;
#pragma unroll
                    for (int q = 0; q < kPix; ++q) {
                        const double v = kind == SPAMM_COMP_NUCLEAR ? nc_v[q]
                                         : (kind == SPAMM_COMP_FE ? fe_v[q] : host_v[q]);
                        if (k < q_bal) m[q] = m[q] + v; else post[q] = post[q] + v;
                    }
                }
            }
        }
        // warp-uniform: does this task hold any Balmer-continuum / pseudo-continuum pixel?  (a non-finite
        // scale keeps the reference's 0 * scale = NaN on every pixel, BC:445-446)
        const bool task_bc = do_bc && (chunk * kChunk <= bc_istar || !bc_finite);
        const bool task_bpc = do_bpc && is_red && chunk * kChunk + kChunk - 1 > bpc_istar;
        if (task_bpc) {
            if (lorentz) {
                // The line sum is a short list of segments, each either a run of line groups summed
                // directly or one cluster summed by its far-field series (one instance of each body):
                //   exact mode            : [direct 0..n_groups)
                //   a far nested superset : [direct 0..g_lo(s)) [series s]
                //   otherwise             : [direct 0..g_direct_end) then per base cluster series|direct
                int sup = -1, nseg = 1;
                double sup_cmin = 0.0;
                if (series_on) {
                    // nested supersets {n >= n_k}, widest first: the first one that is far from every
                    // pixel of the task absorbs all its lines
                    // (lane 0 holds the task's first pixel, the nearest one when the task is redward of c_bar)
                    const double lam_first = __shfl_sync(0xffffffffu, lam[0], 0);
#pragma unroll
                    for (int sidx = kFullSuper - 1; sidx >= 0; --sidx) {     // widest far superset wins
                        if (sidx < n_super && lam_first >= c.cw[n_base + sidx].far_lam) sup = n_base + sidx;
                    }
                    if (sup >= 0) {
                        const double u = lam_first - c.cw[sup].cb;
                        sup_cmin = fma(u, u, c.cw[sup].kcb2);
                    }
                    nseg = sup >= 0 ? 2 : 1 + n_base;
                }
                if (FULL && sup >= 0) {
                    // the common case, straight-line: the 4/8/16 lines before the far superset, then its series
                    lorentz_direct(sL2, 0, pl.cl[sup].g_lo, lam, acc);
                    cluster_series(c.cw[sup], series_terms(c.cw[sup].As, sup_cmin), lam, acc);
                    nseg = 0;
                }
#pragma unroll 1
                for (int sg = 0; sg < nseg; ++sg) {
                    int g0 = 0, g1 = n_groups, kc = 0;
                    double cmin = 0.0;
                    bool series = false;
                    if (series_on) {
                        if (sg == 0) {
                            g1 = sup >= 0 ? pl.cl[sup].g_lo : pl.g_direct_end;
                        } else if (sup >= 0) {
                            kc = sup; cmin = sup_cmin; series = true;
                        } else {
                            kc = sg - 1;
                            cmin = cluster_cmin(c.cw[kc], lam);
                            series = cmin >= c.cw[kc].thr;
                            g0 = pl.cl[kc].g_lo; g1 = pl.cl[kc].g_hi;
                        }
                    }
                    if (series) cluster_series(c.cw[kc], series_terms(c.cw[kc].As, cmin), lam, acc);
                    else lorentz_direct(sL2, g0, g1, lam, acc);
                }
            } else {
                // Gaussian profile exp(-(d / w)^2), w_n = width * c_n (BC:316-319): beyond 27.4 widths the
                // argument is below -750 and exp() is exactly 0.0 in the reference too, so only the lines
                // with a centre in (lam_lo / (1 + 27.4 width), lam_hi / (1 - 27.4 width)) are visited --
                // bit-identical to the full sum.  Centres fall with the line index.
                int n_first = 0, n_last = n_lines;
                {
                    const double kw = 27.4 * (b_lwid / pl.c_kms);
                    const double lam_lo = pl.wl[chunk * kChunk];                         // live: chunk < n_chunks
                    const double lam_hi = pl.wl[min(chunk * kChunk + kChunk, P) - 1];      // last live pixel
                    const double c_min = lam_lo / (1.0 + kw) * (1.0 - 1e-12);
                    const double c_max = kw < 1.0 ? lam_hi / (1.0 - kw) * (1.0 + 1e-12) : CUDART_INF;
                    if (kw >= 0.0 && isfinite(lam_lo) && isfinite(lam_hi)) {
                        int lo = 0, hi = n_lines;                 // first index with centre <= c_max
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (sLine[(mid >> 2) * 12 + (mid & 3)] > c_max) lo = mid + 1; else hi = mid;
                        }
                        n_first = lo;
                        hi = n_lines;                             // first index with centre < c_min
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (sLine[(mid >> 2) * 12 + (mid & 3)] >= c_min) lo = mid + 1; else hi = mid;
                        }
                        n_last = lo;
                    }
                }
                for (int n = n_first; n < n_last; ++n) {
                    int gq = n >> 2, r4 = n & 3;
                    double lc = sLine[gq * 12 + r4], w2 = sLine[gq * 12 + 4 + r4];
                    double rr = sLine[gq * 12 + 8 + r4];
#pragma unroll
                    for (int q = 0; q < kPix; ++q) {
                        double d = lam[q] - lc;
                        acc[q] = fma(rr, exp_fast(-(d * d) / w2), acc[q]);                       // BC:319
                    }
                }
            }
        }
        double chi = 0.0;
#pragma unroll
        for (int q = 0; q < kPix; ++q) {
            const int i = i0 + QS * q;
            double v = m[q];
            if (have_balmer) {
                double bc = 0.0, bpc = 0.0;
                if (task_bc) bc = ((i <= bc_istar) ? bc_conv_at<FULL>(pl, sF0, i) : 0.0) * bc_scale; // BC:445-446
                if (do_bpc) bpc = ((i > bpc_istar) ? acc[q] : 0.0) * bpc_scale;                // BC:386-387
                v = v + ((do_bc && do_bpc) ? bc + bpc : (do_bc ? bc : bpc));                   // BC:462-471
            }
            if (!CANON && q_bal + 1 < pl.n_comp) v = v + post[q];
            if (has_ext) {
                // pow(10, -0.4 * E(B-V) * k) (ReddeningLaw.py:19,45,71,97,122) as exp(y ln 10) with the
                // product carried in double-double: ~1 ulp, a fifth of the cost of pow()
                const double y = (-0.4 * c.p[pl.ext_off]) * pl.ext_k[i];
                const double th = y * 2.302585092994046;
                const double tl = fma(y, 2.302585092994046, -th) + y * (-2.1707562233822494e-16);
                double e = exp_fast<!FULL>(th);
                e = fma(e, tl, e);
                v = (ext_alone ? 1.0 : v) * e;
            }
            const bool live = i < P;
            if (MODE == 0) {
                double r = (dat[q] - v) * ierr[q];                                              // Model.py:440
                if (live) chi += r * r;
            } else if (live) {
                fout[i] = v;
            }
        }
        if (MODE == 0) {
            for (int o = 16; o > 0; o >>= 1) chi += __shfl_xor_sync(0xffffffffu, chi, o);
            if (lane == 0) sChi[chunk] = chi;
        }
    }

#ifdef SPAMM_TRACE
    if (lane == 0 && g_trace != nullptr) {            // last warp to leave the task loop
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(&g_trace[(size_t)blockIdx.x * 8 + 4], t);
        if (warp == 0) g_trace[(size_t)blockIdx.x * 8 + 5] = t;
    }
#endif
    // Sharded call: its consumer (k_gather / k_stretch_accept, launched with programmatic stream serialisation)
    // waits on the ranks' flags, not on this grid, so it may be made resident as soon as every CTA is past its
    // tasks -- the launch latency of the consumer then overlaps this kernel's tail instead of following it.
    if (MODE == 0 && po.bufs != nullptr) cudaTriggerProgrammaticLaunchCompletion();
    // ---- per-walker reduction in fixed chunk order + likelihood (Model.py:440-445) ----------------
    if (MODE == 0 && PARTS && (only_red || only_blue)) {
        __syncthreads();
        double* gchi = pw.chi + (size_t)wloc * n_chunks;
        for (int k = chunk_bot + tid; k < chunk_top; k += kThreads) gchi[k] = sChi[k];
        __threadfence();
        __syncthreads();
        if (tid == 0) c.prior_ok = (int)atomicAdd(pw.cnt + wloc, 1u);     // (slot reused: 1 = the other part is done)
        __syncthreads();
        if (c.prior_ok == 1 && warp == 0) {
            __threadfence();
            double s = 0.0;
            for (int k = lane; k < n_chunks; k += 32) s += __ldcg(gchi + k);
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) {
                pw.cnt[wloc] = 0u;
                double ln_l = -0.5 * (s + pl.sum_ln);
                store_result(out, w, nan_to_num(ln_l) + 0.0, po);
            }
        }
    } else if (MODE == 0) {
        __syncthreads();
        if (SPLIT) cooperative_groups::this_cluster().sync();          // every CTA's partials are in place
        if (warp == 0 && crank == 0) {
            double s = 0.0;
            for (int k = lane; k < n_chunks; k += 32) {
                if (SPLIT) {
                    // chunk k was task t = n_chunks-1-k, handled by cluster rank t % csize (see the task loop)
                    const double* peer = cooperative_groups::this_cluster().map_shared_rank(
                        sChi, (unsigned)((n_chunks - 1 - k) % csize));
                    s += peer[k];
                } else {
                    s += sChi[k];
                }
            }
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) {
                double ln_l = -0.5 * (s + pl.sum_ln);
                store_result(out, w, nan_to_num(ln_l) + 0.0, po);      // + ln_prior (== 0 inside the box)
            }
        }
        if (SPLIT) cooperative_groups::this_cluster().sync();          // peers' shared memory stays alive
    }
}

// Model.likelihood (Model.py:418-445) on a materialised [W][P] model flux
__global__ void __launch_bounds__(kThreads)
k_likelihood(const PlanDev pl, const double* __restrict__ model, double* __restrict__ out) {
    __shared__ double red[kThreads / 32];
    const int w = blockIdx.x, tid = threadIdx.x;
    const double* m = model + (size_t)w * pl.P;
    double chi = 0.0;
    for (int i = tid; i < pl.P; i += kThreads) {
        double r = (pl.flux[i] - m[i]) / pl.err[i];
        chi += r * r;
    }
    for (int o = 16; o > 0; o >>= 1) chi += __shfl_xor_sync(0xffffffffu, chi, o);
    if ((tid & 31) == 0) red[tid >> 5] = chi;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int q = 0; q < kThreads / 32; ++q) s += red[q];
        out[w] = nan_to_num(-0.5 * (s + pl.sum_ln));
    }
}
