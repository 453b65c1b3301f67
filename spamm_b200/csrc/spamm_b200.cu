// spamm_b200 -- hand-written sm_100a kernels for the SPAMM ln-posterior hot path
// and the C ABI declared in include/spamm_b200.h.
//
// Compiled with -fmad=false: every multiply/add is rounded separately, exactly
// as written (the reference is numpy on x86-64 without FMA contraction; the
// integer decisions -- sigma_norm = ceil(..), argmin edge pixel, interp bins --
// must come out bit-identical).  FMAs are used only where written explicitly
// as fma() (hot inner loops where a 1-ulp difference is harmless).
//
// Reference call sites are cited as file:line relative to the reference root.
//   Model.py            = spamm/Model.py
//   NC / Fe / HG / BC   = spamm/components/{NuclearContinuumComponent,FeComponent,
//                          HostGalaxyComponent,BalmerContinuumCombined}.py
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "spamm_b200.h"

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int set_error(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            return set_error(SPAMM_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
        }                                                                                  \
    } while (0)

// ---------------------------------------------------------------------------
// device-side plan
// ---------------------------------------------------------------------------
#ifndef SPAMM_THREADS
#define SPAMM_THREADS 256
#endif
constexpr int kThreads = SPAMM_THREADS;  // threads per CTA of the fused kernel
#ifndef SPAMM_PIX
#define SPAMM_PIX 2
#endif
constexpr int kPix = SPAMM_PIX;        // pixels per lane in the line loop
constexpr int kMaxDim = 64;            // max parameters per walker
constexpr int kMaxLines = 1024;
constexpr int kConvTile = 256;

enum StatusBits : int {
    kStatSigmaRange = 1,   // sigma_norm outside the bank / < 1
    kStatBcKernel = 2,     // log_conv kernel wider than one pixel (unsupported)
    kStatBcStage = 4,      // BC edge pixel beyond the staged f0 range
};

constexpr int kMaxClusters = 8;        // base clusters + nested supersets
constexpr int kMom = 16;               // series terms kept (validated: <= 1e-15 at rho <= 0.1)
constexpr int kMomPoly = 12;           // terms of the moment power series in z = S_ref / kT (|z| <= 0.2)
struct ClusterDev {
    int g_lo, g_hi;                    // line groups (of 4) covered
    double cb0, R0;                    // centre and half-span at zero velocity offset
};

struct TemplDev {
    int n_templ;
    int n_pts[SPAMM_MAX_TEMPLATES];
    int off[SPAMM_MAX_TEMPLATES];        // offsets into logx/logf
    double bin[SPAMM_MAX_TEMPLATES];     // x[2]-x[1]  (Fe:288)
    double norm_wl[SPAMM_MAX_TEMPLATES];
    int bank_row0[SPAMM_MAX_TEMPLATES];
    int bank_sig_max[SPAMM_MAX_TEMPLATES];
    double templ_width, sigma_denom, sqrt_2pi;
    int ksize, clamp;
    int param_off;                        // first parameter of this component
    const double* logx;
    const double* logf;
    const double* bank_f;                 // [rows][P]
    const double* bank_nf;                // [rows]
};

struct PlanDev {
    int P, D;
    const double *wl, *lnwl, *flux, *err;
    double sum_ln;
    const double *lo, *hi;                // may be null
    // components: bit index -> kind ; parameter offsets
    int n_comp;
    int comp_kind[SPAMM_MAX_COMPONENTS];
    int comp_off[SPAMM_MAX_COMPONENTS];
    // NC
    int nc_broken;
    const double *nc_lnx_hi, *nc_lnx_lo;  // ln(wl / norm_wl) as a double-double
    const double* inv_err;                // 1 / flux_error
    int n_chunks;                         // ceil(P / 64): warp tasks of the fused kernel
    // templates
    TemplDev fe, host;
    // Balmer
    int bc_on, bpc_on, line_type, line_min, n_lines, n_lines_pad;
    const double *line_c0, *line_e;
    const double* line_cum;               // n > 50: sum_{m=51..n} line_e[m] (the ratio recursion telescoped)
    double sh_ne[SPAMM_SH95_NE], sh_T[SPAMM_SH95_T];
    const double* sh_z;
    // far-field clusters of the line forest (see DESIGN.md "line-sum"): contiguous groups of
    // lines whose sum is evaluated as a truncated Taylor series in (c_n - c_bar)/R when the
    // pixel is >= 10 R away in the complex plane; cl[n_clusters] is the union ("super") cluster
    int series_on, n_clusters, n_super, g_direct_end;
    ClusterDev cl[kMaxClusters];          // [0,n_clusters) base, then n_super nested supersets, widest first
    const double* line_d[kMaxClusters];   // per cluster: scaled offsets (c_n - c_bar)/R of its lines
    // n > 50 part of the cluster moments as a power series: sum_n c_n d_n^k exp(z u_n), u_n = S_n / S_ref,
    // = sum_p z^p mom_G[cluster][k][p]; valid (<= 1e-17) while |S_ref| / (k_B Te_min) <= 0.2
    const double* mom_G[kMaxClusters];
    double mom_Sref;
    int mom_poly;
    const double *bc_hnu, *bc_K, *bc_t3;
    const int4* bc_idx;                   // 4 source pixels of the log_conv stencil
    const double *bc_tb0, *bc_tb1, *bc_ta;
    int edge_guess;                       // insertion point of balmer_edge in wl
    int bc_nstage;
    double bc_dlnew;
    double c_kms, c_cgs, c_aa, c_cgs2, kb, edge;
    // Extinction (ReddeningLaw.py): k(lambda) of the selected curve, E(B-V) parameter offset; null = absent
    const double* ext_k;
    int ext_off, ext_q;
    int fe_rows, host_rows;               // rows in the sigma banks
    int exp_safe;                         // the prior box bounds every exp() argument of the fused kernel
    int nc_pair_ok;                       // P even and |slope| |ln(wl[i+1]/wl[i])| <= 0.016 inside the prior box
    int* status;
};

// ---------------------------------------------------------------------------
// numpy-compatible helpers
// ---------------------------------------------------------------------------
// numpy binary_search_with_guess semantics (numpy/_core/src/multiarray/compiled_base.c):
//  -1: x < xp[0]; n: x > xp[n-1]; n-1: x == xp[n-1]; else j with xp[j] <= x < xp[j+1]
__device__ __forceinline__ int np_bin(const double* __restrict__ xp, int n, double x) {
    if (x < xp[0]) return -1;
    if (x > xp[n - 1]) return n;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    if (xp[hi] <= x) return hi;
    return lo;
}

// general branch of numpy arr_interp: slope*(x-x0)+f0 with the NaN rescue
__device__ __forceinline__ double np_lerp(double x, double x0, double x1, double f0, double f1) {
    double slope = (f1 - f0) / (x1 - x0);
    double r = slope * (x - x0) + f0;
    if (isnan(r)) {
        r = slope * (x - x1) + f1;
        if (isnan(r) && f0 == f1) r = f0;
    }
    return r;
}

__device__ __forceinline__ double np_interp_at(double x, const double* __restrict__ xp,
                                               const double* __restrict__ fp, int n, double left,
                                               double right) {
    if (isnan(x)) return x;
    int j = np_bin(xp, n, x);
    if (j < 0) return left;
    if (j >= n) return right;
    if (j == n - 1) return fp[j];
    if (xp[j] == x) return fp[j];
    return np_lerp(x, xp[j], xp[j + 1], fp[j], fp[j + 1]);
}

// np.nan_to_num for a scalar
__device__ __forceinline__ double nan_to_num(double v) {
    if (isnan(v)) return 0.0;
    if (isinf(v)) return v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return v;
}

// utils/find_nearest_index.py:33  (np.abs(arr - value)).argmin(), first minimum.
// wl is ascending, so |wl - v| is V-shaped: candidates are the two neighbours
// of the insertion point; equal-distance runs resolve to the lowest index.
__device__ int nearest_index(const double* __restrict__ wl, int n, double v) {
    if (isnan(v)) return 0;   // argmin of all-NaN is 0
    int lo = 0, hi = n;       // first index with wl[i] >= v
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (wl[mid] < v) lo = mid + 1; else hi = mid;
    }
    int c = lo;
    if (c >= n) c = n - 1;
    if (c > 0 && fabs(wl[c - 1] - v) <= fabs(wl[c] - v)) c = c - 1;
    while (c > 0 && fabs(wl[c - 1] - v) == fabs(wl[c] - v)) --c;
    return c;
}

// exp(x) for |x| < 700: k = rint(x log2 e) by the magic-number trick, r = x - k ln2 in two FMA
// steps, degree-11 minimax polynomial (coefficients in constant memory so each DFMA reads its
// coefficient as a constant-bank operand instead of materialising a 64-bit immediate), scaling by
// 2^k through the exponent field.  ~25 instructions; agrees with exp() to 1 ulp.  Outside the
// range (never on the ln-posterior path) the library exp() keeps its overflow/underflow semantics.
__constant__ double kExpPoly[10] = {
    2.502232253650299e-08, 2.763090348817311e-07, 2.755751454588244e-06, 2.4801491039099165e-05,
    0.00019841269589115497, 0.001388888894591638, 0.008333333333455043, 0.041666666666519754,
    0.16666666666666477, 0.5000000000000012};

template <bool CHECK = true>
__device__ __forceinline__ double exp_fast(double x) {
    if (CHECK && !(fabs(x) < 700.0)) return exp(x);
    const double magic = 6755399441055744.0;
    double t = fma(x, 1.4426950408889634, magic);
    int ki = __double2loint(t);
    double k = t - magic;
    double r = fma(k, -0.6931471805599453, x);
    r = fma(k, -2.3190468138462996e-17, r);
    double p = kExpPoly[0];
#pragma unroll
    for (int j = 1; j < 10; ++j) p = fma(p, r, kExpPoly[j]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (ki << 20), __double2loint(p));
}

// reciprocal to <= 1 ulp: MUFU.RCP64H seed (measured 2^-20 on B200, tools/rcp_probe.cu)
// + one cubic Newton step y(1 + e + e^2), e = 1 - x y  (3 DFMA), no special-case slow path
// (arguments here are products of (dl^2 + w^2) terms, always normal and positive)
__device__ __forceinline__ double rcp_fast(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

// ---------------------------------------------------------------------------
// setup kernels (plan creation)
// ---------------------------------------------------------------------------
// Gaussian broadening in ln-lambda space for a list of rows (template t, sigma_norm):
//   kernel = signal.gaussian(K*sigma, sigma)/(sqrt(2 pi) sigma)          Fe:291-293 / HG:298-300
//   conv   = np.convolve(log_flux, kernel, mode="same")                  Fe:304     / HG:311
// One CTA = one row x one tile of kConvTile output points; the template tile
// (+ halo) and the kernel taps are staged in shared memory.
__global__ void __launch_bounds__(kConvTile)
k_conv_rows(TemplDev ts, const int* __restrict__ row_t, const int* __restrict__ row_sig,
            double* __restrict__ conv, int stride) {
    extern __shared__ double smem[];
    const int r = blockIdx.y;
    const int t = row_t[r];
    const int sig = row_sig[r];
    const int N = ts.n_pts[t];
    const int m0 = blockIdx.x * kConvTile;
    if (m0 >= N) return;
    double* out = conv + (size_t)r * stride;
    if (sig < 1) {   // reference: empty kernel -> ValueError; we poison the row
        int m = m0 + threadIdx.x;
        if (m < N) out[m] = CUDART_NAN;
        return;
    }
    const int M = ts.ksize * sig;
    const int off = (M - 1) / 2;
    double* kw = smem;            // [M]
    double* gs = smem + M;        // [kConvTile + M - 1]
    const double* g = ts.logf + ts.off[t];
    const double sg = (double)sig;
    const double sig2 = 2 * sg * sg;
    const double nrm = ts.sqrt_2pi * sg;
    const double half = ((double)M - 1.0) / 2.0;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        double n = (double)j - half;
        kw[j] = exp(-(n * n) / sig2) / nrm;
    }
    // output m needs g[m + off - j], j = 0..M-1  ->  indices [m0+off-(M-1), m0+kConvTile-1+off]
    const int lo = m0 + off - (M - 1);
    const int span = kConvTile + M - 1;
    for (int s = threadIdx.x; s < span; s += blockDim.x) {
        int idx = lo + s;
        gs[s] = (idx >= 0 && idx < N) ? g[idx] : 0.0;
    }
    __syncthreads();
    const int m = m0 + threadIdx.x;
    if (m >= N) return;
    // gs index of g[m + off - j] = (m + off - j) - lo = threadIdx.x + (M-1) - j
    double acc = 0.0;
    const double* gp = gs + threadIdx.x + (M - 1);
    for (int j = 0; j < M; ++j) acc = fma(gp[-j], kw[j], acc);
    out[m] = acc;
}

// conv (ln-lambda grid) -> data grid:  np.interp(np.log(wl), log_wl, conv [,left=0,right=0])
__global__ void k_interp_rows(TemplDev ts, const int* __restrict__ row_t,
                              const double* __restrict__ conv, int stride,
                              const double* __restrict__ lnwl, int P, double* __restrict__ frows) {
    const int r = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int t = row_t[r];
    const int N = ts.n_pts[t];
    const double* xp = ts.logx + ts.off[t];
    const double* fp = conv + (size_t)r * stride;
    double left = ts.clamp ? fp[0] : 0.0;
    double right = ts.clamp ? fp[N - 1] : 0.0;
    frows[(size_t)r * P + i] = np_interp_at(lnwl[i], xp, fp, N, left, right);
}

// nf = np.nan_to_num(np.interp(median(template wl), wl, f))     Fe:325-331 / HG:330-336
__global__ void k_norm_rows(TemplDev ts, const int* __restrict__ row_t, int n_rows,
                            const double* __restrict__ wl, int P,
                            const double* __restrict__ frows, double* __restrict__ nf) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const double* f = frows + (size_t)r * P;
    double v = np_interp_at(ts.norm_wl[row_t[r]], wl, f, P, f[0], f[P - 1]);
    nf[r] = nan_to_num(v);
}

// sigma_norm = np.ceil(sqrt(width^2 - templ_width^2)/(c_kms 2 sqrt(2 ln 2)) / bin)   Fe:285-290
__device__ __forceinline__ double sigma_norm_of(const TemplDev& ts, int t, double width) {
    double sc = sqrt(width * width - ts.templ_width * ts.templ_width) / ts.sigma_denom;
    return ceil(sc / ts.bin[t]);
}

// rows for the direct mode: one row per (walker, template)
__global__ void k_rows_from_params(TemplDev ts, const double* __restrict__ params, int D, int w0,
                                   int nW, int* __restrict__ row_t, int* __restrict__ row_sig) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nW * ts.n_templ) return;
    int w = idx / ts.n_templ, t = idx % ts.n_templ;
    double width = params[(size_t)(w0 + w) * D + ts.param_off + ts.n_templ];
    double sn = sigma_norm_of(ts, t, width);
    int sig = (sn >= 1.0 && sn <= 1.0e6) ? (int)sn : 0;
    row_t[idx] = t;
    row_sig[idx] = sig;
}

// flag (device word) whether every walker's sigma_norm is inside the bank
__global__ void k_check_bank(TemplDev ts, const double* __restrict__ params, int D, int nW,
                             int* __restrict__ flag) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nW * ts.n_templ) return;
    int w = idx / ts.n_templ, t = idx % ts.n_templ;
    double width = params[(size_t)w * D + ts.param_off + ts.n_templ];
    double sn = sigma_norm_of(ts, t, width);
    if (!(sn >= 1.0 && sn <= (double)ts.bank_sig_max[t])) atomicOr(flag, 1);
}

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

// Multi-GPU result exchange fused into the kernel: instead of writing lnp[w] locally and
// all-gathering afterwards, the finishing thread stores the value straight into every peer's
// symmetric buffer over NVLink (P2P stores); bufs == nullptr selects the plain local store.
struct PeerOut {
    const unsigned long long* bufs;   // device array of n_peers base pointers (incl. this rank)
    int n_peers;
    long long offset;                 // element offset of this launch's walker 0 in the gathered vector
};

__device__ __forceinline__ void store_result(double* out, int w, double v, const PeerOut& po) {
    if (po.bufs == nullptr) { out[w] = v; return; }
    for (int r = 0; r < po.n_peers; ++r)
        reinterpret_cast<double*>(po.bufs[r])[po.offset + w] = v;
}

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

// log_conv output at data pixel i (BC:225-239 with kernel == [1.0]): the two np.interp
// resamplings collapse to a 4-tap stencil on f0 with walker-independent weights
//   fr_j = f0[k] + (f0[k+1]-f0[k]) tB_j ;  out_i = fr_j + (fr_{j+1}-fr_j) tA_i
__device__ __forceinline__ double bc_conv_at(const PlanDev& pl, const double* __restrict__ f0s, int i) {
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
constexpr int kSpecs[] = {1, 3, 5, 9, 15, 31};    // NC | NC+Fe | NC+host | NC+Balmer | full | full+Extinction
template <int MODE, bool CANON, bool SPLIT, int SPEC>
__global__ void __launch_bounds__(kThreads, SPAMM_MIN_BLOCKS)
k_walkers(const PlanDev pl, const double* __restrict__ params, int w0, double* __restrict__ out,
          uint32_t mask, DirectRows dr, PeerOut po) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WalkerCtx& c = *reinterpret_cast<WalkerCtx*>(smem_raw);
    double* sLine = reinterpret_cast<double*>(smem_raw + sizeof(WalkerCtx));   // [n_pad/4][12] grouped
    double* sG = sLine + 3 * pl.n_lines_pad;                                    // [n_lines_pad]
    double* sW = sG + pl.n_lines_pad;                                           // [n_lines_pad]
    double* sChi = sW + pl.n_lines_pad;                                         // [n_chunks]
    double* sF0 = sChi + pl.n_chunks;                                           // [bc_nstage]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int csize = 1, crank = 0;
    if (SPLIT) {
        cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
        csize = (int)cl.num_blocks();
        crank = (int)cl.block_rank();
    }
    const int wloc = SPLIT ? blockIdx.x / csize : blockIdx.x;   // walker within this launch
    const int w = w0 + wloc;
    const int P = pl.P;
    const double* pw = params + (size_t)w * pl.D;

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
    if (tid == 0) c.prior_ok = 1;
    __syncthreads();
    if (tid < pl.D) {
        double v = pw[tid];
        c.p[tid] = v;
        if (MODE == 0 && pl.lo != nullptr) {
            if (!(pl.lo[tid] < v && v < pl.hi[tid])) c.prior_ok = 0;   // strict; NaN fails
        }
    }
    __syncthreads();
    if (MODE == 0 && !c.prior_ok) {       // ln_posterior returns -inf before any flux (Model.py:67-69)
        if (tid == 0 && crank == 0) store_result(out, w, -CUDART_INF, po);
        return;
    }

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
    const bool do_bc = FULL ? have_balmer : have_balmer && pl.bc_on;
    const bool do_bpc = FULL ? have_balmer : have_balmer && pl.bpc_on;
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
    if (has_fe && warp == 1 && lane < n_fe) {
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
            else { atomicOr(pl.status, kStatSigmaRange); norm = CUDART_NAN; }
            int row = ts.bank_row0[t] + sig - 1;
            f = ts.bank_f + (size_t)row * P; nf = ts.bank_nf[row];
        }
        c.tc[t].f = f;
        c.tc[t].a = norm / nf;                                          // Fe:334
    }
    if (has_host && warp == 2 && lane < n_host) {
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
            else { atomicOr(pl.status, kStatSigmaRange); norm = CUDART_NAN; }
            int row = ts.bank_row0[t] + sig - 1;
            f = ts.bank_f + (size_t)row * P; nf = ts.bank_nf[row];
        }
        c.tc[n_fe + t].f = f;
        c.tc[n_fe + t].a = norm / nf;                                        // HG:339
    }
    if (warp == 3 && lane == 0) {
        c.bc_istar = -1;
        c.kT = kT; c.tauBE = b_tau;
        if (do_bc) {
            double edge_wl = pl.edge * (1 - b_loff / pl.c_cgs);          // BC:432 (cm/s!)
            int is = nearest_index_guess(pl.wl, P, edge_wl, pl.edge_guess);   // BC:443
            c.bc_istar = is;
            // log_conv kernel: kernel_width = round(5*dpix) must be 0          BC:229-233
            double dpix = (b_lwid / pl.c_cgs) / pl.bc_dlnew;
            if (rint(5 * dpix) != 0.0) atomicOr(pl.status, kStatBcKernel);
            const int4 kk = pl.bc_idx[is];
            if (max(kk.y, kk.w) >= pl.bc_nstage) atomicOr(pl.status, kStatBcStage);
        }
    }
    if (warp == 4 && lane == 0) {
        c.bpc_istar = -1;
        if (do_bpc) {
            double edge_wl = pl.edge * (1 - b_loff / pl.c_kms);          // BC:374
            c.bpc_istar = nearest_index_guess(pl.wl, P, edge_wl, pl.edge_guess);   // BC:382
        }
    }
    if (do_bpc) {
        const double shift = b_loff / pl.c_kms;          // BC:379
        const double width = b_lwid / pl.c_kms;
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
                else rg = exp(pl.line_cum[n] / kT);                          // BC:296, telescoped
            }
            sLine[(n >> 2) * 12 + (n & 3)] = lc;
            sW[n] = wv;
            sG[n] = rg;      // ratio (n <= 50) or r_n / r_50 (n > 50)
        }
    }
    // Balmer-continuum samples f0 for the staged pixels; loads batched four deep so one L2 round
    // trip is exposed per batch instead of per pixel
    if (do_bc) {
        const int nst = min(pl.bc_nstage, P);
        const double inv_kT = 1.0 / kT;
        for (int k0 = tid; k0 < nst; k0 += 4 * kThreads) {
            double h[4], kk[4], t3[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = min(k0 + j * kThreads, nst - 1);
                h[j] = pl.bc_hnu[k]; kk[j] = pl.bc_K[k]; t3[j] = pl.bc_t3[k];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + j * kThreads;
                if (k < nst) sF0[k] = bc_f0<FULL>(h[j], kk[j], t3[j], kT, inv_kT, b_tau);
            }
        }
    }
    __syncthreads();

    // (the ratio recursion flux_ratios[i] = flux_ratios[i-1] * exp(e_i / kT), BC:292-296, is telescoped:
    //  r_n = r_50 * exp(sum_{m=51..n} e_m / kT); line_cum holds the partial sums, P3 applies r_50)
    // ---- P3: finish the line table; BC and BpC normalisers; cluster moments -------------------------
    const int bc_istar = c.bc_istar, bpc_istar = c.bpc_istar;
    double bpc_scale = 0.0;
    if (do_bc && warp == 7 && lane == 0) {
        double fnorm = bc_conv_at(pl, sF0, bc_istar);                    // BC:444
        c.bc_scale = b_norm / fnorm;                                     // BC:446
        c.bc_finite = isfinite(c.bc_scale) ? 1 : 0;
    }
    if (do_bpc) {
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
        if (warp < n_cl) {
            const ClusterDev& cd = pl.cl[warp];
            const double* dd = pl.line_d[warp];
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
                const double* G = pl.mom_G[warp] + (lane >> 1) * kMomPoly;
                double g[kMomPoly];
#pragma unroll
                for (int q = 0; q < kMomPoly; ++q) g[q] = G[q];
                const double z = pl.mom_Sref / kT;
                double acc = g[kMomPoly - 1];
#pragma unroll
                for (int q = kMomPoly - 2; q >= 0; --q) acc = fma(acc, z, g[q]);
                const double shift = b_loff / pl.c_kms, width = b_lwid / pl.c_kms;
                mk[0] = fma(r50 * (width * (1.0 - shift)), acc, mk[0]);
            }
            if (!(lane & 1)) c.cw[warp].mom[lane >> 1] = mk[0];
            if (lane == 0) {
                const double shift = b_loff / pl.c_kms, width = b_lwid / pl.c_kms;
                const double kappa = width * width;
                const double cb = cd.cb0 - shift * cd.cb0, R = cd.R0 - shift * cd.R0;
                ClusterW& cw = c.cw[warp];
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
    if (do_bpc) {
        double tot = 0.0;
        for (int q = 0; q < kThreads / 32; ++q) tot += c.red[q];
        bpc_scale = b_norm / tot;                                        // BC:387
    }
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
    for (int t = SPLIT ? warp * csize + crank : warp; t < n_chunks; t += (kThreads / 32) * csize) {
        const int chunk = n_chunks - 1 - t;
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
        const bool task_bpc = do_bpc && chunk * kChunk + kChunk - 1 > bpc_istar;
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
                for (int n = 0; n < n_lines; ++n) {
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
                if (task_bc) bc = ((i <= bc_istar) ? bc_conv_at(pl, sF0, i) : 0.0) * bc_scale; // BC:445-446
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

    // ---- per-walker reduction in fixed chunk order + likelihood (Model.py:440-445) ----------------
    if (MODE == 0) {
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

// ---------------------------------------------------------------------------
// sampler kernels (emcee 2.2.1 stretch move)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                           uint32_t c2, uint32_t c3, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in [0,1) with 53 bits
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    uint64_t v = ((uint64_t)a << 32) | b;
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void k_stretch_propose(const double* __restrict__ walkers, int K, int D, int half,
                                  double a, uint64_t seed, uint64_t iter, double* __restrict__ q,
                                  double* __restrict__ z) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int Ns = K / 2;
    if (i >= Ns) return;
    uint32_t r[4];
    philox4x32((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)i, (uint32_t)half,
               (uint32_t)iter, (uint32_t)(iter >> 32) ^ 0x50524f50u, r);
    double u = u53(r[0], r[1]);
    double zz = ((a - 1.0) * u + 1.0);
    zz = zz * zz / a;
    int Nc = K - Ns;
    int rint_ = (int)(u53(r[2], r[3]) * Nc);
    if (rint_ >= Nc) rint_ = Nc - 1;
    const double* s = walkers + (size_t)(half == 0 ? i : Ns + i) * D;
    const double* cc = walkers + (size_t)(half == 0 ? Ns + rint_ : rint_) * D;
    for (int d = 0; d < D; ++d) q[(size_t)i * D + d] = cc[d] - zz * (cc[d] - s[d]);
    z[i] = zz;
}

__global__ void k_stretch_accept(double* __restrict__ walkers, double* __restrict__ lnp, int K,
                                 int D, int half, const double* __restrict__ q,
                                 const double* __restrict__ z, const double* __restrict__ lnp_q,
                                 uint64_t seed, uint64_t iter, long long* __restrict__ nacc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int Ns = K / 2;
    if (i >= Ns) return;
    uint32_t r[4];
    philox4x32((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)i, (uint32_t)half,
               (uint32_t)iter, (uint32_t)(iter >> 32) ^ 0x41434350u, r);
    double u = u53(r[0], r[1]);
    int wi = half == 0 ? i : Ns + i;
    double lnpdiff = ((double)D - 1.0) * log(z[i]) + lnp_q[i] - lnp[wi];
    if (lnpdiff > log(u)) {
        for (int d = 0; d < D; ++d) walkers[(size_t)wi * D + d] = q[(size_t)i * D + d];
        lnp[wi] = lnp_q[i];
        if (nacc) nacc[wi] += 1;
    }
}

__global__ void k_chain_record(const double* __restrict__ walkers, const double* __restrict__ lnp,
                               int K, int D, long long iter, long long n_iter,
                               double* __restrict__ chain, double* __restrict__ lnp_chain) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    if (chain)
        for (int d = 0; d < D; ++d) chain[((size_t)i * n_iter + iter) * D + d] = walkers[(size_t)i * D + d];
    if (lnp_chain) lnp_chain[(size_t)i * n_iter + iter] = lnp[i];
}

// ---------------------------------------------------------------------------
// FP64 FMA peak probe
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_probe(int iters, double* sink) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-9, a2 = a0 + 2e-9, a3 = a0 + 3e-9;
    double a4 = a0 + 4e-9, a5 = a0 + 5e-9, a6 = a0 + 6e-9, a7 = a0 + 7e-9;
    const double m = 0.999999999, b = 1e-12;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

// ---------------------------------------------------------------------------
// host side: the plan
// ---------------------------------------------------------------------------
struct SpammPlan {
    PlanDev dev{};
    std::vector<void*> allocs;
    int template_mode = SPAMM_TEMPLATES_BANK;
    int bank_rows = 0;
    int max_walkers = 0;
    int n_sm = 148;
    size_t smem_bytes = 0;
    int64_t launches = 0;
    bool has_data = false;
    // direct-mode workspace
    int direct_chunk = 0;
    double *d_conv = nullptr, *d_fe_f = nullptr, *d_fe_nf = nullptr, *d_host_f = nullptr, *d_host_nf = nullptr;
    int *d_row_t = nullptr, *d_row_sig = nullptr;
    int conv_stride = 0;
    int* d_flag = nullptr;
    // host staging for the *_host entry point
    double *h_params = nullptr, *h_lnp = nullptr, *d_params = nullptr, *d_lnp = nullptr;
    cudaStream_t own_stream = nullptr;
    int max_sigma_seen = 0;
    int last_bank_rows = 0;
    int split_mode = 1;          // 0: never split a walker over a cluster; 1: auto; 2/4/8: forced size
    bool use_full_kernel = getenv("SPAMM_NO_FULL_KERNEL") == nullptr;   // tests compare it with the generic one
};

// Every device array gets kPad zeroed elements of slack: the fused kernel's last warp task
// reads pixels [P, P + 63] (results discarded) instead of clamping indices.
constexpr size_t kPad = 64;

template <typename T>
static int upload(SpammPlan* pl, const T* host, size_t n, const T** dev_out) {
    T* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (n + kPad) * sizeof(T)));
    pl->allocs.push_back(d);
    CUDA_TRY(cudaMemset(d, 0, (n + kPad) * sizeof(T)));
    if (n) CUDA_TRY(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *dev_out = d;
    return SPAMM_OK;
}

template <typename T>
static int dalloc(SpammPlan* pl, size_t n, T** dev_out) {
    T* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (n + kPad) * sizeof(T)));
    pl->allocs.push_back(d);
    CUDA_TRY(cudaMemset(d, 0, (n + kPad) * sizeof(T)));
    *dev_out = d;
    return SPAMM_OK;
}

static int conv_smem_bytes(int ksize, int max_sigma) {
    int M = ksize * max_sigma;
    return (M + kConvTile + M - 1) * (int)sizeof(double);
}

// run the broadening pipeline for n_rows rows -> frows[n_rows][P], nf[n_rows]
static int run_rows(SpammPlan* pl, const TemplDev& ts, const int* d_row_t, const int* d_row_sig,
                    int n_rows, int max_sigma, double* d_conv, int stride, double* d_frows,
                    double* d_nf, cudaStream_t st) {
    if (n_rows == 0) return SPAMM_OK;
    int smem = conv_smem_bytes(ts.ksize, std::max(max_sigma, 1));
    if (smem > 200 * 1024)
        return set_error(SPAMM_ERANGE, "broadening kernel too wide for shared memory");
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(k_conv_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int nmax = 0;
    for (int t = 0; t < ts.n_templ; ++t) nmax = std::max(nmax, ts.n_pts[t]);
    dim3 g1((nmax + kConvTile - 1) / kConvTile, n_rows);
    k_conv_rows<<<g1, kConvTile, smem, st>>>(ts, d_row_t, d_row_sig, d_conv, stride);
    dim3 g2((pl->dev.P + 255) / 256, n_rows);
    k_interp_rows<<<g2, 256, 0, st>>>(ts, d_row_t, d_conv, stride, pl->dev.lnwl, pl->dev.P, d_frows);
    k_norm_rows<<<(n_rows + 127) / 128, 128, 0, st>>>(ts, d_row_t, n_rows, pl->dev.wl, pl->dev.P,
                                                       d_frows, d_nf);
    pl->launches += 3;
    CUDA_TRY(cudaGetLastError());
    return SPAMM_OK;
}

static double host_sigma_norm(const SpammTemplateSet& s, double bin, double width) {
    volatile double a = width * width;
    volatile double b = s.template_width * s.template_width;
    volatile double sc = std::sqrt(a - b) / s.sigma_denom;
    return std::ceil(sc / bin);
}

static int setup_template_set(SpammPlan* pl, const SpammTemplateSet& src, int param_off,
                              double width_hi, bool want_bank, TemplDev* out, bool* bank_ok) {
    TemplDev ts{};
    ts.n_templ = src.n_templates;
    if (ts.n_templ < 1 || ts.n_templ > SPAMM_MAX_TEMPLATES)
        return set_error(SPAMM_EINVAL, "template count out of range");
    size_t total = 0;
    for (int t = 0; t < ts.n_templ; ++t) {
        if (src.n_points[t] < 3) return set_error(SPAMM_EINVAL, "template needs >= 3 points");
        ts.n_pts[t] = src.n_points[t];
        ts.off[t] = (int)total;
        // bin_size = log grid [2] - [1]                                       Fe:288 / HG:295
        ts.bin[t] = src.log_wl[total + 2] - src.log_wl[total + 1];
        ts.norm_wl[t] = src.norm_wl[t];
        total += src.n_points[t];
    }
    ts.templ_width = src.template_width;
    ts.sigma_denom = src.sigma_denom;
    ts.sqrt_2pi = src.sqrt_2pi;
    ts.ksize = src.kernel_size_sigma;
    ts.clamp = src.clamp_edges;
    ts.param_off = param_off;
    int rc;
    if ((rc = upload(pl, src.log_wl, total, &ts.logx))) return rc;
    if ((rc = upload(pl, src.log_flux, total, &ts.logf))) return rc;
    for (int t = 0; t < ts.n_templ; ++t) pl->conv_stride = std::max(pl->conv_stride, ts.n_pts[t]);

    *bank_ok = false;
    if (want_bank && std::isfinite(width_hi)) {
        int rows = 0;
        bool ok = true;
        for (int t = 0; t < ts.n_templ; ++t) {
            double sn = host_sigma_norm(src, ts.bin[t], width_hi);
            if (!(sn >= 1.0) || sn > 512.0) { ok = false; break; }
            ts.bank_row0[t] = rows;
            ts.bank_sig_max[t] = (int)sn;
            rows += (int)sn;
        }
        if (ok && (size_t)rows * pl->dev.P * sizeof(double) <= ((size_t)2 << 30)) {
            std::vector<int> rt(rows), rs(rows);
            int max_sigma = 0;
            for (int t = 0; t < ts.n_templ; ++t)
                for (int s = 1; s <= ts.bank_sig_max[t]; ++s) {
                    rt[ts.bank_row0[t] + s - 1] = t;
                    rs[ts.bank_row0[t] + s - 1] = s;
                    max_sigma = std::max(max_sigma, s);
                }
            const int *d_rt, *d_rs;
            if ((rc = upload(pl, rt.data(), rows, &d_rt))) return rc;
            if ((rc = upload(pl, rs.data(), rows, &d_rs))) return rc;
            double *d_conv, *d_f, *d_nf;
            int stride = 0;
            for (int t = 0; t < ts.n_templ; ++t) stride = std::max(stride, ts.n_pts[t]);
            CUDA_TRY(cudaMalloc(&d_conv, (size_t)rows * stride * sizeof(double)));
            if ((rc = dalloc(pl, (size_t)rows * pl->dev.P, &d_f))) { cudaFree(d_conv); return rc; }
            if ((rc = dalloc(pl, (size_t)rows, &d_nf))) { cudaFree(d_conv); return rc; }
            rc = run_rows(pl, ts, d_rt, d_rs, rows, max_sigma, d_conv, stride, d_f, d_nf, 0);
            cudaError_t e = cudaDeviceSynchronize();
            cudaFree(d_conv);
            if (rc) return rc;
            if (e != cudaSuccess) return set_error(SPAMM_ECUDA, cudaGetErrorString(e));
            ts.bank_f = d_f;
            ts.bank_nf = d_nf;
            pl->last_bank_rows = rows;
            pl->bank_rows += rows;
            *bank_ok = true;
        }
    }
    *out = ts;
    return SPAMM_OK;
}

extern "C" int spamm_abi_version(void) { return SPAMM_ABI_VERSION; }
extern "C" const char* spamm_last_error(void) { return g_last_error.c_str(); }

extern "C" void spamm_plan_destroy(SpammPlan* pl) {
    if (!pl) return;
    for (void* p : pl->allocs) cudaFree(p);
    cudaFree(pl->d_conv); cudaFree(pl->d_fe_f); cudaFree(pl->d_fe_nf);
    cudaFree(pl->d_host_f); cudaFree(pl->d_host_nf); cudaFree(pl->d_row_t); cudaFree(pl->d_row_sig);
    cudaFree(pl->d_params); cudaFree(pl->d_lnp);
    if (pl->h_params) cudaFreeHost(pl->h_params);
    if (pl->h_lnp) cudaFreeHost(pl->h_lnp);
    if (pl->own_stream) cudaStreamDestroy(pl->own_stream);
    delete pl;
}

static int plan_create_impl(const SpammPlanDesc* d, SpammPlan* pl) {
    if (d->abi_version != SPAMM_ABI_VERSION) return set_error(SPAMM_EINVAL, "ABI version mismatch");
    const int P = d->n_pix;
    if (P < 3 || !d->wl || !d->ln_wl) return set_error(SPAMM_EINVAL, "need n_pix >= 3, wl and ln_wl");
    for (int i = 1; i < P; ++i)
        if (!(d->wl[i] > d->wl[i - 1]))
            return set_error(SPAMM_EINVAL, "wl must be strictly ascending");
    if (d->n_components < 1 || d->n_components > SPAMM_MAX_COMPONENTS)
        return set_error(SPAMM_EINVAL, "n_components out of range");
    if (d->n_dim < 1 || d->n_dim > kMaxDim) return set_error(SPAMM_EINVAL, "n_dim out of range");
    if (d->max_walkers < 1) return set_error(SPAMM_EINVAL, "max_walkers must be >= 1");

    int dev_id = 0;
    CUDA_TRY(cudaGetDevice(&dev_id));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev_id));
    pl->n_sm = prop.multiProcessorCount;
    pl->max_walkers = d->max_walkers;

    PlanDev& pd = pl->dev;
    pd.P = P;
    pd.D = d->n_dim;
    int rc;
    if ((rc = upload(pl, d->wl, P, &pd.wl))) return rc;
    if ((rc = upload(pl, d->ln_wl, P, &pd.lnwl))) return rc;
    pl->has_data = d->flux && d->flux_error;
    if (pl->has_data) {
        if ((rc = upload(pl, d->flux, P, &pd.flux))) return rc;
        if ((rc = upload(pl, d->flux_error, P, &pd.err))) return rc;
        std::vector<double> ie(P);
        for (int i = 0; i < P; ++i) ie[i] = 1.0 / d->flux_error[i];
        if ((rc = upload(pl, ie.data(), P, &pd.inv_err))) return rc;
    }
    pd.n_chunks = (P + kChunk - 1) / kChunk;
    pd.sum_ln = d->sum_ln_2pi_err2;
    if (d->prior_lo && d->prior_hi) {
        if ((rc = upload(pl, d->prior_lo, d->n_dim, &pd.lo))) return rc;
        if ((rc = upload(pl, d->prior_hi, d->n_dim, &pd.hi))) return rc;
    }
    if ((rc = dalloc(pl, 1, &pd.status))) return rc;
    CUDA_TRY(cudaMemset(pd.status, 0, sizeof(int)));
    if ((rc = dalloc(pl, 1, &pl->d_flag))) return rc;

    // component layout
    pd.n_comp = d->n_components;
    int off = 0, seen[5] = {0, 0, 0, 0, 0};
    pd.fe_rows = 0; pd.host_rows = 0;
    pd.ext_k = nullptr; pd.ext_off = 0; pd.ext_q = 0;
    int off_fe = -1, off_host = -1;
    for (int q = 0; q < d->n_components; ++q) {
        int kind = d->component_kind[q];
        if (kind < 0 || kind > 4) return set_error(SPAMM_EINVAL, "unknown component kind");
        if (seen[kind]++) return set_error(SPAMM_EINVAL, "each component kind may appear once");
        pd.comp_kind[q] = kind;
        pd.comp_off[q] = off;
        if (kind == SPAMM_COMP_NUCLEAR) off += d->nc_broken_pl ? 4 : 2;
        else if (kind == SPAMM_COMP_FE) { off_fe = off; off += d->fe.n_templates + 1; }
        else if (kind == SPAMM_COMP_HOST) { off_host = off; off += d->host.n_templates + 1; }
        else if (kind == SPAMM_COMP_EXTINCTION) {
            if (q != d->n_components - 1)
                return set_error(SPAMM_EINVAL, "Extinction must be the last component (run_spamm.py:123-125)");
            if (!d->ext_k) return set_error(SPAMM_EINVAL, "missing extinction curve");
            pd.ext_off = off; pd.ext_q = q;
            off += 1;
        } else off += 6;
    }
    if (seen[SPAMM_COMP_EXTINCTION])
        if ((rc = upload(pl, d->ext_k, P, &pd.ext_k))) return rc;
    if (off != d->n_dim) return set_error(SPAMM_EINVAL, "n_dim does not match the component list");

    // NC: ln(wl / norm_wl) in double-double (PowerLaw1D evaluates (x / x_0) ** (-alpha))
    pd.nc_broken = d->nc_broken_pl;
    if (seen[SPAMM_COMP_NUCLEAR]) {
        std::vector<double> hi(P), lo(P);
        for (int i = 0; i < P; ++i) {
            double x = d->wl[i] / d->nc_norm_wl;
            long double L = logl((long double)x);
            hi[i] = (double)L;
            lo[i] = (double)(L - (long double)hi[i]);
        }
        if ((rc = upload(pl, hi.data(), P, &pd.nc_lnx_hi))) return rc;
        if ((rc = upload(pl, lo.data(), P, &pd.nc_lnx_lo))) return rc;
    }

    // templates
    bool want_bank = d->template_mode != SPAMM_TEMPLATES_DIRECT;
    bool all_bank = true;
    if (seen[SPAMM_COMP_FE]) {
        bool ok = false;
        double hi = d->prior_hi ? d->prior_hi[off_fe + d->fe.n_templates] : INFINITY;
        if ((rc = setup_template_set(pl, d->fe, off_fe, hi, want_bank, &pd.fe, &ok))) return rc;
        pd.fe_rows = ok ? pl->last_bank_rows : 0;
        all_bank = all_bank && ok;
    }
    if (seen[SPAMM_COMP_HOST]) {
        bool ok = false;
        double hi = d->prior_hi ? d->prior_hi[off_host + d->host.n_templates] : INFINITY;
        if ((rc = setup_template_set(pl, d->host, off_host, hi, want_bank, &pd.host, &ok))) return rc;
        pd.host_rows = ok ? pl->last_bank_rows : 0;
        all_bank = all_bank && ok;
    }
    if (d->template_mode == SPAMM_TEMPLATES_BANK && !all_bank)
        return set_error(SPAMM_EINVAL, "template bank requested but the width prior is unbounded/too wide");
    pl->template_mode = (want_bank && all_bank) ? SPAMM_TEMPLATES_BANK : SPAMM_TEMPLATES_DIRECT;

    // Balmer
    pd.bc_on = d->balmer_bc;
    pd.bpc_on = d->balmer_bpc;
    pd.line_type = d->balmer_line_type;
    pd.n_lines = 0;
    pd.n_lines_pad = 0;
    pd.bc_nstage = 0;
    if (seen[SPAMM_COMP_BALMER]) {
        if (!pd.bc_on && !pd.bpc_on)
            return set_error(SPAMM_EINVAL, "BalmerCombined needs BalmerContinuum and/or BalmerPseudocContinuum");
        pd.c_kms = d->c_kms; pd.c_cgs = d->c_cgs; pd.c_aa = d->c_aa;
        pd.c_cgs2 = d->c_cgs * d->c_cgs;                      // c.cgs.value**2
        pd.kb = d->kb_cgs; pd.edge = d->balmer_edge;
        if (pd.bpc_on) {
            if (d->balmer_line_min < 3) return set_error(SPAMM_EINVAL, "bc_lines_min must be >= 3");
            if (d->balmer_n_lines < 1 || d->balmer_n_lines > kMaxLines)
                return set_error(SPAMM_EINVAL, "number of Balmer lines out of range");
            if (d->balmer_line_min <= 50 && !d->sh95_z) return set_error(SPAMM_EINVAL, "missing SH95 table");
            pd.line_min = d->balmer_line_min;
            pd.n_lines = d->balmer_n_lines;
            pd.n_lines_pad = (pd.n_lines + 3) & ~3;
            // far-field clusters over line-table indices; boundaries are multiples of 4 lines.
            // For bc_lines_min = 3: base clusters n = 19..30 | 31..50 | 51..98 | 99..max (lines
            // 3..18 always direct next to them) and nested supersets {n >= 7}, {n >= 11}, {n >= 19}
            // for pixels far from the whole forest.
            pd.series_on = (!d->balmer_exact_line_sum && pd.line_type == SPAMM_LINE_LORENTZIAN) ? 1 : 0;
            pd.n_clusters = 0; pd.n_super = 0; pd.g_direct_end = pd.n_lines_pad / 4;
            {
                const int bounds[5] = {16, 28, 48, 96, pd.n_lines_pad};
                const int super_lo[3] = {4, 8, 16};
                const double* cc = d->balmer_line_center;
                auto make = [&](int lo, int hi, int slot) -> int {
                    int hi_line = std::min(hi, pd.n_lines);            // exclude padding lines
                    double cmax = cc[lo], cmin = cc[lo];
                    for (int n = lo; n < hi_line; ++n) { cmax = std::max(cmax, cc[n]); cmin = std::min(cmin, cc[n]); }
                    ClusterDev* out = &pd.cl[slot];
                    out->g_lo = lo / 4; out->g_hi = hi / 4;
                    out->cb0 = 0.5 * (cmax + cmin);
                    out->R0 = 0.5 * (cmax - cmin);
                    std::vector<double> dd(pd.n_lines_pad, 0.0);
                    for (int n = lo; n < hi_line; ++n) dd[n] = (cc[n] - out->cb0) / out->R0;
                    return upload(pl, dd.data(), dd.size(), &pd.line_d[slot]);
                };
                if (pd.series_on && pd.n_lines >= 32) {
                    pd.g_direct_end = bounds[0] / 4;
                    for (int b = 0; b < 4; ++b) {
                        int lo = bounds[b], hi = std::min(bounds[b + 1], pd.n_lines_pad);
                        if (lo >= pd.n_lines) break;
                        if (std::min(hi, pd.n_lines) - lo < 2) { hi = pd.n_lines_pad; }
                        if ((rc = make(lo, hi, pd.n_clusters))) return rc;
                        pd.n_clusters++;
                        if (hi == pd.n_lines_pad) break;
                    }
                    for (int sidx = 0; sidx < 3; ++sidx) {
                        if (pd.n_lines - super_lo[sidx] < 8) continue;
                        if ((rc = make(super_lo[sidx], pd.n_lines_pad, pd.n_clusters + pd.n_super))) return rc;
                        pd.n_super++;
                    }
                } else {
                    pd.series_on = 0;
                }
            }
            if ((rc = upload(pl, d->balmer_line_center, pd.n_lines, &pd.line_c0))) return rc;
            if ((rc = upload(pl, d->balmer_line_exparg, pd.n_lines, &pd.line_e))) return rc;
            {
                std::vector<double> cum(pd.n_lines, 0.0);
                long double acc = 0.0L;
                for (int n = 0; n < pd.n_lines; ++n) {
                    if (pd.line_min + n > 50) { acc += (long double)d->balmer_line_exparg[n]; cum[n] = (double)acc; }
                }
                if ((rc = upload(pl, cum.data(), pd.n_lines, &pd.line_cum))) return rc;
                // moment power series of the n > 50 lines (see PlanDev::mom_G)
                pd.mom_poly = 0;
                pd.mom_Sref = cum[pd.n_lines - 1];
                const int n50 = std::min(std::max(51 - pd.line_min, 0), pd.n_lines);
                int o_bal = -1;
                for (int q = 0; q < pd.n_comp; ++q) if (pd.comp_kind[q] == SPAMM_COMP_BALMER) o_bal = pd.comp_off[q];
                const double te_lo = (d->prior_lo && o_bal >= 0) ? d->prior_lo[o_bal + 1] : 0.0;
                if (pd.series_on && n50 < pd.n_lines && pd.mom_Sref < 0.0 && te_lo > 0.0 &&
                    std::fabs(pd.mom_Sref) / (d->kb_cgs * te_lo) <= 0.2) {
                    const double* cc = d->balmer_line_center;
                    for (int k = 0; k < pd.n_clusters + pd.n_super; ++k) {
                        std::vector<long double> G((size_t)kMom * kMomPoly, 0.0L);
                        const int lo = std::max(pd.cl[k].g_lo * 4, n50), hi = std::min(pd.cl[k].g_hi * 4, pd.n_lines);
                        for (int n = lo; n < hi; ++n) {
                            const long double dn = ((long double)cc[n] - (long double)pd.cl[k].cb0) / (long double)pd.cl[k].R0;
                            const long double dn_dev = (long double)(double)dn;   // the device uses the rounded offset
                            const long double un = (long double)cum[n] / (long double)pd.mom_Sref;
                            long double dj = 1.0L;
                            for (int j = 0; j < kMom; ++j) {
                                long double up = 1.0L;                              // u^p / p!
                                for (int q = 0; q < kMomPoly; ++q) {
                                    G[(size_t)j * kMomPoly + q] += (long double)cc[n] * dj * up;
                                    up *= un / (long double)(q + 1);
                                }
                                dj *= dn_dev;
                            }
                        }
                        std::vector<double> Gd(G.size());
                        for (size_t i = 0; i < G.size(); ++i) Gd[i] = (double)G[i];
                        if ((rc = upload(pl, Gd.data(), Gd.size(), &pd.mom_G[k]))) return rc;
                    }
                    pd.mom_poly = 1;
                }
            }
            memcpy(pd.sh_ne, d->sh95_ne, sizeof(pd.sh_ne));
            memcpy(pd.sh_T, d->sh95_T, sizeof(pd.sh_T));
            if (d->sh95_z)
                if ((rc = upload(pl, d->sh95_z, (size_t)SPAMM_SH95_LINES * SPAMM_SH95_NE * SPAMM_SH95_T, &pd.sh_z)))
                    return rc;
        }
        if (pd.bc_on) {
            if (!d->bc_ln_wl_new || !d->bc_exp_ln_wl_new)
                return set_error(SPAMM_EINVAL, "missing log_conv grids");
            // walker-independent pieces of blackbody_lambda / tau
            const double c2 = d->c_cgs * d->c_cgs;                        // c.cgs.value ** 2
            std::vector<double> hnu(P), K(P), t3(P);
            for (int i = 0; i < P; ++i) {
                double freq = d->c_aa / d->wl[i];                         // c / lambda
                hnu[i] = d->h_cgs * freq;                                 // h * freq
                double b1 = (2.0 * d->h_cgs) * std::pow(freq, 3.0);       // 2 h freq^3
                double lam2 = d->wl[i] * d->wl[i];
                K[i] = b1 * d->c_aa / (c2 * lam2);                        // B_lambda = K / expm1(h nu / kT)
                t3[i] = std::pow(d->wl[i] / d->balmer_edge, 3.0);         // (wl/balmer_edge)**3
            }
            if ((rc = upload(pl, hnu.data(), P, &pd.bc_hnu))) return rc;
            if ((rc = upload(pl, K.data(), P, &pd.bc_K))) return rc;
            if ((rc = upload(pl, t3.data(), P, &pd.bc_t3))) return rc;
            pd.bc_dlnew = d->bc_ln_wl_new[1] - d->bc_ln_wl_new[0];
            // log_conv (BC:218-241) = np.interp onto the uniform ln-lambda grid, identity kernel,
            // np.interp back: per output pixel a 4-tap stencil with walker-independent weights
            const double* lnnew = d->bc_ln_wl_new;
            const double* expnew = d->bc_exp_ln_wl_new;
            auto bin = [](const double* xp, int n, double x) {
                if (x < xp[0]) return -1;
                if (x > xp[n - 1]) return n;
                int lo = 0, hi = n - 1;
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (xp[mid] <= x) lo = mid; else hi = mid; }
                return xp[hi] <= x ? hi : lo;
            };
            auto tap = [&](const double* xp, double x, int* k, double* t) {
                int j = bin(xp, P, x);
                if (j < 0) { *k = 0; *t = 0.0; }
                else if (j >= P - 1) { *k = P - 1; *t = 0.0; }
                else if (xp[j] == x) { *k = j; *t = 0.0; }
                else { *k = j; *t = (x - xp[j]) / (xp[j + 1] - xp[j]); }
            };
            std::vector<int4> idx(P);
            std::vector<double> tb0(P), tb1(P), ta(P);
            for (int i = 0; i < P; ++i) {
                int j0; double tA;
                tap(expnew, d->wl[i], &j0, &tA);                       // np.interp(wl, exp(ln_wavenew), .)
                int j1 = std::min(j0 + 1, P - 1);
                int k0, k1; double t0, t1;
                tap(d->ln_wl, lnnew[j0], &k0, &t0);                    // np.interp(ln_wavenew, ln_wave, .)
                tap(d->ln_wl, lnnew[j1], &k1, &t1);
                idx[i] = make_int4(k0, std::min(k0 + 1, P - 1), k1, std::min(k1 + 1, P - 1));
                tb0[i] = t0; tb1[i] = t1; ta[i] = tA;
            }
            if ((rc = upload(pl, idx.data(), P, &pd.bc_idx))) return rc;
            if ((rc = upload(pl, tb0.data(), P, &pd.bc_tb0))) return rc;
            if ((rc = upload(pl, tb1.data(), P, &pd.bc_tb1))) return rc;
            if ((rc = upload(pl, ta.data(), P, &pd.bc_ta))) return rc;
            // stage f0 up to the furthest source pixel of (nearest pixel to the edge) + 2
            int ibe = 0;
            double best = INFINITY;
            for (int i = 0; i < P; ++i) {
                double v = std::fabs(d->wl[i] - d->balmer_edge);
                if (v < best) { best = v; ibe = i; }
            }
            int itop = std::min(ibe + 2, P - 1), top = 0;
            for (int i = 0; i <= itop; ++i) top = std::max(top, std::max(idx[i].y, idx[i].w));
            pd.bc_nstage = std::min(P, top + 1);
        }
        // insertion point of the Balmer edge: start of the per-walker edge-pixel search
        pd.edge_guess = (int)(std::lower_bound(d->wl, d->wl + P, d->balmer_edge) - d->wl);
    }

    // exp_safe: inside the prior box every exp() argument of the fused kernel stays below 690 in
    // magnitude, so the FULL instantiation (ln-posterior only: out-of-prior walkers exit first) can
    // drop the range checks.  NC: |slope| max|ln(wl/wl0)|; BC: h nu / (k Te) at the bluest pixel and
    // Te_min; tau_BE (wl/edge)^3 over the staged pixels.
    pd.exp_safe = 0;
    pd.nc_pair_ok = 0;
    if (d->prior_lo && d->prior_hi && seen[SPAMM_COMP_NUCLEAR] && !d->nc_broken_pl) {
        int o_nc = -1, o_bal = -1;
        for (int q = 0; q < pd.n_comp; ++q) {
            if (pd.comp_kind[q] == SPAMM_COMP_NUCLEAR) o_nc = pd.comp_off[q];
            if (pd.comp_kind[q] == SPAMM_COMP_BALMER) o_bal = pd.comp_off[q];
        }
        double lmax = 0.0;
        for (int i = 0; i < P; ++i) lmax = std::max(lmax, std::fabs(std::log(d->wl[i] / d->nc_norm_wl)));
        const double smax = std::max(std::fabs(d->prior_lo[o_nc + 1]), std::fabs(d->prior_hi[o_nc + 1]));
        bool ok = smax * lmax < 690.0;                                  // NaN bounds compare false
        if (o_bal >= 0 && pd.bc_on) {
            const double te_lo = d->prior_lo[o_bal + 1];
            const double taumax = std::max(std::fabs(d->prior_lo[o_bal + 2]), std::fabs(d->prior_hi[o_bal + 2]));
            const double hnu0 = d->h_cgs * (d->c_aa / d->wl[0]);
            const double t3max = std::pow(d->wl[std::max(pd.bc_nstage, 1) - 1] / d->balmer_edge, 3.0);
            ok = ok && te_lo > 0.0 && hnu0 / (d->kb_cgs * te_lo) < 690.0 && taumax * t3max < 690.0;
        }
        if (seen[SPAMM_COMP_EXTINCTION]) {
            double kmax = 0.0;
            for (int i = 0; i < P; ++i) kmax = std::max(kmax, std::fabs(d->ext_k[i]));
            const double emax = std::max(std::fabs(d->prior_lo[pd.ext_off]), std::fabs(d->prior_hi[pd.ext_off]));
            ok = ok && 0.4 * emax * kmax * 2.302585092994046 < 690.0;
        }
        pd.exp_safe = ok ? 1 : 0;
        double dmax = 0.0;
        for (int i = 0; i + 1 < P; ++i) dmax = std::max(dmax, std::fabs(std::log(d->wl[i + 1] / d->wl[i])));
        pd.nc_pair_ok = (P % 2 == 0 && smax * dmax <= 0.016) ? 1 : 0;
    }

    // shared memory of the fused kernel
    pl->smem_bytes = sizeof(WalkerCtx) + (size_t)(5 * pd.n_lines_pad + pd.n_chunks + pd.bc_nstage) * sizeof(double);
    if (pl->smem_bytes > 200 * 1024)
        return set_error(SPAMM_EINVAL, "spectrum has too many pixels blueward of the Balmer edge for the shared-memory stage");
    const int sb = (int)pl->smem_bytes;
#define SPAMM_SMEM_ATTR(...) \
    CUDA_TRY(cudaFuncSetAttribute(k_walkers<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb))
    SPAMM_SMEM_ATTR(0, true, false, 0); SPAMM_SMEM_ATTR(1, true, false, 0);
    SPAMM_SMEM_ATTR(0, false, false, 0); SPAMM_SMEM_ATTR(1, false, false, 0);
    SPAMM_SMEM_ATTR(0, true, true, 0); SPAMM_SMEM_ATTR(1, true, true, 0);
    SPAMM_SMEM_ATTR(0, false, true, 0); SPAMM_SMEM_ATTR(1, false, true, 0);
    SPAMM_SMEM_ATTR(0, true, false, 1); SPAMM_SMEM_ATTR(0, true, true, 1);
    SPAMM_SMEM_ATTR(0, true, false, 3); SPAMM_SMEM_ATTR(0, true, true, 3);
    SPAMM_SMEM_ATTR(0, true, false, 5); SPAMM_SMEM_ATTR(0, true, true, 5);
    SPAMM_SMEM_ATTR(0, true, false, 9); SPAMM_SMEM_ATTR(0, true, true, 9);
    SPAMM_SMEM_ATTR(0, true, false, 15); SPAMM_SMEM_ATTR(0, true, true, 15);
    SPAMM_SMEM_ATTR(0, true, false, 31); SPAMM_SMEM_ATTR(0, true, true, 31);
#undef SPAMM_SMEM_ATTR
    CUDA_TRY(cudaDeviceSynchronize());
    return SPAMM_OK;
}

extern "C" int spamm_plan_create(const SpammPlanDesc* desc, SpammPlan** out_plan) {
    if (!desc || !out_plan) return set_error(SPAMM_EINVAL, "null argument");
    SpammPlan* pl = new SpammPlan();
    int rc = plan_create_impl(desc, pl);
    if (rc) { spamm_plan_destroy(pl); *out_plan = nullptr; return rc; }
    *out_plan = pl;
    return SPAMM_OK;
}

extern "C" int spamm_plan_n_dim(const SpammPlan* pl) { return pl ? pl->dev.D : SPAMM_EINVAL; }
extern "C" int spamm_plan_n_pix(const SpammPlan* pl) { return pl ? pl->dev.P : SPAMM_EINVAL; }
extern "C" int spamm_plan_template_mode(const SpammPlan* pl) { return pl ? pl->template_mode : SPAMM_EINVAL; }
extern "C" int spamm_plan_bank_rows(const SpammPlan* pl) { return pl ? pl->bank_rows : SPAMM_EINVAL; }
extern "C" int64_t spamm_plan_launch_count(const SpammPlan* pl) { return pl ? pl->launches : 0; }
extern "C" int spamm_plan_set_walker_split(SpammPlan* pl, int32_t mode) {
    if (!pl) return set_error(SPAMM_EINVAL, "null plan");
    if (mode != 0 && mode != 1 && mode != 2 && mode != 4 && mode != 8)
        return set_error(SPAMM_EINVAL, "walker split must be 0 (off), 1 (auto), 2, 4 or 8");
    pl->split_mode = mode;
    return SPAMM_OK;
}

extern "C" int spamm_plan_status(SpammPlan* pl, void* stream) {
    if (!pl) return set_error(SPAMM_EINVAL, "null plan");
    int h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, pl->dev.status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return h;
}

// direct-mode workspace (per-walker broadened rows), allocated on first use
static int ensure_direct_ws(SpammPlan* pl) {
    if (pl->direct_chunk) return SPAMM_OK;
    const PlanDev& pd = pl->dev;
    int nt = pd.fe.n_templ + pd.host.n_templ;
    if (nt == 0) { pl->direct_chunk = pl->max_walkers; return SPAMM_OK; }
    size_t per_walker = (size_t)nt * ((size_t)pd.P + pl->conv_stride) * sizeof(double);
    size_t budget = (size_t)1 << 30;
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)pl->max_walkers, budget / per_walker));
    int ntm = std::max(pd.fe.n_templ, pd.host.n_templ);
    CUDA_TRY(cudaMalloc(&pl->d_conv, (size_t)chunk * ntm * pl->conv_stride * sizeof(double)));
    if (pd.fe.n_templ) {
        CUDA_TRY(cudaMalloc(&pl->d_fe_f, ((size_t)chunk * pd.fe.n_templ * pd.P + kPad) * sizeof(double)));
        CUDA_TRY(cudaMemset(pl->d_fe_f, 0, ((size_t)chunk * pd.fe.n_templ * pd.P + kPad) * sizeof(double)));
        CUDA_TRY(cudaMalloc(&pl->d_fe_nf, (size_t)chunk * pd.fe.n_templ * sizeof(double)));
    }
    if (pd.host.n_templ) {
        CUDA_TRY(cudaMalloc(&pl->d_host_f, ((size_t)chunk * pd.host.n_templ * pd.P + kPad) * sizeof(double)));
        CUDA_TRY(cudaMemset(pl->d_host_f, 0, ((size_t)chunk * pd.host.n_templ * pd.P + kPad) * sizeof(double)));
        CUDA_TRY(cudaMalloc(&pl->d_host_nf, (size_t)chunk * pd.host.n_templ * sizeof(double)));
    }
    CUDA_TRY(cudaMalloc(&pl->d_row_t, (size_t)chunk * ntm * sizeof(int)));
    CUDA_TRY(cudaMalloc(&pl->d_row_sig, (size_t)chunk * ntm * sizeof(int)));
    pl->direct_chunk = chunk;
    return SPAMM_OK;
}

static int direct_rows_for_chunk(SpammPlan* pl, const TemplDev& ts, const double* params_dev, int w0,
                                 int nW, double* d_f, double* d_nf, cudaStream_t st) {
    int n_rows = nW * ts.n_templ;
    k_rows_from_params<<<(n_rows + 127) / 128, 128, 0, st>>>(ts, params_dev, pl->dev.D, w0, nW,
                                                               pl->d_row_t, pl->d_row_sig);
    pl->launches += 1;
    // the widest kernel decides the shared-memory size: read the sigmas back (direct mode is the
    // cross-check path, not the throughput path)
    std::vector<int> sig(n_rows);
    CUDA_TRY(cudaMemcpyAsync(sig.data(), pl->d_row_sig, n_rows * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    int max_sigma = 1;
    for (int s : sig) max_sigma = std::max(max_sigma, s);
    return run_rows(pl, ts, pl->d_row_t, pl->d_row_sig, n_rows, max_sigma, pl->d_conv, pl->conv_stride,
                    d_f, d_nf, st);
}

// canonical = live components appear as NC < Fe < host < Balmer in the list
static bool canonical_order(const PlanDev& pd, uint32_t mask) {
    int last = -1;
    for (int q = 0; q < pd.n_comp; ++q) {
        if (!((mask >> q) & 1u)) continue;
        if (pd.comp_kind[q] < last) return false;
        last = pd.comp_kind[q];
    }
    return true;
}

// One launch of the fused kernel.  Ensembles too small to fill the GPU (fewer walkers than
// resident CTA slots) run one walker per thread-block cluster of 2/4/8 CTAs (SPLIT instantiation).
template <int MODE, bool CANON, bool SPLIT, int SPEC = 0>
static int launch_one(SpammPlan* pl, int n_walkers, int csize, cudaStream_t st, const PlanDev& pd,
                      const double* params_dev, int w0, double* out_dev, uint32_t mask, DirectRows dr,
                      PeerOut po) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_walkers * csize), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = pl->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)csize;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    if (SPLIT) { cfg.attrs = &attr; cfg.numAttrs = 1; }
    CUDA_TRY(cudaLaunchKernelEx(&cfg, k_walkers<MODE, CANON, SPLIT, SPEC>, pd, params_dev, w0, out_dev, mask, dr, po));
    return SPAMM_OK;
}

static int cluster_size_for(const SpammPlan* pl, int n_walkers) {
    if (pl->split_mode == 0) return 1;
    int slots = pl->n_sm * SPAMM_MIN_BLOCKS;
    int c = 1;
    while (c < 8 && n_walkers * c * 4 <= slots) c *= 2;       // keep the split grid within half the slots
    if (pl->split_mode > 1) c = pl->split_mode;          // forced (tests)
    return c;
}

// Which specialised instantiation (SPEC bit set, 0 = none) serves ln-posterior calls on this plan
static int spec_of(const SpammPlan* pl, const PlanDev& pd, uint32_t mask, const DirectRows& dr) {
    if (!pl->use_full_kernel || mask != (1u << pd.n_comp) - 1u || dr.fe_f || dr.host_f) return 0;
    if (!pd.exp_safe || !pd.nc_pair_ok || pd.lo == nullptr || pd.nc_broken) return 0;
    int spec = 0, last = -1;
    for (int q = 0; q < pd.n_comp; ++q) {
        const int kind = pd.comp_kind[q];
        if (kind <= last) return 0;                       // run_spamm's order only
        last = kind;
        if (kind == SPAMM_COMP_NUCLEAR) spec |= kSpecNc;
        else if (kind == SPAMM_COMP_FE) { if (pd.fe.n_templ != kFullFe) return 0; spec |= kSpecFe; }
        else if (kind == SPAMM_COMP_HOST) { if (pd.host.n_templ != kFullHost) return 0; spec |= kSpecHost; }
        else if (kind == SPAMM_COMP_BALMER) {
            if (!(pd.bc_on && pd.bpc_on && pd.line_type == SPAMM_LINE_LORENTZIAN && pd.series_on &&
                  pd.n_clusters == kFullClusters && pd.n_super == kFullSuper)) return 0;
            spec |= kSpecBalmer;
        } else if (kind == SPAMM_COMP_EXTINCTION) spec |= kSpecExt;
    }
    for (int k : kSpecs) if (k == spec) return spec;
    return 0;
}

extern "C" int spamm_plan_specialised(const SpammPlan* pl) {
    if (!pl) return SPAMM_EINVAL;
    DirectRows none{nullptr, nullptr, nullptr, nullptr};
    const bool bank = pl->template_mode != SPAMM_TEMPLATES_DIRECT;
    return bank && spec_of(pl, pl->dev, (1u << pl->dev.n_comp) - 1u, none) != 0 ? 1 : 0;
}

template <int MODE>
static int launch_k(SpammPlan* pl, bool canon, int n_walkers, cudaStream_t st, const PlanDev& pd,
                    const double* params_dev, int w0, double* out_dev, uint32_t mask, DirectRows dr,
                    PeerOut po) {
    const int c = cluster_size_for(pl, n_walkers);
    if (MODE == 0) {
#define SPAMM_LAUNCH_SPEC(S)                                                                                     \
    case S:                                                                                                      \
        return c > 1 ? launch_one<0, true, true, S>(pl, n_walkers, c, st, pd, params_dev, w0, out_dev, mask, dr, po) \
                     : launch_one<0, true, false, S>(pl, n_walkers, 1, st, pd, params_dev, w0, out_dev, mask, dr, po);
        switch (spec_of(pl, pd, mask, dr)) {
            SPAMM_LAUNCH_SPEC(1) SPAMM_LAUNCH_SPEC(3) SPAMM_LAUNCH_SPEC(5) SPAMM_LAUNCH_SPEC(9)
            SPAMM_LAUNCH_SPEC(15) SPAMM_LAUNCH_SPEC(31)
            default: break;
        }
#undef SPAMM_LAUNCH_SPEC
    }
    if (c > 1) {
        return canon ? launch_one<MODE, true, true>(pl, n_walkers, c, st, pd, params_dev, w0, out_dev, mask, dr, po)
                     : launch_one<MODE, false, true>(pl, n_walkers, c, st, pd, params_dev, w0, out_dev, mask, dr, po);
    }
    return canon ? launch_one<MODE, true, false>(pl, n_walkers, 1, st, pd, params_dev, w0, out_dev, mask, dr, po)
                 : launch_one<MODE, false, false>(pl, n_walkers, 1, st, pd, params_dev, w0, out_dev, mask, dr, po);
}

template <int MODE>
static int launch_walkers(SpammPlan* pl, uint32_t mask, const double* params_dev, int n_walkers,
                          double* out_dev, cudaStream_t st, bool force_direct,
                          PeerOut po = PeerOut{nullptr, 0, 0}) {
    const PlanDev& pd = pl->dev;
    bool need_fe = false, need_host = false;
    for (int q = 0; q < pd.n_comp; ++q) {
        if (!((mask >> q) & 1u)) continue;
        if (pd.comp_kind[q] == SPAMM_COMP_FE) need_fe = true;
        if (pd.comp_kind[q] == SPAMM_COMP_HOST) need_host = true;
    }
    bool direct = (need_fe || need_host) && (force_direct || pl->template_mode == SPAMM_TEMPLATES_DIRECT);
    if (!direct) {
        DirectRows dr{nullptr, nullptr, nullptr, nullptr};
        int rc = launch_k<MODE>(pl, canonical_order(pd, mask), n_walkers, st, pd, params_dev, 0, out_dev, mask, dr, po);
        if (rc) return rc;
        pl->launches += 1;
        return SPAMM_OK;
    }
    int rc = ensure_direct_ws(pl);
    if (rc) return rc;
    for (int w0 = 0; w0 < n_walkers; w0 += pl->direct_chunk) {
        int nW = std::min(pl->direct_chunk, n_walkers - w0);
        DirectRows dr{nullptr, nullptr, nullptr, nullptr};
        if (need_fe) {
            if ((rc = direct_rows_for_chunk(pl, pd.fe, params_dev, w0, nW, pl->d_fe_f, pl->d_fe_nf, st))) return rc;
            dr.fe_f = pl->d_fe_f; dr.fe_nf = pl->d_fe_nf;
        }
        if (need_host) {
            if ((rc = direct_rows_for_chunk(pl, pd.host, params_dev, w0, nW, pl->d_host_f, pl->d_host_nf, st))) return rc;
            dr.host_f = pl->d_host_f; dr.host_nf = pl->d_host_nf;
        }
        double* o = MODE == 0 ? out_dev : out_dev + (size_t)w0 * pd.P;
        if ((rc = launch_k<MODE>(pl, canonical_order(pd, mask), nW, st, pd, params_dev, w0, o, mask, dr, po))) return rc;
        pl->launches += 1;
    }
    return SPAMM_OK;
}

static uint32_t full_mask(const SpammPlan* pl) { return (1u << pl->dev.n_comp) - 1u; }

extern "C" int spamm_lnpost_batch(SpammPlan* pl, const double* params_dev, int32_t n_walkers,
                                  double* lnp_dev, void* stream) {
    if (!pl || !params_dev || !lnp_dev) return set_error(SPAMM_EINVAL, "null argument");
    if (n_walkers < 0) return set_error(SPAMM_EINVAL, "negative walker count");
    if (n_walkers == 0) return SPAMM_OK;
    if (!pl->has_data) return set_error(SPAMM_EINVAL, "plan was created without data flux/flux_error");
    return launch_walkers<0>(pl, full_mask(pl), params_dev, n_walkers, lnp_dev, (cudaStream_t)stream, false);
}

extern "C" int spamm_lnpost_batch_p2p(SpammPlan* pl, const double* params_dev, int32_t n_walkers,
                                      const uint64_t* peer_bufs_dev, int32_t n_peers, int64_t elem_offset,
                                      void* stream) {
    if (!pl || !params_dev || !peer_bufs_dev) return set_error(SPAMM_EINVAL, "null argument");
    if (n_walkers < 0 || n_peers < 1 || elem_offset < 0) return set_error(SPAMM_EINVAL, "bad argument");
    if (n_walkers == 0) return SPAMM_OK;
    if (!pl->has_data) return set_error(SPAMM_EINVAL, "plan was created without data flux/flux_error");
    PeerOut po{reinterpret_cast<const unsigned long long*>(peer_bufs_dev), n_peers, (long long)elem_offset};
    return launch_walkers<0>(pl, full_mask(pl), params_dev, n_walkers, nullptr, (cudaStream_t)stream, false, po);
}

extern "C" int spamm_lnpost_batch_host(SpammPlan* pl, const double* params_host, int32_t n_walkers,
                                       double* lnp_host) {
    if (!pl || !params_host || !lnp_host) return set_error(SPAMM_EINVAL, "null argument");
    if (n_walkers < 0) return set_error(SPAMM_EINVAL, "negative walker count");
    if (n_walkers == 0) return SPAMM_OK;
    if (!pl->has_data) return set_error(SPAMM_EINVAL, "plan was created without data flux/flux_error");
    if (!pl->own_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&pl->own_stream, cudaStreamNonBlocking));
        size_t np = (size_t)pl->max_walkers * pl->dev.D;
        CUDA_TRY(cudaMallocHost(&pl->h_params, np * sizeof(double)));
        CUDA_TRY(cudaMallocHost(&pl->h_lnp, (size_t)pl->max_walkers * sizeof(double)));
        CUDA_TRY(cudaMalloc(&pl->d_params, np * sizeof(double)));
        CUDA_TRY(cudaMalloc(&pl->d_lnp, (size_t)pl->max_walkers * sizeof(double)));
    }
    cudaStream_t st = pl->own_stream;
    // caller buffers that are already page-locked (cudaHostAlloc / cudaHostRegister / torch
    // pin_memory) are copied from/to directly; pageable ones go through the plan's pinned staging
    auto is_pinned = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    const bool pin_in = is_pinned(params_host), pin_out = is_pinned(lnp_host);
    for (int w0 = 0; w0 < n_walkers; w0 += pl->max_walkers) {
        int nW = std::min(pl->max_walkers, n_walkers - w0);
        size_t nb = (size_t)nW * pl->dev.D * sizeof(double);
        const double* src = params_host + (size_t)w0 * pl->dev.D;
        if (!pin_in) { memcpy(pl->h_params, src, nb); src = pl->h_params; }
        CUDA_TRY(cudaMemcpyAsync(pl->d_params, src, nb, cudaMemcpyHostToDevice, st));
        int rc = launch_walkers<0>(pl, full_mask(pl), pl->d_params, nW, pl->d_lnp, st, false);
        if (rc) return rc;
        double* dst = pin_out ? lnp_host + w0 : pl->h_lnp;
        CUDA_TRY(cudaMemcpyAsync(dst, pl->d_lnp, (size_t)nW * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (!pin_out) memcpy(lnp_host + w0, pl->h_lnp, (size_t)nW * sizeof(double));
    }
    return SPAMM_OK;
}

extern "C" int spamm_model_flux_batch(SpammPlan* pl, uint32_t mask, const double* params_dev,
                                      int32_t n_walkers, double* flux_dev, void* stream) {
    if (!pl || !params_dev || !flux_dev) return set_error(SPAMM_EINVAL, "null argument");
    if (n_walkers < 0) return set_error(SPAMM_EINVAL, "negative walker count");
    if (n_walkers == 0) return SPAMM_OK;
    mask &= full_mask(pl);
    if (!mask) return set_error(SPAMM_EINVAL, "empty component mask");
    cudaStream_t st = (cudaStream_t)stream;
    const PlanDev& pd = pl->dev;
    // flux() may be called with widths the prior (and so the bank) never admits
    bool force_direct = false;
    if (pl->template_mode == SPAMM_TEMPLATES_BANK) {
        CUDA_TRY(cudaMemsetAsync(pl->d_flag, 0, sizeof(int), st));
        bool any = false;
        for (int q = 0; q < pd.n_comp; ++q) {
            if (!((mask >> q) & 1u)) continue;
            const TemplDev* ts = pd.comp_kind[q] == SPAMM_COMP_FE ? &pd.fe
                                 : pd.comp_kind[q] == SPAMM_COMP_HOST ? &pd.host : nullptr;
            if (!ts) continue;
            int n = n_walkers * ts->n_templ;
            k_check_bank<<<(n + 127) / 128, 128, 0, st>>>(*ts, params_dev, pd.D, n_walkers, pl->d_flag);
            pl->launches += 1;
            any = true;
        }
        if (any) {
            int h = 0;
            CUDA_TRY(cudaMemcpyAsync(&h, pl->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            force_direct = h != 0;
        }
    }
    return launch_walkers<1>(pl, mask, params_dev, n_walkers, flux_dev, st, force_direct);
}

extern "C" int spamm_likelihood_batch(SpammPlan* pl, const double* model_flux_dev, int32_t n_walkers,
                                      double* lnl_dev, void* stream) {
    if (!pl || !model_flux_dev || !lnl_dev) return set_error(SPAMM_EINVAL, "null argument");
    if (n_walkers < 0) return set_error(SPAMM_EINVAL, "negative walker count");
    if (n_walkers == 0) return SPAMM_OK;
    if (!pl->has_data) return set_error(SPAMM_EINVAL, "plan was created without data flux/flux_error");
    k_likelihood<<<n_walkers, kThreads, 0, (cudaStream_t)stream>>>(pl->dev, model_flux_dev, lnl_dev);
    pl->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return SPAMM_OK;
}

// ---------------------------------------------------------------------------
// sampler + probe entry points
// ---------------------------------------------------------------------------
extern "C" int spamm_stretch_propose(const double* walkers_dev, int32_t K, int32_t D, int32_t half,
                                     double a, uint64_t seed, uint64_t iteration, double* q_dev,
                                     double* z_dev, void* stream) {
    if (!walkers_dev || !q_dev || !z_dev) return set_error(SPAMM_EINVAL, "null argument");
    if (K < 2 || (K & 1) || D < 1) return set_error(SPAMM_EINVAL, "need an even walker count");
    int Ns = K / 2;
    k_stretch_propose<<<(Ns + 127) / 128, 128, 0, (cudaStream_t)stream>>>(walkers_dev, K, D, half, a, seed,
                                                                          iteration, q_dev, z_dev);
    CUDA_TRY(cudaGetLastError());
    return SPAMM_OK;
}

extern "C" int spamm_stretch_accept(double* walkers_dev, double* lnp_dev, int32_t K, int32_t D,
                                    int32_t half, const double* q_dev, const double* z_dev,
                                    const double* lnp_q_dev, uint64_t seed, uint64_t iteration,
                                    int64_t* naccepted_dev, void* stream) {
    if (!walkers_dev || !lnp_dev || !q_dev || !z_dev || !lnp_q_dev) return set_error(SPAMM_EINVAL, "null argument");
    if (K < 2 || (K & 1) || D < 1) return set_error(SPAMM_EINVAL, "need an even walker count");
    int Ns = K / 2;
    k_stretch_accept<<<(Ns + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        walkers_dev, lnp_dev, K, D, half, q_dev, z_dev, lnp_q_dev, seed, iteration,
        reinterpret_cast<long long*>(naccepted_dev));
    CUDA_TRY(cudaGetLastError());
    return SPAMM_OK;
}

extern "C" int spamm_chain_record(const double* walkers_dev, const double* lnp_dev, int32_t K, int32_t D,
                                  int64_t iter, int64_t n_iter, double* chain_dev, double* lnp_chain_dev,
                                  void* stream) {
    if (!walkers_dev || !lnp_dev) return set_error(SPAMM_EINVAL, "null argument");
    k_chain_record<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(walkers_dev, lnp_dev, K, D, iter, n_iter,
                                                                      chain_dev, lnp_chain_dev);
    CUDA_TRY(cudaGetLastError());
    return SPAMM_OK;
}

// n_iter full iterations of the stretch move (both half-steps, optional chain record) queued on
// `stream` without returning to the host language in between
extern "C" int spamm_stretch_run(SpammPlan* pl, double* walkers_dev, double* lnp_dev, int32_t K, double a,
                                 uint64_t seed, uint64_t iter0, int32_t n_iter, double* q_dev, double* z_dev,
                                 double* lnp_q_dev, int64_t* naccepted_dev, double* chain_dev,
                                 double* lnp_chain_dev, int64_t chain_len, void* stream) {
    if (!pl || !walkers_dev || !lnp_dev || !q_dev || !z_dev || !lnp_q_dev)
        return set_error(SPAMM_EINVAL, "null argument");
    if (K < 2 || (K & 1) || n_iter < 0) return set_error(SPAMM_EINVAL, "need an even walker count");
    if (!pl->has_data) return set_error(SPAMM_EINVAL, "plan was created without data flux/flux_error");
    if ((chain_dev || lnp_chain_dev) && (int64_t)iter0 + n_iter > chain_len)
        return set_error(SPAMM_EINVAL, "chain buffer too short");
    const int D = pl->dev.D;
    for (int it = 0; it < n_iter; ++it) {
        const uint64_t iteration = iter0 + (uint64_t)it;
        for (int half = 0; half < 2; ++half) {
            int rc = spamm_stretch_propose(walkers_dev, K, D, half, a, seed, iteration, q_dev, z_dev, stream);
            if (rc) return rc;
            if ((rc = launch_walkers<0>(pl, full_mask(pl), q_dev, K / 2, lnp_q_dev, (cudaStream_t)stream, false)))
                return rc;
            if ((rc = spamm_stretch_accept(walkers_dev, lnp_dev, K, D, half, q_dev, z_dev, lnp_q_dev, seed,
                                           iteration, naccepted_dev, stream)))
                return rc;
        }
        if (chain_dev || lnp_chain_dev) {
            int rc = spamm_chain_record(walkers_dev, lnp_dev, K, D, (int64_t)iteration, chain_len, chain_dev,
                                        lnp_chain_dev, stream);
            if (rc) return rc;
        }
    }
    return SPAMM_OK;
}

extern "C" int spamm_fp64_peak_probe(int32_t iters, int32_t repeats, double* tflops_out) {
    if (!tflops_out || iters < 1 || repeats < 1) return set_error(SPAMM_EINVAL, "bad argument");
    int dev_id = 0;
    CUDA_TRY(cudaGetDevice(&dev_id));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev_id));
    double* sink;
    CUDA_TRY(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    int blocks = prop.multiProcessorCount * 8;
    k_dfma_probe<<<blocks, 256>>>(iters, sink);       // warm-up
    double best = 0.0;
    for (int r = 0; r < repeats; ++r) {
        CUDA_TRY(cudaEventRecord(e0));
        k_dfma_probe<<<blocks, 256>>>(iters, sink);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 8.0 * (double)iters * 256.0 * blocks;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *tflops_out = best;
    return SPAMM_OK;
}
