// spamm_b200 -- template broadening kernels (sigma bank / per-walker rows)
// (part of the single translation unit spamm_b200.cu; reference call sites are cited as
// file:line relative to the reference root)
#pragma once
#include "plan_dev.cuh"
#include "device_math.cuh"
// ---------------------------------------------------------------------------
// setup kernels (plan creation)
// ---------------------------------------------------------------------------
// Gaussian broadening in ln-lambda space for a list of rows (template t, sigma_norm):
//   kernel = signal.gaussian(K*sigma, sigma)/(sqrt(2 pi) sigma)          Fe:291-293 / HG:298-300
//   conv   = np.convolve(log_flux, kernel, mode="same")                  Fe:304     / HG:311
// One CTA = one row x one tile of kConvTile = 1024 output points: 128 threads x kConvR = 8 CONSECUTIVE outputs
// each.  The kernel taps and the template tile (+ halo) are staged in shared memory; a thread keeps the 8
// template values under its outputs in registers and slides the window one point per tap, so a tap costs one
// 8-byte shared load (the new point) + one broadcast load (the tap) for 8 DFMA -- 3 shared-memory wavefronts
// per 8 warp-DFMA, below the FP64 pipe's 4 cycles: FP64-bound instead of the 3 wavefronts PER DFMA of one
// output per thread.  The tile is stored transposed in groups of 8 (element s at (s mod 8) S + s / 8) so that,
// for a given tap, the 32 lanes read consecutive words.  Every output accumulates its taps in ascending order
// with one FMA each, whatever the tiling (rows are bit-identical to the one-output-per-thread form).
__global__ void __launch_bounds__(kConvThreads)
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
        for (int m = m0 + threadIdx.x; m < min(m0 + kConvTile, N); m += kConvThreads) out[m] = CUDART_NAN;
        return;
    }
    const int M = ts.ksize * sig;
    const int Mp = (M + kConvR - 1) & ~(kConvR - 1);     // taps padded with zeros to a multiple of 8
    const int off = (M - 1) / 2;
    const int S = conv_group_stride(Mp);                 // words per residue class of the transposed tile
    double* kw = smem;                                   // [Mp]
    double* gs = smem + Mp;                              // [kConvR][S]
    const double* g = ts.logf + ts.off[t];
    const double sg = (double)sig;
    const double sig2 = 2 * sg * sg;
    const double nrm = ts.sqrt_2pi * sg;
    const double half = ((double)M - 1.0) / 2.0;
    for (int j = threadIdx.x; j < Mp; j += kConvThreads) {
        double n = (double)j - half;
        kw[j] = j < M ? exp(-(n * n) / sig2) / nrm : 0.0;
    }
    // output m needs g[m + off - j], j = 0..Mp-1: template indices [m0 + off - (Mp-1), m0 + kConvTile - 1 + off]
    // (tile position s <-> template index lo + s)
    const int lo = m0 + off - (Mp - 1);
    const int span = kConvTile + Mp - 1;
    for (int s = threadIdx.x; s < span; s += kConvThreads) {
        const int idx = lo + s;
        gs[(s & (kConvR - 1)) * S + (s >> 3)] = (idx >= 0 && idx < N) ? g[idx] : 0.0;
    }
    __syncthreads();
    // this thread's outputs m0 + 8 tid + q; at tap j output q needs tile position 8 tid + (Mp-1) + q - j
    double acc[kConvR], w[kConvR];
    const int c0 = kConvR * threadIdx.x + (Mp - 1);
#pragma unroll
    for (int q = 0; q < kConvR; ++q) {
        acc[q] = 0.0;
        const int s = c0 + q;
        w[q] = gs[(s & (kConvR - 1)) * S + (s >> 3)];
    }
    // c0 = 8 tid + Mp - 1 == 7 (mod 8): the point entering the window at tap j + 1 is s = c0 - j - 1, residue
    // class 6 - (j mod 8), word tid + (Mp - 8 - j) / 8 -- compile-time class inside the 8-fold unrolled body
    for (int j0 = 0; j0 < Mp; j0 += kConvR) {
        const double* gw = gs + threadIdx.x + ((Mp - kConvR - j0) >> 3);
#pragma unroll
        for (int u = 0; u < kConvR; ++u) {
            const double k = kw[j0 + u];
            // window value of output q at this tap sits in register (q - u) mod 8
#pragma unroll
            for (int q = 0; q < kConvR; ++q) acc[q] = fma(w[(q - u) & (kConvR - 1)], k, acc[q]);
            // slide: output 0's next point replaces the register output 7 just used
            if (u < kConvR - 1) w[(kConvR - 1 - u) & (kConvR - 1)] = gw[(kConvR - 2 - u) * S];
            else if (j0 + kConvR < Mp) w[0] = gw[(kConvR - 1) * S - 1];
        }
    }
#pragma unroll
    for (int q = 0; q < kConvR; ++q) {
        const int m = m0 + kConvR * threadIdx.x + q;
        if (m < N) out[m] = acc[q];
    }
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
    if (ts.rb_ptr != nullptr) {
        // rebin_spec: True (Fe:320-323 / HG:326-329): bin-averaged trapezoid integral of the convolved row,
        // a sparse matrix row built by the host (spamm_b200/utils/rebin_spec.py)
        const int* ptr = ts.rb_ptr + (size_t)t * (P + 1);
        double acc = 0.0;
        for (int k = ptr[i]; k < ptr[i + 1]; ++k) acc += ts.rb_w[k] * fp[ts.rb_idx[k]];
        frows[(size_t)r * P + i] = acc;
        return;
    }
    double left = ts.clamp ? fp[0] : 0.0;
    double right = ts.clamp ? fp[N - 1] : 0.0;
    frows[(size_t)r * P + i] = np_interp_at<true>(lnwl[i], xp, fp, N, left, right);
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

// rows for the direct mode: one row per (walker, template).  sig_cap > 0 (ln-posterior path): the widest kernel
// the width prior admits -- a walker beyond it fails the prior test of the walker kernel before its rows are
// looked at, so it gets the narrowest kernel instead of stretching the shared-memory stage of the whole batch.
__global__ void k_rows_from_params(TemplDev ts, const double* __restrict__ params, int D, int w0,
                                   int nW, int* __restrict__ row_t, int* __restrict__ row_sig, int sig_cap) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nW * ts.n_templ) return;
    int w = idx / ts.n_templ, t = idx % ts.n_templ;
    double width = params[(size_t)(w0 + w) * D + ts.param_off + ts.n_templ];
    double sn = sigma_norm_of(ts, t, width);
    int sig = (sn >= 1.0 && sn <= 1.0e6) ? (int)sn : 0;
    if (sig_cap > 0 && sig > sig_cap) sig = 1;
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
