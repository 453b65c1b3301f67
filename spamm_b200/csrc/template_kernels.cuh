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
