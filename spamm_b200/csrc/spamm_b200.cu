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

#include "plan_dev.cuh"
#include "device_math.cuh"
#include "template_kernels.cuh"
#include "walker_kernel.cuh"
#include "sampler_kernels.cuh"

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
    int sig_cap_fe = 0, sig_cap_host = 0;   // widest sigma_norm the width priors admit (0: no bounded prior)
    int last_bank_rows = 0;
    int last_sig_cap = 0;
    int split_mode = 1;          // 0: never split a walker; 1: auto; 2/4/8: forced cluster size; 3: forced two-part split
    int* h_status = nullptr;           // pinned copy of the device status word (mapped; see deferred_status)
    double* d_part_chi = nullptr;      // walker parts: per-chunk chi^2 partials [max_walkers][n_chunks]
    unsigned* d_part_cnt = nullptr;    //               arrival counters [max_walkers]
    int parts_max_walkers = getenv("SPAMM_PARTS_MAX") ? atoi(getenv("SPAMM_PARTS_MAX")) : 0;   // auto: mixed launches up to this many waves (0: default)
    int parts_mixed = getenv("SPAMM_PARTS_MIXED") ? atoi(getenv("SPAMM_PARTS_MIXED")) : 1;     // 0: never mix; 2: mix from one wave on
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
    const int Mp = (ksize * max_sigma + kConvR - 1) & ~(kConvR - 1);
    return (Mp + kConvR * conv_group_stride(Mp)) * (int)sizeof(double);
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
    k_conv_rows<<<g1, kConvThreads, smem, st>>>(ts, d_row_t, d_row_sig, d_conv, stride);
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
    ts.rb_ptr = nullptr; ts.rb_idx = nullptr; ts.rb_w = nullptr;
    if (src.rebin_indptr && src.rebin_index && src.rebin_weight) {
        const size_t np1 = (size_t)ts.n_templ * (pl->dev.P + 1);
        const int nnz = src.rebin_indptr[np1 - 1];
        for (int t = 0; t < ts.n_templ; ++t) {
            const int32_t* ptr = src.rebin_indptr + (size_t)t * (pl->dev.P + 1);
            for (int i = 0; i < pl->dev.P; ++i) {
                if (ptr[i + 1] < ptr[i] || ptr[i + 1] > nnz) return set_error(SPAMM_EINVAL, "rebin operator: bad row pointers");
                for (int k = ptr[i]; k < ptr[i + 1]; ++k)
                    if (src.rebin_index[k] < 0 || src.rebin_index[k] >= ts.n_pts[t])
                        return set_error(SPAMM_EINVAL, "rebin operator: column outside the template");
            }
        }
        if ((rc = upload(pl, src.rebin_indptr, np1, &ts.rb_ptr))) return rc;
        if ((rc = upload(pl, src.rebin_index, (size_t)nnz, &ts.rb_idx))) return rc;
        if ((rc = upload(pl, src.rebin_weight, (size_t)nnz, &ts.rb_w))) return rc;
    }

    *bank_ok = false;
    pl->last_sig_cap = 0;
    if (std::isfinite(width_hi)) {
        int cap = 0;
        for (int t = 0; t < ts.n_templ; ++t) {
            double sn = host_sigma_norm(src, ts.bin[t], width_hi);
            if (!(sn >= 1.0) || sn > 1.0e6) { cap = 0; break; }
            cap = std::max(cap, (int)sn);
        }
        pl->last_sig_cap = cap;
    }
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
    cudaFree(pl->d_part_chi); cudaFree(pl->d_part_cnt);
    if (pl->h_params) cudaFreeHost(pl->h_params);
    if (pl->h_lnp) cudaFreeHost(pl->h_lnp);
    if (pl->h_status) cudaFreeHost(pl->h_status);
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
    CUDA_TRY(cudaHostAlloc(&pl->h_status, sizeof(int), cudaHostAllocMapped));
    *pl->h_status = 0;
    CUDA_TRY(cudaHostGetDevicePointer(&pd.status_host, pl->h_status, 0));
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
        pl->sig_cap_fe = pl->last_sig_cap;
        all_bank = all_bank && ok;
    }
    if (seen[SPAMM_COMP_HOST]) {
        bool ok = false;
        double hi = d->prior_hi ? d->prior_hi[off_host + d->host.n_templates] : INFINITY;
        if ((rc = setup_template_set(pl, d->host, off_host, hi, want_bank, &pd.host, &ok))) return rc;
        pd.host_rows = ok ? pl->last_bank_rows : 0;
        pl->sig_cap_host = pl->last_sig_cap;
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
            pd.bc_rb_ptr = nullptr; pd.bc_rb_idx = nullptr; pd.bc_rb_w = nullptr;
            const bool bc_rebin = d->bc_rebin_indptr && d->bc_rebin_index && d->bc_rebin_weight;
            if (bc_rebin) {
                const int nnz = d->bc_rebin_indptr[P];
                for (int i = 0; i < P; ++i) {
                    if (d->bc_rebin_indptr[i + 1] < d->bc_rebin_indptr[i] || d->bc_rebin_indptr[i + 1] > nnz)
                        return set_error(SPAMM_EINVAL, "log_conv rebin operator: bad row pointers");
                    for (int k = d->bc_rebin_indptr[i]; k < d->bc_rebin_indptr[i + 1]; ++k) {
                        if (d->bc_rebin_index[k] < 0 || d->bc_rebin_index[k] >= P)
                            return set_error(SPAMM_EINVAL, "log_conv rebin operator: column outside the grid");
                        if (k > d->bc_rebin_indptr[i] && d->bc_rebin_index[k] <= d->bc_rebin_index[k - 1])
                            return set_error(SPAMM_EINVAL, "log_conv rebin operator: columns must ascend within a row");
                    }
                }
                if ((rc = upload(pl, d->bc_rebin_indptr, (size_t)P + 1, &pd.bc_rb_ptr))) return rc;
                if ((rc = upload(pl, d->bc_rebin_index, (size_t)nnz, &pd.bc_rb_idx))) return rc;
                if ((rc = upload(pl, d->bc_rebin_weight, (size_t)nnz, &pd.bc_rb_w))) return rc;
            }
            // stage f0 up to the furthest source pixel of (nearest pixel to the edge) + 2
            int ibe = 0;
            double best = INFINITY;
            for (int i = 0; i < P; ++i) {
                double v = std::fabs(d->wl[i] - d->balmer_edge);
                if (v < best) { best = v; ibe = i; }
            }
            int itop = std::min(ibe + 2, P - 1), top = 0;
            for (int i = 0; i <= itop; ++i) {
                top = std::max(top, std::max(idx[i].y, idx[i].w));
                if (bc_rebin && d->bc_rebin_indptr[i + 1] > d->bc_rebin_indptr[i])
                    top = std::max(top, (int)d->bc_rebin_index[d->bc_rebin_indptr[i + 1] - 1]);
            }
            pd.bc_nstage = std::min(P, top + 1);
        }
        // insertion point of the Balmer edge: start of the per-walker edge-pixel search
        pd.edge_guess = (int)(std::lower_bound(d->wl, d->wl + P, d->balmer_edge) - d->wl);
    }
    // walker parts: chunk boundary just blueward of the bluest Balmer edge the velocity-offset prior admits
    pd.part_chunk = -1;
    pd.part_stage0 = 0;
    if (seen[SPAMM_COMP_BALMER] && pd.bc_on && pd.bpc_on && d->prior_lo && d->prior_hi) {
        int o_bal = -1;
        for (int q = 0; q < pd.n_comp; ++q) if (pd.comp_kind[q] == SPAMM_COMP_BALMER) o_bal = pd.comp_off[q];
        const double lmax = std::max(std::fabs(d->prior_lo[o_bal + 3]), std::fabs(d->prior_hi[o_bal + 3]));
        const double edge_min = d->balmer_edge * (1.0 - lmax / std::min(d->c_kms, d->c_cgs));
        if (std::isfinite(edge_min)) {
            const int i_min = (int)(std::lower_bound(d->wl, d->wl + P, edge_min) - d->wl) - 2;
            const int pc = i_min / kChunk;
            if (i_min > 0 && pc >= 1 && pc < pd.n_chunks) {
                std::vector<int4> idx(P);
                CUDA_TRY(cudaMemcpy(idx.data(), pd.bc_idx, (size_t)P * sizeof(int4), cudaMemcpyDeviceToHost));
                int lo = pc * kChunk;
                for (int i = pc * kChunk; i < P && i < pd.bc_nstage + kChunk; ++i)
                    lo = std::min(lo, std::min(idx[i].x, idx[i].z));
                pd.part_chunk = pc;
                pd.part_stage0 = std::max(lo, 0) & ~1;
                CUDA_TRY(cudaMalloc(&pl->d_part_chi, (size_t)pl->max_walkers * pd.n_chunks * sizeof(double)));
                CUDA_TRY(cudaMalloc(&pl->d_part_cnt, (size_t)pl->max_walkers * sizeof(unsigned)));
                CUDA_TRY(cudaMemset(pl->d_part_cnt, 0, (size_t)pl->max_walkers * sizeof(unsigned)));
            }
        }
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
        pd.bc_pair_ok = 0;
        pd.bc_dhnu_max = 0.0;
        if (ok && o_bal >= 0 && pd.bc_on && pd.bc_nstage >= 2) {
            const double taumax = std::max(std::fabs(d->prior_lo[o_bal + 2]), std::fabs(d->prior_hi[o_bal + 2]));
            double dt3 = 0.0, dh = 0.0;
            for (int k = 0; k + 1 < pd.bc_nstage; k += 2) {
                const double a0 = std::pow(d->wl[k] / d->balmer_edge, 3.0), a1 = std::pow(d->wl[k + 1] / d->balmer_edge, 3.0);
                dt3 = std::max(dt3, std::fabs(a1 - a0));
                dh = std::max(dh, std::fabs(d->h_cgs * (d->c_aa / d->wl[k + 1]) - d->h_cgs * (d->c_aa / d->wl[k])));
            }
            pd.bc_pair_ok = taumax * dt3 <= 0.004 ? 1 : 0;
            pd.bc_dhnu_max = dh;
        }
        double dmax = 0.0;
        for (int i = 0; i + 1 < P; ++i) dmax = std::max(dmax, std::fabs(std::log(d->wl[i + 1] / d->wl[i])));
        pd.nc_pair_ok = (P % 2 == 0 && smax * dmax <= 0.016) ? 1 : 0;
    }

    // shared memory of the fused kernel
    pl->smem_bytes = sizeof(WalkerCtx) + (size_t)(5 * pd.n_lines_pad + pd.n_chunks + pd.bc_nstage) * sizeof(double);
    if (pl->smem_bytes > 200 * 1024)
        return set_error(SPAMM_EINVAL, "spectrum has too many pixels blueward of the Balmer edge for the shared-memory stage");
    // The dynamic shared-memory limit is a per-function, process-wide attribute: never lower it (an
    // earlier, larger plan may still be live), keep a high-water mark.
    static int smem_high_water_dev[64] = {0};                 // per device
    int& smem_high_water = smem_high_water_dev[dev_id & 63];
    if ((int)pl->smem_bytes > smem_high_water) {
        const int sb = (int)pl->smem_bytes;
#define SPAMM_SMEM_ATTR(...) \
    CUDA_TRY(cudaFuncSetAttribute(k_walkers<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb))
        SPAMM_SMEM_ATTR(0, true, false, 0); SPAMM_SMEM_ATTR(1, true, false, 0);
        SPAMM_SMEM_ATTR(0, false, false, 0); SPAMM_SMEM_ATTR(1, false, false, 0);
        SPAMM_SMEM_ATTR(0, true, true, 0); SPAMM_SMEM_ATTR(1, true, true, 0);
        SPAMM_SMEM_ATTR(0, false, true, 0); SPAMM_SMEM_ATTR(1, false, true, 0);
#define SPAMM_SMEM_SPEC(S) SPAMM_SMEM_ATTR(0, true, false, S); SPAMM_SMEM_ATTR(0, true, true, S);
        SPAMM_SPEC_LIST(SPAMM_SMEM_SPEC)
#undef SPAMM_SMEM_SPEC
#define SPAMM_SMEM_PARTS(S) SPAMM_SMEM_ATTR(0, true, false, S | kSpecParts);
        SPAMM_PARTS_LIST(SPAMM_SMEM_PARTS)
#undef SPAMM_SMEM_PARTS
#undef SPAMM_SMEM_ATTR
        smem_high_water = sb;
    }
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
    if (mode != 0 && mode != 1 && mode != 2 && mode != 3 && mode != 4 && mode != 8)
        return set_error(SPAMM_EINVAL, "walker split must be 0 (off), 1 (auto), 2, 4, 8 (cluster) or 3 (two parts)");
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


// The device status word is sticky (kernels only OR bits in).  A non-zero word means some walker needed a
// table entry the plan does not hold and its ln-posterior is not the model's -- that must surface as an
// error, never as a number.  A kernel that sets a bit also writes 1 to a mapped pinned host word, so the host
// notices for free: synchronous entry points look at it after their synchronisation (check_status),
// asynchronous ones at the start of the next call on the plan (deferred_status); spamm_plan_check does
// the same on demand.  The details come from the device word.
static int status_error(SpammPlan* pl, cudaStream_t st) {
    int h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, pl->dev.status, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaMemsetAsync(pl->dev.status, 0, sizeof(int), st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *reinterpret_cast<volatile int*>(pl->h_status) = 0;
    if (h == 0) return SPAMM_OK;
    std::string msg = "device status";
    if (h & kStatSigmaRange) msg += ": a template width is outside the broadening bank (sigma_norm < 1 or beyond the prior's maximum)";
    if (h & kStatBcKernel) msg += ": the Balmer-continuum log_conv kernel is wider than one pixel (unsupported grid/width)";
    if (h & kStatBcStage) msg += ": the Balmer edge pixel lies beyond the staged continuum samples";
    if (h & kStatPeerTimeout) msg += ": a peer rank's results did not arrive within 5 s (multi-GPU exchange)";
    return set_error(SPAMM_ERANGE, msg);
}

static int deferred_status(SpammPlan* pl, cudaStream_t st) {
    return *reinterpret_cast<volatile int*>(pl->h_status) ? status_error(pl, st) : SPAMM_OK;
}

// `st` was just synchronised
static int check_status(SpammPlan* pl, cudaStream_t st) { return deferred_status(pl, st); }

extern "C" int spamm_plan_check(SpammPlan* pl, void* stream) {
    if (!pl) return set_error(SPAMM_EINVAL, "null plan");
    const int h = spamm_plan_status(pl, stream);
    if (h < 0) return h;
    return h ? status_error(pl, (cudaStream_t)stream) : SPAMM_OK;
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

// sig_cap > 0: ln-posterior path with a bounded width prior -- the shared-memory stage is sized for the prior's
// widest kernel and nothing is read back (the call stays asynchronous); 0: flux() with arbitrary widths
static int direct_rows_for_chunk(SpammPlan* pl, const TemplDev& ts, const double* params_dev, int w0,
                                 int nW, double* d_f, double* d_nf, cudaStream_t st, int sig_cap) {
    int n_rows = nW * ts.n_templ;
    k_rows_from_params<<<(n_rows + 127) / 128, 128, 0, st>>>(ts, params_dev, pl->dev.D, w0, nW,
                                                               pl->d_row_t, pl->d_row_sig, sig_cap);
    pl->launches += 1;
    if (sig_cap > 0)
        return run_rows(pl, ts, pl->d_row_t, pl->d_row_sig, n_rows, sig_cap, pl->d_conv, pl->conv_stride, d_f, d_nf, st);
    // the widest kernel decides the shared-memory size: read the sigmas back (synchronises the stream)
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

static int parts_count(const SpammPlan* pl, int n_walkers);

// One launch of the fused kernel.  Ensembles too small to fill the GPU (fewer walkers than
// resident CTA slots) run one walker per thread-block cluster of 2/4/8 CTAs (SPLIT instantiation).
template <int MODE, bool CANON, bool SPLIT, int SPEC = 0>
static int launch_one(SpammPlan* pl, int n_walkers, int csize, cudaStream_t st, const PlanDev& pd,
                      const double* params_dev, int w0, double* out_dev, uint32_t mask, DirectRows dr,
                      PeerOut po, Proposal pr) {
    // (SPEC & kSpecParts: the launch's last n_parts walkers run as two independent CTAs each -- red parts first --
    //  after the n_whole walkers that run one CTA each; no cluster)
    int n_parts = 0;
    if (SPEC & kSpecParts) n_parts = parts_count(pl, n_walkers);
    PartsWs pw{pl->d_part_chi, pl->d_part_cnt, n_walkers - n_parts, n_parts};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((SPEC & kSpecParts) ? n_walkers + n_parts : n_walkers * csize), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = pl->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)csize;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    if (SPLIT) { cfg.attrs = &attr; cfg.numAttrs = 1; }
    CUDA_TRY(cudaLaunchKernelEx(&cfg, k_walkers<MODE, CANON, SPLIT, SPEC>, pd, params_dev, w0, out_dev, mask, dr, po, pw, pr));
    return SPAMM_OK;
}

static int cluster_size_for(const SpammPlan* pl, int n_walkers) {
    if (pl->split_mode == 0 || pl->split_mode == 3) return 1;
    int slots = pl->n_sm * SPAMM_MIN_BLOCKS;
    int c = 1;
    while (c < 8 && n_walkers * c * 5 <= slots * 2) c *= 2;   // keep the split grid within ~0.8 of the slots
                                                              // (profiles/r02_split_table.txt: 96 walkers x4 35 us, x2 41)
    if (pl->split_mode > 1) c = pl->split_mode;          // forced (tests)
    return c;
}

// How many walkers of an n-walker launch run as two independent CTAs (kSpecParts; the rest one CTA each).
// Forced by split mode 3 (all of them).  Automatic (profiles/r02_split_table.txt): below a wave and a half of
// CTAs every walker -- the two parts of a walker can land on different SMs, which evens out the 3-versus-4
// walkers per SM of a partial wave -- and for a few waves the walkers of the last, partial wave only, so its tail
// is made of finer work items; long launches (where CTAs of all phases overlap anyway) and ensembles small
// enough for the cluster split are left alone.
static int parts_count(const SpammPlan* pl, int n_walkers) {
    const PlanDev& pd = pl->dev;
    if (pd.part_chunk < 0 || !pl->d_part_chi || n_walkers > pl->max_walkers) return 0;
    if (pl->split_mode == 3) return n_walkers;
    if (pl->split_mode != 1) return 0;
    const int slots = pl->n_sm * SPAMM_MIN_BLOCKS;
    if (5 * n_walkers <= slots) return 0;                                 // below ~120 walkers: cluster split
    if (2 * n_walkers <= 3 * slots && pl->parts_mixed < 2) return n_walkers;
    const int max_waves = pl->parts_max_walkers > 0 ? pl->parts_max_walkers : 4;
    if (!pl->parts_mixed || n_walkers > max_waves * slots) return 0;
    return n_walkers % slots;                                             // the partial wave
}

static bool parts_for(const SpammPlan* pl, const PlanDev& pd, int n_walkers) {
    (void)pd;
    return parts_count(pl, n_walkers) > 0;
}

// Which specialised instantiation (SPEC bit set, 0 = none) serves ln-posterior calls on this plan
static int spec_of(const SpammPlan* pl, const PlanDev& pd, uint32_t mask, const DirectRows& dr) {
    if (!pl->use_full_kernel || mask != (1u << pd.n_comp) - 1u || dr.fe_f || dr.host_f) return 0;
    if (!pd.exp_safe || !pd.nc_pair_ok || pd.lo == nullptr || pd.nc_broken) return 0;
    if (pd.bc_rb_ptr != nullptr) return 0;            // rebin_spec: True runs the generic instantiation
    int spec = 0, last = -1;
    for (int q = 0; q < pd.n_comp; ++q) {
        const int kind = pd.comp_kind[q];
        if (kind <= last) return 0;                       // run_spamm's order only
        last = kind;
        if (kind == SPAMM_COMP_NUCLEAR) spec |= kSpecNc;
        else if (kind == SPAMM_COMP_FE) { if (pd.fe.n_templ != kFullFe) return 0; spec |= kSpecFe; }
        else if (kind == SPAMM_COMP_HOST) { if (pd.host.n_templ != kFullHost) return 0; spec |= kSpecHost; }
        else if (kind == SPAMM_COMP_BALMER) {
            if (pd.bpc_on && !(pd.line_type == SPAMM_LINE_LORENTZIAN && pd.series_on &&
                               pd.n_clusters == kFullClusters && pd.n_super == kFullSuper)) return 0;
            spec |= kSpecBalmer;
            if (!pd.bpc_on) spec |= kSpecNoBpc;
            if (!pd.bc_on) spec |= kSpecNoBc;
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
                    PeerOut po, Proposal pr) {
    if (MODE == 0 && parts_for(pl, pd, n_walkers)) {
#define SPAMM_LAUNCH_PARTS(S)                                                                                    \
    case S:                                                                                                      \
        return launch_one<0, true, false, S | kSpecParts>(pl, n_walkers, 2, st, pd, params_dev, w0, out_dev, mask, dr, po, pr);
        switch (spec_of(pl, pd, mask, dr)) {
            SPAMM_PARTS_LIST(SPAMM_LAUNCH_PARTS)
            default: break;
        }
#undef SPAMM_LAUNCH_PARTS
    }
    const int c = cluster_size_for(pl, n_walkers);
    if (MODE == 0) {
#define SPAMM_LAUNCH_SPEC(S)                                                                                     \
    case S:                                                                                                      \
        return c > 1 ? launch_one<0, true, true, S>(pl, n_walkers, c, st, pd, params_dev, w0, out_dev, mask, dr, po, pr) \
                     : launch_one<0, true, false, S>(pl, n_walkers, 1, st, pd, params_dev, w0, out_dev, mask, dr, po, pr);
        switch (spec_of(pl, pd, mask, dr)) {
            SPAMM_SPEC_LIST(SPAMM_LAUNCH_SPEC)
            default: break;
        }
#undef SPAMM_LAUNCH_SPEC
    }
    if (c > 1) {
        return canon ? launch_one<MODE, true, true>(pl, n_walkers, c, st, pd, params_dev, w0, out_dev, mask, dr, po, pr)
                     : launch_one<MODE, false, true>(pl, n_walkers, c, st, pd, params_dev, w0, out_dev, mask, dr, po, pr);
    }
    return canon ? launch_one<MODE, true, false>(pl, n_walkers, 1, st, pd, params_dev, w0, out_dev, mask, dr, po, pr)
                 : launch_one<MODE, false, false>(pl, n_walkers, 1, st, pd, params_dev, w0, out_dev, mask, dr, po, pr);
}

template <int MODE>
static int launch_walkers(SpammPlan* pl, uint32_t mask, const double* params_dev, int n_walkers,
                          double* out_dev, cudaStream_t st, bool force_direct,
                          PeerOut po = PeerOut{nullptr, 0, 0, nullptr, 0u, 0, 0ull},
                          Proposal pr = Proposal{nullptr, 0, 0, 0, 0.0, 0ull, 0ull}) {
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
        int rc = launch_k<MODE>(pl, canonical_order(pd, mask), n_walkers, st, pd, params_dev, 0, out_dev, mask, dr, po, pr);
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
            if ((rc = direct_rows_for_chunk(pl, pd.fe, params_dev, w0, nW, pl->d_fe_f, pl->d_fe_nf, st,
                                            MODE == 0 && pd.lo ? pl->sig_cap_fe : 0))) return rc;
            dr.fe_f = pl->d_fe_f; dr.fe_nf = pl->d_fe_nf;
        }
        if (need_host) {
            if ((rc = direct_rows_for_chunk(pl, pd.host, params_dev, w0, nW, pl->d_host_f, pl->d_host_nf, st,
                                            MODE == 0 && pd.lo ? pl->sig_cap_host : 0))) return rc;
            dr.host_f = pl->d_host_f; dr.host_nf = pl->d_host_nf;
        }
        double* o = MODE == 0 ? out_dev : out_dev + (size_t)w0 * pd.P;
        if ((rc = launch_k<MODE>(pl, canonical_order(pd, mask), nW, st, pd, params_dev, w0, o, mask, dr, po, pr))) return rc;
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
    int rc = deferred_status(pl, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_walkers<0>(pl, full_mask(pl), params_dev, n_walkers, lnp_dev, (cudaStream_t)stream, false);
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
        if ((rc = check_status(pl, st))) return rc;
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
        reinterpret_cast<long long*>(naccepted_dev), PeerIn{nullptr, 0, 0, 0, 1, 0ull, nullptr, nullptr},
        StretchFused{0, 0.0, nullptr, nullptr, 0, 0});
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
    const int D = pl->dev.D, Ns = K / 2;
    cudaStream_t st = (cudaStream_t)stream;
    // Pre-broadened bank: two launches per half-step -- the walker kernel makes the proposal itself (Proposal) and
    // the accept kernel re-makes it from the same counter and records the chain (StretchFused).  Per-walker
    // broadening needs the proposals in memory for its row kernels and keeps the propose / record launches.
    const bool fused = pl->template_mode != SPAMM_TEMPLATES_DIRECT;
    for (int it = 0; it < n_iter; ++it) {
        const uint64_t iteration = iter0 + (uint64_t)it;
        for (int half = 0; half < 2; ++half) {
            int rc;
            if (!fused) {
                if ((rc = spamm_stretch_propose(walkers_dev, K, D, half, a, seed, iteration, q_dev, z_dev, stream)))
                    return rc;
                if ((rc = launch_walkers<0>(pl, full_mask(pl), q_dev, Ns, lnp_q_dev, st, false))) return rc;
                if ((rc = spamm_stretch_accept(walkers_dev, lnp_dev, K, D, half, q_dev, z_dev, lnp_q_dev, seed,
                                               iteration, naccepted_dev, stream)))
                    return rc;
                continue;
            }
            Proposal pr{walkers_dev, K, half, 0, a, seed, iteration};
            if ((rc = launch_walkers<0>(pl, full_mask(pl), walkers_dev, Ns, lnp_q_dev, st, false,
                                        PeerOut{nullptr, 0, 0, nullptr, 0u, 0, 0ull}, pr)))
                return rc;
            k_stretch_accept<<<(Ns + 127) / 128, 128, 0, st>>>(
                walkers_dev, lnp_dev, K, D, half, nullptr, nullptr, lnp_q_dev, seed, iteration,
                reinterpret_cast<long long*>(naccepted_dev), PeerIn{nullptr, 0, 0, 0, 1, 0ull, nullptr, nullptr},
                StretchFused{1, a, chain_dev, lnp_chain_dev, (long long)chain_len, (long long)iteration});
            CUDA_TRY(cudaGetLastError());
        }
        if (!fused && (chain_dev || lnp_chain_dev)) {
            int rc = spamm_chain_record(walkers_dev, lnp_dev, K, D, (int64_t)iteration, chain_len, chain_dev,
                                        lnp_chain_dev, stream);
            if (rc) return rc;
        }
    }
    return SPAMM_OK;
}

// ---------------------------------------------------------------------------
// multi-GPU: walker shards, results exchanged by the compute kernel itself
// ---------------------------------------------------------------------------
// Region of one rank (device memory, mapped into every peer by CUDA IPC or handed in as raw pointers):
//   words [0, kFlagWords)                       flag block, word r = last epoch rank r has completed
//   words [kFlagWords, kFlagWords + max_n)      result buffer of even epochs
//   words [kFlagWords + max_n, ... + 2 max_n)   result buffer of odd epochs
// Two buffers suffice: a rank can start epoch e + 2 only after it has seen every peer's flag for e + 1, which
// a peer raises after (in stream order) it consumed the results of epoch e.
struct SpammExchange {
    int n_ranks = 1, rank = 0;
    int64_t max_n = 0;
    uint64_t epoch = 0;                    // identical on every rank: all ranks make the same calls
    void* region = nullptr;                // this rank's region (cudaMalloc)
    std::vector<void*> peer;               // every rank's region as mapped here (peer[rank] == region)
    std::vector<bool> opened;              // opened through cudaIpcOpenMemHandle
    unsigned long long* d_peer_ptrs = nullptr;   // device copy of `peer`
    unsigned* d_done = nullptr;            // retired-walker counter of the running call
    bool connected = false;
    bool pdl = getenv("SPAMM_EXCHANGE_PDL") ? atoi(getenv("SPAMM_EXCHANGE_PDL")) != 0 : true;
    // staging of the *_host entry point
    double *h_params = nullptr, *h_lnp = nullptr, *d_params = nullptr;
    cudaStream_t own_stream = nullptr;
};

extern "C" void spamm_shard_bounds(int64_t n, int32_t n_ranks, int32_t rank, int64_t* lo, int64_t* hi) {
    const int64_t per = n_ranks > 0 ? (n + n_ranks - 1) / n_ranks : n;
    const int64_t a = std::min<int64_t>((int64_t)rank * per, n);
    if (lo) *lo = a;
    if (hi) *hi = std::min<int64_t>(a + per, n);
}

extern "C" int64_t spamm_exchange_region_bytes(int64_t max_walkers) {
    return (int64_t)sizeof(double) * (kFlagWords + 2 * max_walkers);
}

extern "C" void spamm_exchange_destroy(SpammExchange* ex) {
    if (!ex) return;
    for (size_t r = 0; r < ex->peer.size(); ++r)
        if (ex->opened[r] && ex->peer[r]) cudaIpcCloseMemHandle(ex->peer[r]);
    cudaFree(ex->region); cudaFree(ex->d_peer_ptrs); cudaFree(ex->d_done); cudaFree(ex->d_params);
    if (ex->h_params) cudaFreeHost(ex->h_params);
    if (ex->h_lnp) cudaFreeHost(ex->h_lnp);
    if (ex->own_stream) cudaStreamDestroy(ex->own_stream);
    delete ex;
}

static int exchange_create_impl(SpammExchange* ex, int32_t n_ranks, int32_t rank, int64_t max_walkers,
                                void* ipc_handle_out) {
    if (n_ranks < 1 || n_ranks > kFlagWords || rank < 0 || rank >= n_ranks || max_walkers < 1)
        return set_error(SPAMM_EINVAL, "exchange: need 1 <= n_ranks <= 64, 0 <= rank < n_ranks, max_walkers >= 1");
    ex->n_ranks = n_ranks; ex->rank = rank; ex->max_n = max_walkers;
    const size_t bytes = (size_t)spamm_exchange_region_bytes(max_walkers);
    CUDA_TRY(cudaMalloc(&ex->region, bytes));
    CUDA_TRY(cudaMemset(ex->region, 0, bytes));
    CUDA_TRY(cudaMalloc(&ex->d_done, sizeof(unsigned)));
    CUDA_TRY(cudaMemset(ex->d_done, 0, sizeof(unsigned)));
    CUDA_TRY(cudaMalloc(&ex->d_peer_ptrs, (size_t)n_ranks * sizeof(unsigned long long)));
    CUDA_TRY(cudaDeviceSynchronize());
    ex->peer.assign(n_ranks, nullptr);
    ex->opened.assign(n_ranks, false);
    ex->peer[rank] = ex->region;
    if (ipc_handle_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "SPAMM_IPC_HANDLE_BYTES");
        cudaIpcMemHandle_t h;
        CUDA_TRY(cudaIpcGetMemHandle(&h, ex->region));
        memcpy(ipc_handle_out, &h, sizeof(h));
    }
    // Load every kernel of the exchange path now: with lazy module loading the first launch of a kernel
    // can wait for running kernels to drain -- and a consumer spinning on a peer's flag never drains if
    // that peer (another stream of this process) is the one stuck loading.
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_raise_flags));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_gather));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_stretch_propose));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_stretch_accept));
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_chain_record));
    if (n_ranks == 1) {
        unsigned long long p = (unsigned long long)ex->region;
        CUDA_TRY(cudaMemcpy(ex->d_peer_ptrs, &p, sizeof(p), cudaMemcpyHostToDevice));
        ex->connected = true;
    }
    return SPAMM_OK;
}

extern "C" int spamm_exchange_create(int32_t n_ranks, int32_t rank, int64_t max_walkers, SpammExchange** out,
                                     void* ipc_handle_out) {
    if (!out) return set_error(SPAMM_EINVAL, "null argument");
    SpammExchange* ex = new SpammExchange();
    int rc = exchange_create_impl(ex, n_ranks, rank, max_walkers, ipc_handle_out);
    if (rc) { spamm_exchange_destroy(ex); *out = nullptr; return rc; }
    *out = ex;
    return SPAMM_OK;
}

static int exchange_finish_connect(SpammExchange* ex) {
    std::vector<unsigned long long> p(ex->n_ranks);
    for (int r = 0; r < ex->n_ranks; ++r) p[r] = (unsigned long long)ex->peer[r];
    CUDA_TRY(cudaMemcpy(ex->d_peer_ptrs, p.data(), p.size() * sizeof(p[0]), cudaMemcpyHostToDevice));
    ex->connected = true;
    return SPAMM_OK;
}

extern "C" int spamm_exchange_connect(SpammExchange* ex, const void* all_handles) {
    if (!ex || !all_handles) return set_error(SPAMM_EINVAL, "null argument");
    const cudaIpcMemHandle_t* hs = reinterpret_cast<const cudaIpcMemHandle_t*>(all_handles);
    for (int r = 0; r < ex->n_ranks; ++r) {
        if (r == ex->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + r, sizeof(h));
        void* p = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ex->peer[r] = p;
        ex->opened[r] = true;
    }
    return exchange_finish_connect(ex);
}

extern "C" int spamm_exchange_connect_ptrs(SpammExchange* ex, const uint64_t* region_ptrs) {
    if (!ex || !region_ptrs) return set_error(SPAMM_EINVAL, "null argument");
    for (int r = 0; r < ex->n_ranks; ++r)
        if (r != ex->rank) ex->peer[r] = reinterpret_cast<void*>(region_ptrs[r]);
    return exchange_finish_connect(ex);
}

extern "C" uint64_t spamm_exchange_region_ptr(const SpammExchange* ex) { return ex ? (uint64_t)ex->region : 0; }
extern "C" uint64_t spamm_exchange_epoch(const SpammExchange* ex) { return ex ? ex->epoch : 0; }

// queue one sharded evaluation of params_dev[n_total][D]: this rank's slice through k_walkers, results into this
// rank's region, flag raised in every region; *pi_out tells a consumer where the gathered vector lives
static int sharded_eval(SpammPlan* pl, SpammExchange* ex, const double* params_dev, int32_t n_total,
                        cudaStream_t st, PeerIn* pi_out, Proposal pr = Proposal{nullptr, 0, 0, 0, 0.0, 0ull, 0ull}) {
    if (!ex->connected) return set_error(SPAMM_EINVAL, "exchange is not connected");
    if (n_total > ex->max_n) return set_error(SPAMM_EINVAL, "exchange region holds fewer walkers than n_total");
    const uint64_t e = ++ex->epoch;
    int64_t lo, hi;
    spamm_shard_bounds(n_total, ex->n_ranks, ex->rank, &lo, &hi);
    const long long boff = kFlagWords + (long long)(e & 1) * ex->max_n;
    PeerOut po{ex->d_peer_ptrs, ex->n_ranks, boff + lo, ex->d_done, (unsigned)(hi - lo), ex->rank, e};
    if (hi > lo) {
        pr.i0 = (int)lo;                 // (proposal mode: walker w of the launch is proposal lo + w of the half)
        int rc = launch_walkers<0>(pl, full_mask(pl), params_dev + (size_t)lo * pl->dev.D, (int)(hi - lo), nullptr, st,
                                   false, po, pr);
        if (rc) return rc;
    } else {
        k_raise_flags<<<1, 1, 0, st>>>(po);
        pl->launches += 1;
        CUDA_TRY(cudaGetLastError());
    }
    *pi_out = PeerIn{ex->d_peer_ptrs, ex->n_ranks, ex->rank, boff, (int)((n_total + ex->n_ranks - 1) / ex->n_ranks), e,
                     pl->dev.status, pl->dev.status_host};
    return SPAMM_OK;
}

// wait for every rank's flag and pull the other ranks' slices into this rank's region (and into `copy`)
static int queue_gather(SpammPlan* pl, SpammExchange* ex, const PeerIn& pi, int32_t n_total, double* copy,
                        cudaStream_t st) {
    // programmatic stream serialisation: the gather kernel may start while the walker kernel before it drains
    // (it synchronises through the exchange flags, never through the grid dependency)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((n_total + 255) / 256), 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = ex->pdl ? 1 : 0;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, k_gather, pi, n_total, copy));
    pl->launches += 1;
    return SPAMM_OK;
}

extern "C" int spamm_lnpost_batch_sharded(SpammPlan* pl, SpammExchange* ex, const double* params_dev,
                                          int32_t n_total, double* lnp_dev, const double** lnp_region_out,
                                          void* stream) {
    if (!pl || !ex || !params_dev) return set_error(SPAMM_EINVAL, "null argument");
    if (n_total < 1) return set_error(SPAMM_EINVAL, "need at least one walker");
    if (!pl->has_data) return set_error(SPAMM_EINVAL, "plan was created without data flux/flux_error");
    cudaStream_t st = (cudaStream_t)stream;
    PeerIn pi;
    int rc = deferred_status(pl, st);
    if (rc) return rc;
    if ((rc = sharded_eval(pl, ex, params_dev, n_total, st, &pi))) return rc;
    if ((rc = queue_gather(pl, ex, pi, n_total, lnp_dev, st))) return rc;
    if (lnp_region_out) *lnp_region_out = reinterpret_cast<const double*>(ex->region) + pi.boff;
    return SPAMM_OK;
}

extern "C" int spamm_lnpost_batch_sharded_host(SpammPlan* pl, SpammExchange* ex, const double* params_host,
                                               int32_t n_total, double* lnp_host) {
    if (!pl || !ex || !params_host || !lnp_host) return set_error(SPAMM_EINVAL, "null argument");
    if (n_total < 1 || n_total > ex->max_n) return set_error(SPAMM_EINVAL, "walker count outside the exchange region");
    if (!pl->has_data) return set_error(SPAMM_EINVAL, "plan was created without data flux/flux_error");
    const int D = pl->dev.D;
    if (!ex->own_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ex->own_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaMallocHost(&ex->h_params, (size_t)ex->max_n * D * sizeof(double)));
        CUDA_TRY(cudaMallocHost(&ex->h_lnp, (size_t)ex->max_n * sizeof(double)));
        CUDA_TRY(cudaMalloc(&ex->d_params, (size_t)ex->max_n * D * sizeof(double)));
    }
    cudaStream_t st = ex->own_stream;
    auto is_pinned = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    // only this rank's slice of the parameters travels to the device; the whole result vector travels back
    int64_t lo, hi;
    spamm_shard_bounds(n_total, ex->n_ranks, ex->rank, &lo, &hi);
    if (hi > lo) {
        const size_t nb = (size_t)(hi - lo) * D * sizeof(double);
        const double* src = params_host + (size_t)lo * D;
        if (!is_pinned(params_host)) { memcpy(ex->h_params, src, nb); src = ex->h_params; }
        CUDA_TRY(cudaMemcpyAsync(ex->d_params + (size_t)lo * D, src, nb, cudaMemcpyHostToDevice, st));
    }
    PeerIn pi;
    int rc = sharded_eval(pl, ex, ex->d_params, n_total, st, &pi);
    if (rc) return rc;
    if ((rc = queue_gather(pl, ex, pi, n_total, nullptr, st))) return rc;
    const bool pin_out = is_pinned(lnp_host);
    double* dst = pin_out ? lnp_host : ex->h_lnp;
    CUDA_TRY(cudaMemcpyAsync(dst, reinterpret_cast<const double*>(ex->region) + pi.boff,
                             (size_t)n_total * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (!pin_out) memcpy(lnp_host, ex->h_lnp, (size_t)n_total * sizeof(double));
    return check_status(pl, st);
}

// spamm_stretch_run with the walkers of every half-step sharded over the ranks of an exchange: every rank
// holds the replicated state and the same counter-based RNG, proposes the whole half-ensemble, evaluates its
// slice (results stored into every rank's region by the kernel) and accepts after the flag wait fused into
// k_stretch_accept -- no host code and no collective per half-step.
extern "C" int spamm_stretch_run_sharded(SpammPlan* pl, SpammExchange* ex, double* walkers_dev, double* lnp_dev,
                                         int32_t K, double a, uint64_t seed, uint64_t iter0, int32_t n_iter,
                                         double* q_dev, double* z_dev, int64_t* naccepted_dev, double* chain_dev,
                                         double* lnp_chain_dev, int64_t chain_len, void* stream) {
    if (!pl || !ex || !walkers_dev || !lnp_dev || !q_dev || !z_dev) return set_error(SPAMM_EINVAL, "null argument");
    if (K < 2 || (K & 1) || n_iter < 0) return set_error(SPAMM_EINVAL, "need an even walker count");
    if (!pl->has_data) return set_error(SPAMM_EINVAL, "plan was created without data flux/flux_error");
    if ((chain_dev || lnp_chain_dev) && (int64_t)iter0 + n_iter > chain_len)
        return set_error(SPAMM_EINVAL, "chain buffer too short");
    const int D = pl->dev.D, Ns = K / 2;
    cudaStream_t st = (cudaStream_t)stream;
    const bool fused = pl->template_mode != SPAMM_TEMPLATES_DIRECT;      // see spamm_stretch_run
    for (int it = 0; it < n_iter; ++it) {
        const uint64_t iteration = iter0 + (uint64_t)it;
        for (int half = 0; half < 2; ++half) {
            int rc;
            if (!fused && (rc = spamm_stretch_propose(walkers_dev, K, D, half, a, seed, iteration, q_dev, z_dev, stream)))
                return rc;
            PeerIn pi;
            Proposal pr{fused ? walkers_dev : nullptr, K, half, 0, a, seed, iteration};
            if ((rc = sharded_eval(pl, ex, fused ? walkers_dev : q_dev, Ns, st, &pi, pr))) return rc;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)((Ns + 127) / 128), 1, 1);
            cfg.blockDim = dim3(128, 1, 1);
            cfg.stream = st;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr.val.programmaticStreamSerializationAllowed = ex->pdl ? 1 : 0;
            cfg.attrs = &attr;
            cfg.numAttrs = 1;
            StretchFused fu{fused ? 1 : 0, a, chain_dev, lnp_chain_dev, (long long)chain_len, (long long)iteration};
            CUDA_TRY(cudaLaunchKernelEx(&cfg, k_stretch_accept, walkers_dev, lnp_dev, (int)K, D, half,
                                        (const double*)q_dev, (const double*)z_dev, (const double*)nullptr, seed,
                                        iteration, reinterpret_cast<long long*>(naccepted_dev), pi, fu));
        }
        if (!fused && (chain_dev || lnp_chain_dev)) {
            int rc = spamm_chain_record(walkers_dev, lnp_dev, K, D, (int64_t)iteration, chain_len, chain_dev,
                                        lnp_chain_dev, stream);
            if (rc) return rc;
        }
    }
    return SPAMM_OK;
}

#ifdef SPAMM_TRACE
extern "C" int spamm_debug_set_trace(unsigned long long* buf) {
    CUDA_TRY(cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)));
    return SPAMM_OK;
}
#endif

// Timing probe of the broadening kernels on their own (roofline evidence for the per-walker DIRECT path):
// n_rows rows of family 0 (Fe) / 1 (host), templates cycled, every row with kernel width sigma_norm, run
// `repeats` times; returns the mean duration of k_conv_rows and of k_interp_rows + k_norm_rows (CUDA events)
// and the FMA count of one k_conv_rows launch (N_t * M per row: the SURVEY 8(d) convolution term / 2).
extern "C" int spamm_broadening_probe(SpammPlan* pl, int32_t family, int32_t sigma_norm, int32_t n_rows,
                                      int32_t repeats, double* conv_ms_out, double* interp_ms_out,
                                      double* conv_fma_out) {
    if (!pl || n_rows < 1 || repeats < 1 || sigma_norm < 1) return set_error(SPAMM_EINVAL, "bad argument");
    const TemplDev& ts = family == 0 ? pl->dev.fe : pl->dev.host;
    if (ts.n_templ < 1 || !ts.logf) return set_error(SPAMM_EINVAL, "plan has no such template family");
    std::vector<int> rt(n_rows), rs(n_rows, sigma_norm);
    int stride = 0;
    double fma_count = 0.0;
    for (int t = 0; t < ts.n_templ; ++t) stride = std::max(stride, ts.n_pts[t]);
    for (int r = 0; r < n_rows; ++r) {
        rt[r] = r % ts.n_templ;
        fma_count += (double)ts.n_pts[rt[r]] * ts.ksize * sigma_norm;
    }
    int *d_rt = nullptr, *d_rs = nullptr;
    double *d_conv = nullptr, *d_f = nullptr, *d_nf = nullptr;
    CUDA_TRY(cudaMalloc(&d_rt, n_rows * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_rs, n_rows * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_conv, (size_t)n_rows * stride * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_f, ((size_t)n_rows * pl->dev.P + kPad) * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_nf, (size_t)n_rows * sizeof(double)));
    CUDA_TRY(cudaMemcpy(d_rt, rt.data(), n_rows * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_rs, rs.data(), n_rows * sizeof(int), cudaMemcpyHostToDevice));
    const int smem = conv_smem_bytes(ts.ksize, sigma_norm);
    int rc = SPAMM_OK;
    if (smem > 200 * 1024) rc = set_error(SPAMM_ERANGE, "broadening kernel too wide for shared memory");
    if (!rc && smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_conv_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        rc = set_error(SPAMM_ECUDA, "cudaFuncSetAttribute(k_conv_rows)");
    if (!rc) {
        int nmax = stride;
        dim3 g1((nmax + kConvTile - 1) / kConvTile, n_rows), g2((pl->dev.P + 255) / 256, n_rows);
        cudaEvent_t e0, e1, e2;
        cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
        double tc = 0.0, ti = 0.0;
        for (int it = 0; it <= repeats; ++it) {          // iteration 0 warms up
            cudaEventRecord(e0);
            k_conv_rows<<<g1, kConvThreads, smem>>>(ts, d_rt, d_rs, d_conv, stride);
            cudaEventRecord(e1);
            k_interp_rows<<<g2, 256>>>(ts, d_rt, d_conv, stride, pl->dev.lnwl, pl->dev.P, d_f);
            k_norm_rows<<<(n_rows + 127) / 128, 128>>>(ts, d_rt, n_rows, pl->dev.wl, pl->dev.P, d_f, d_nf);
            cudaEventRecord(e2);
            cudaEventSynchronize(e2);
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, e0, e1);
            cudaEventElapsedTime(&b, e1, e2);
            if (it) { tc += a; ti += b; }
        }
        pl->launches += 3 * (repeats + 1);
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        if (cudaGetLastError() != cudaSuccess) rc = set_error(SPAMM_ECUDA, "broadening probe launch failed");
        if (conv_ms_out) *conv_ms_out = tc / repeats;
        if (interp_ms_out) *interp_ms_out = ti / repeats;
        if (conv_fma_out) *conv_fma_out = fma_count;
    }
    cudaFree(d_rt); cudaFree(d_rs); cudaFree(d_conv); cudaFree(d_f); cudaFree(d_nf);
    return rc;
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
