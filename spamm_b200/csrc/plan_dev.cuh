// spamm_b200 -- device-side view of a plan: constants, PlanDev, template and cluster tables
// (part of the single translation unit spamm_b200.cu; reference call sites are cited as
// file:line relative to the reference root)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "spamm_b200.h"
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
constexpr int kConvR = 8, kConvThreads = 128;          // broadening kernel: outputs per thread, threads per CTA
constexpr int kConvTile = kConvR * kConvThreads;       // output points per CTA
// words per residue class of the transposed template tile (tile + halo of Mp - 1 points, Mp = padded taps)
__host__ __device__ inline int conv_group_stride(int Mp) { return (kConvTile + Mp - 1 + kConvR - 1) / kConvR + 1; }

enum StatusBits : int {
    kStatSigmaRange = 1,   // sigma_norm outside the bank / < 1
    kStatBcKernel = 2,     // log_conv kernel wider than one pixel (unsupported)
    kStatBcStage = 4,      // BC edge pixel beyond the staged f0 range
    kStatPeerTimeout = 8,  // multi-GPU exchange: a peer's flag did not arrive within 5 s
};
constexpr int kFlagWords = 64;         // flag block at the start of an exchange region (one word per rank)

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
    // rebin_spec: True -- the flux-conserving rebin onto the data grid as a CSR matrix per template
    // (rb_ptr [n_templ][P + 1], template-local columns); null: np.interp
    const int* rb_ptr;
    const int* rb_idx;
    const double* rb_w;
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
    // rebin_spec: True -- log_conv's two rebins composed into one sparse operator on f0 (CSR); null: the stencil
    const int* bc_rb_ptr;
    const int* bc_rb_idx;
    const double* bc_rb_w;
    const int4* bc_idx;                   // 4 source pixels of the log_conv stencil
    const double *bc_tb0, *bc_tb1, *bc_ta;
    int edge_guess;                       // insertion point of balmer_edge in wl
    int bc_nstage;
    // walker parts (k_walkers, kSpecParts): chunks [part_chunk, n_chunks) hold every pixel beyond the Balmer
    // edge of any in-prior walker; a red part stages f0 from part_stage0 (even) on; part_chunk < 0: not available
    int part_chunk, part_stage0;
    // adjacent staged pixels 2j, 2j+1: the second sample's two exponentials follow from the first through
    // expm1 of the (small) argument differences; bc_pair_ok: tau_max |t3[2j+1] - t3[2j]| <= 0.004 everywhere,
    // bc_dhnu_max: max |h nu[2j+1] - h nu[2j]| (the kernel needs bc_dhnu_max / kT <= 0.01 for the walker)
    int bc_pair_ok;
    double bc_dhnu_max;
    double bc_dlnew;
    double c_kms, c_cgs, c_aa, c_cgs2, kb, edge;
    // Extinction (ReddeningLaw.py): k(lambda) of the selected curve, E(B-V) parameter offset; null = absent
    const double* ext_k;
    int ext_off, ext_q;
    int fe_rows, host_rows;               // rows in the sigma banks
    int exp_safe;                         // the prior box bounds every exp() argument of the fused kernel
    int nc_pair_ok;                       // P even and |slope| |ln(wl[i+1]/wl[i])| <= 0.016 inside the prior box
    int* status;                          // sticky device status word (StatusBits)
    int* status_host;                     // mapped pinned host word: set non-zero together with `status`, so the
                                          // host notices without queueing a copy behind every kernel
};

__device__ __forceinline__ void flag_status(int* status, int* status_host, int bit) {
    atomicOr(status, bit);
    *reinterpret_cast<volatile int*>(status_host) = 1;
}
