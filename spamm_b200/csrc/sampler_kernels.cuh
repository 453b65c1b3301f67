// spamm_b200 -- stretch-move sampler kernels and the FP64 peak probe
// (part of the single translation unit spamm_b200.cu; reference call sites are cited as
// file:line relative to the reference root)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
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

// One stretch proposal (emcee 2.2.1 _propose_stretch) for walker i of the active half: z = ((a-1) u + 1)^2 / a and
// the index of the complementary walker, from the Philox counter (seed, i, half, iteration).
__device__ __forceinline__ void stretch_draw(uint64_t seed, int i, int half, uint64_t iter, double a, int K,
                                             double& zz, int& partner) {
    uint32_t r[4];
    philox4x32((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)i, (uint32_t)half,
               (uint32_t)iter, (uint32_t)(iter >> 32) ^ 0x50524f50u, r);
    const double u = u53(r[0], r[1]);
    zz = ((a - 1.0) * u + 1.0);
    zz = zz * zz / a;
    const int Ns = K / 2, Nc = K - Ns;
    int rint_ = (int)(u53(r[2], r[3]) * Nc);
    if (rint_ >= Nc) rint_ = Nc - 1;
    partner = half == 0 ? Ns + rint_ : rint_;
}

__global__ void k_stretch_propose(const double* __restrict__ walkers, int K, int D, int half,
                                  double a, uint64_t seed, uint64_t iter, double* __restrict__ q,
                                  double* __restrict__ z) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int Ns = K / 2;
    if (i >= Ns) return;
    double zz;
    int partner;
    stretch_draw(seed, i, half, iter, a, K, zz, partner);
    const double* s = walkers + (size_t)(half == 0 ? i : Ns + i) * D;
    const double* cc = walkers + (size_t)partner * D;
    for (int d = 0; d < D; ++d) q[(size_t)i * D + d] = cc[d] - zz * (cc[d] - s[d]);
    z[i] = zz;
}

struct StretchFused {
    int on;                    // 0: q / z come from k_stretch_propose and the chain is recorded by k_chain_record
    double a;
    double* chain;             // [K][chain_len][D] or null
    double* lnp_chain;         // [K][chain_len] or null
    long long chain_len, chain_iter;
};

// pi.bufs != nullptr (multi-GPU): lnp_q is unused; every rank's walker kernel has left its slice of the
// proposals' ln-posteriors in its own exchange region.  Each block waits until all ranks have raised their
// flag for the epoch and a thread takes its walker's value from the owner's region (wait_flags, gathered_value).
__global__ void k_stretch_accept(double* __restrict__ walkers, double* __restrict__ lnp, int K,
                                 int D, int half, const double* __restrict__ q,
                                 const double* __restrict__ z, const double* lnp_q,
                                 uint64_t seed, uint64_t iter, long long* __restrict__ nacc, PeerIn pi,
                                 StretchFused fu) {
    if (pi.bufs != nullptr) wait_flags(pi);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int Ns = K / 2;
    if (i >= Ns) return;
    uint32_t r[4];
    philox4x32((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)i, (uint32_t)half,
               (uint32_t)iter, (uint32_t)(iter >> 32) ^ 0x41434350u, r);
    double u = u53(r[0], r[1]);
    int wi = half == 0 ? i : Ns + i;
    const double lq = pi.bufs != nullptr ? gathered_value(pi, i) : lnp_q[i];
    // fu.on: the library's own loop -- the proposal was made inside the walker kernel (k_walkers, Proposal) and is
    // re-made here from the same counter instead of being read from q / z (the complementary half does not move
    // during this half-step and this thread is the only writer of its own walker)
    double zz = 0.0;
    int partner = 0;
    if (fu.on) stretch_draw(seed, i, half, iter, fu.a, K, zz, partner);
    else zz = z[i];
    double lnpdiff = ((double)D - 1.0) * log(zz) + lq - lnp[wi];
    if (lnpdiff > log(u)) {
        if (fu.on) {
            const double* cc = walkers + (size_t)partner * D;
            double* sw = walkers + (size_t)wi * D;
            for (int d = 0; d < D; ++d) sw[d] = cc[d] - zz * (cc[d] - sw[d]);
        } else {
            for (int d = 0; d < D; ++d) walkers[(size_t)wi * D + d] = q[(size_t)i * D + d];
        }
        lnp[wi] = lq;
        if (nacc) nacc[wi] += 1;
    }
    // ... and this half's walkers are final for the iteration: record them (k_chain_record fused)
    if (fu.on) {
        if (fu.chain)
            for (int d = 0; d < D; ++d)
                fu.chain[((size_t)wi * fu.chain_len + fu.chain_iter) * D + d] = walkers[(size_t)wi * D + d];
        if (fu.lnp_chain) fu.lnp_chain[(size_t)wi * fu.chain_len + fu.chain_iter] = lnp[wi];
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
