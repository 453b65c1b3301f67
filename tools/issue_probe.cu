// Probe: does a warp-wide FP64 instruction (16 FP64 lanes per SM sub-partition, two pipe cycles)
// also hold the sub-partition's ISSUE port for two cycles, or can another pipe's instruction
// issue in the second cycle?  Times independent DFMA chains, independent integer chains
// (ALU: LOP3/IADD3; FMA-pipe: IMAD) and their mixes at the fused kernel's occupancy
// (256 threads x 4 CTAs per SM = 8 warps per sub-partition).
//   additive  : t(mix) ~ t(fp64 alone) + t(int alone)   -> FP64 costs two issue slots
//   overlapped: t(mix) ~ max(t(fp64 alone), t(int alone))
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/issue_probe.bin tools/issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ND, int NA, int NM>
__global__ void __launch_bounds__(256, 4) k(double* out, unsigned* iout, int iters, double b, double c, unsigned m) {
    double a[8];
    unsigned x[8], y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = 1.0 + 1e-9 * (threadIdx.x + j);
        x[j] = threadIdx.x * 8 + j;
        y[j] = threadIdx.x * 16 + j;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {       // interleave the three kinds, 8 independent chains each
#pragma unroll
            for (int j = 0; j < ND; ++j) a[(r * ND + j) & 7] = fma(a[(r * ND + j) & 7], b, c);
#pragma unroll
            for (int j = 0; j < NA; ++j)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[(r * NA + j) & 7]) : "r"(m), "r"(i));
#pragma unroll
            for (int j = 0; j < NM; ++j)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[(r * NM + j) & 7]) : "r"(m), "r"(i));
        }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += a[j]; t += x[j] + y[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int ND, int NA, int NM>
static double run(double* o, unsigned* io, int sms, int clock_khz) {
    const int iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ND, NA, NM><<<sms * 4, 256>>>(o, io, iters, 0.999999999, 1e-12, 0x9e3779b9u);
    cudaEventRecord(e0);
    k<ND, NA, NM><<<sms * 4, 256>>>(o, io, iters, 0.999999999, 1e-12, 0x9e3779b9u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // cycles per (one warp, one group of ND + NA + NM instructions) per sub-partition:
    // 8 warps per sub-partition, 8 groups per iteration
    const double cyc = ms * 1e-3 * clock_khz * 1e3 / (iters * 8.0 * 8.0);
    printf("per group: %d DFMA + %d LOP3 + %d IMAD : %7.3f ms  %6.2f issue cycles per group (%.2f per instruction)\n",
           ND, NA, NM, ms, cyc, cyc / (ND + NA + NM));
    return cyc;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, khz);
    double* o; unsigned* io;
    cudaMalloc(&o, (size_t)p.multiProcessorCount * 4 * 256 * 8);
    cudaMalloc(&io, (size_t)p.multiProcessorCount * 4 * 256 * 4);
    const int s = p.multiProcessorCount;
    double d = run<8, 0, 0>(o, io, s, khz);
    double a = run<0, 8, 0>(o, io, s, khz);
    double m = run<0, 0, 8>(o, io, s, khz);
    double da = run<8, 8, 0>(o, io, s, khz);
    double dm = run<8, 0, 8>(o, io, s, khz);
    double dam = run<8, 4, 4>(o, io, s, khz);
    run<4, 8, 0>(o, io, s, khz);
    run<8, 12, 0>(o, io, s, khz);
    printf("DFMA+LOP3: mix %.2f vs additive %.2f vs overlapped %.2f\n", da, d + a, d > a ? d : a);
    printf("DFMA+IMAD: mix %.2f vs additive %.2f vs overlapped %.2f\n", dm, d + m, d > m ? d : m);
    printf("DFMA+4 LOP3+4 IMAD: mix %.2f vs additive %.2f\n", dam, d + 0.5 * a + 0.5 * m);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
