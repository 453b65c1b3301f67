// Probe: accuracy of rcp.approx.ftz.f64 (MUFU.RCP64H) and of Newton refinements,
// and DFMA/DMUL/DADD dependent-issue latency on this GPU.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/rcp_probe tools/rcp_probe.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ double seed(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }

__global__ void k(const double* x, int n, double* e0, double* e1, double* e2, double* e3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = x[i];
    double y = seed(v);
    double ex = 1.0 / v;
    e0[i] = fabs(y - ex) / ex;
    double e = fma(-v, y, 1.0);
    double yq = fma(y, e, y);                 // one quadratic step
    e1[i] = fabs(yq - ex) / ex;
    double yc = fma(y, fma(e, e, e), y);      // one cubic step
    e2[i] = fabs(yc - ex) / ex;
    double e_ = fma(-v, yq, 1.0);
    double yqq = fma(yq, e_, yq);             // two quadratic steps
    e3[i] = fabs(yqq - ex) / ex;
}

__global__ void lat(double* out, long long* cyc, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, 0.999999999, 1e-12);
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    const int n = 1 << 20;
    double* hx = new double[n];
    for (int i = 0; i < n; ++i) {
        double u = (i + 0.5) / n;
        hx[i] = (1.0 + u) * pow(10.0, (i % 127) - 20.0);   // mantissas over [1,2), exponents 1e-20..1e106
    }
    double *dx, *e[4];
    cudaMalloc(&dx, n * 8);
    cudaMemcpy(dx, hx, n * 8, cudaMemcpyHostToDevice);
    for (int j = 0; j < 4; ++j) cudaMalloc(&e[j], n * 8);
    k<<<n / 256, 256>>>(dx, n, e[0], e[1], e[2], e[3]);
    double* h = new double[n];
    const char* names[4] = {"seed", "1 quadratic", "1 cubic", "2 quadratic"};
    for (int j = 0; j < 4; ++j) {
        cudaMemcpy(h, e[j], n * 8, cudaMemcpyDeviceToHost);
        double mx = 0;
        for (int i = 0; i < n; ++i) mx = fmax(mx, h[i]);
        printf("%-12s max rel err %.3e (2^%.1f)\n", names[j], mx, log2(mx));
    }
    double* o; long long* c; cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 8);
    lat<<<1, 32>>>(o, c, 4096); lat<<<1, 32>>>(o, c, 4096);
    long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles\n", hc / 4096.0);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
