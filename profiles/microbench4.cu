// Does an FP64 instruction cost the scheduler one issue slot or two? (not part of the product)
// 16 warps on one SM (4 per scheduler), 4 independent DFMA chains per thread, with 0 / 1 / 2 / 3 independent integer
// instructions (LOP3/IADD chains) interleaved per DFMA. If integer work rides in the shadow of the half-rate FP64 pipe the time
// stays flat until the issue port is full (1 DFMA + 1 INT per 2 cycles); if a DFMA takes both slots it grows at once.
// Same with FRND.F64 / MUFU.RSQ64H streams beside the DFMAs (do they share the FP64 pipe or run on the XU pipe in parallel?).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench4 profiles/microbench4.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 1024

template <int NI>
__global__ void k_mix(double* out, long long* cyc, double a, double b, unsigned m) {
    double x[4];
    unsigned u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { x[k] = threadIdx.x + k; u[k] = threadIdx.x * 7 + k; }
    long long t0 = clock64();
#pragma unroll 2
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            x[k] = fma(x[k], a, b);
#pragma unroll
            for (int j = 0; j < NI; ++j) u[(k + j) & 3] = (u[(k + j) & 3] ^ m) + (u[(k + j + 1) & 3] | 1u);   // LOP3 + IADD3-ish
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += x[k] + (double)u[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int KIND>   // 0: DFMA only, 1: + rint per 2 DFMA, 2: + rsqrt seed per 4 DFMA
__global__ void k_xu(double* out, long long* cyc, double a, double b) {
    double x[4], y[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = threadIdx.x + k;
    y[0] = 1.25 + threadIdx.x; y[1] = 2.5 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 2
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = fma(x[k], a, b);
        if (KIND == 1) { y[0] = rint(y[0]) ; y[1] = rint(y[1]); asm volatile("" : "+d"(y[0]), "+d"(y[1])); }
        if (KIND == 2) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y[0])); y[0] = r; }
    }
    long long t1 = clock64();
    double s = y[0] + y[1];
#pragma unroll
    for (int k = 0; k < 4; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
#define RUN(label, kern, threads, ...) \
    kern<<<1, threads>>>(__VA_ARGS__); kern<<<1, threads>>>(__VA_ARGS__); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-64s %7.2f cycles per iteration (4 DFMA per thread)\n", label, (double)h / ITERS);
    for (int th : {512, 768}) {
        printf("-- %d warps on one SM\n", th / 32);
        RUN("4 DFMA", k_mix<0>, th, out, cyc, 0.999, 1e-9, 0x55u)
        RUN("4 DFMA + 4 x 1 int pair", k_mix<1>, th, out, cyc, 0.999, 1e-9, 0x55u)
        RUN("4 DFMA + 4 x 2 int pairs", k_mix<2>, th, out, cyc, 0.999, 1e-9, 0x55u)
        RUN("4 DFMA + 4 x 3 int pairs", k_mix<3>, th, out, cyc, 0.999, 1e-9, 0x55u)
        RUN("4 DFMA (xu kernel)", k_xu<0>, th, out, cyc, 0.999, 1e-9)
        RUN("4 DFMA + 2 FRND.F64", k_xu<1>, th, out, cyc, 0.999, 1e-9)
        RUN("4 DFMA + 1 MUFU.RSQ64H", k_xu<2>, th, out, cyc, 0.999, 1e-9)
    }
    return 0;
}
