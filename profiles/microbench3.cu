// FP64 pipe micro-benchmarks behind the pair-kernel design (not part of the product): dependent-issue latency of DFMA /
// MUFU.RSQ64H / F2F-style rounding, and how many independent chains per scheduler saturate the FP64 pipe.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench3 profiles/microbench3.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048

template <int ILP>
__global__ void k_dfma_ilp(double* out, long long* cyc, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x + k;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rsq(double* out, long long* cyc) {
    double x = 1.5 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y + 1.5; }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rint(double* out, long long* cyc) {
    double x = 1.5 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) x = rint(x * 1.0000001) + 0.25;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds_rmw(double* out, long long* cyc) {   // shared-memory read-modify-write chain, as the reaction-force update
    __shared__ double s[96];
    s[threadIdx.x % 96] = 0.0;
    __syncwarp();
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITERS; ++i) { const int j = (threadIdx.x + i) & 31; s[j] = fma(1.0000001, s[j], 0.5); __syncwarp(); }
    long long t1 = clock64();
    out[threadIdx.x] = s[threadIdx.x & 31];
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
#define RUN(label, kern, blocks, threads, per_iter, ...) \
    kern<<<blocks, threads>>>(__VA_ARGS__); kern<<<blocks, threads>>>(__VA_ARGS__); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-52s %7.2f cycles per iteration, %6.2f per op\n", label, (double)h / ITERS, (double)h / ITERS / (per_iter));
    RUN("DFMA dependent chain, 1 warp", k_dfma_ilp<1>, 1, 32, 1, out, cyc, 0.999, 1e-9)
    RUN("DFMA 2 chains/warp, 1 warp", k_dfma_ilp<2>, 1, 32, 2, out, cyc, 0.999, 1e-9)
    RUN("DFMA 4 chains/warp, 1 warp", k_dfma_ilp<4>, 1, 32, 4, out, cyc, 0.999, 1e-9)
    RUN("DFMA 8 chains/warp, 1 warp", k_dfma_ilp<8>, 1, 32, 8, out, cyc, 0.999, 1e-9)
    RUN("DFMA 1 chain, 4 warps (1 per scheduler)", k_dfma_ilp<1>, 1, 128, 1, out, cyc, 0.999, 1e-9)
    RUN("DFMA 1 chain, 8 warps (2 per scheduler)", k_dfma_ilp<1>, 1, 256, 1, out, cyc, 0.999, 1e-9)
    RUN("DFMA 1 chain, 16 warps (4 per scheduler)", k_dfma_ilp<1>, 1, 512, 1, out, cyc, 0.999, 1e-9)
    RUN("DFMA 1 chain, 24 warps (6 per scheduler)", k_dfma_ilp<1>, 1, 768, 1, out, cyc, 0.999, 1e-9)
    RUN("DFMA 1 chain, 32 warps (8 per scheduler)", k_dfma_ilp<1>, 1, 1024, 1, out, cyc, 0.999, 1e-9)
    RUN("DFMA 2 chains, 16 warps", k_dfma_ilp<2>, 1, 512, 2, out, cyc, 0.999, 1e-9)
    RUN("DFMA 4 chains, 16 warps", k_dfma_ilp<4>, 1, 512, 4, out, cyc, 0.999, 1e-9)
    RUN("MUFU.RSQ64H + DADD chain, 1 warp", k_rsq, 1, 32, 1, out, cyc)
    RUN("rint (FRND.F64) + DMUL + DADD chain, 1 warp", k_rint, 1, 32, 1, out, cyc)
    RUN("shared RMW (LDS, DFMA, STS, syncwarp) chain, 1 warp", k_lds_rmw, 1, 32, 1, out, cyc)
    return 0;
}
