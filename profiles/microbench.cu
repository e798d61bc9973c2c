// Latency micro-benchmarks used to design the exchange recurrence chain (not part of the product).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench profiles/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < ITERS; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dmul(double* out, long long* cyc, double a) {
    double x = threadIdx.x + 1.0;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < ITERS; ++i) x = x * a;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_iadd(int* out, long long* cyc, int a) {
    int x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < ITERS; ++i) x = max(x + a, x ^ a);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* cyc) {
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < ITERS; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc) {
    __shared__ double s[64];
    s[threadIdx.x & 63] = (double)((threadIdx.x + 1) & 63);
    __syncthreads();
    int idx = threadIdx.x & 63;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < ITERS; ++i) idx = (int)s[idx] & 63;
    long long t1 = clock64();
    out[threadIdx.x] = idx;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(double* out, long long* cyc) {
    long long t0 = clock64();
    for (int i = 0; i < ITERS; ++i) __syncthreads();
    long long t1 = clock64();
    out[threadIdx.x] = 0;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// smem publish / poll ping-pong between two warps of one block (one round trip = 2 hand-offs)
__global__ void k_pingpong(double* out, long long* cyc) {
    __shared__ volatile int flag[2];
    if (threadIdx.x == 0) { flag[0] = 0; flag[1] = 0; }
    __syncthreads();
    const int w = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int i = 1; i <= ITERS; ++i) {
        if (w == 0) {
            if ((threadIdx.x & 31) == 0) flag[0] = i;
            while (flag[1] < i) {}
        } else if (w == 1) {
            while (flag[0] < i) {}
            if ((threadIdx.x & 31) == 0) flag[1] = i;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = 0;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_membar(double* out, long long* cyc) {
    __shared__ volatile int s[32];
    long long t0 = clock64();
    for (int i = 0; i < ITERS; ++i) { s[threadIdx.x & 31] = i; __threadfence_block(); }
    long long t1 = clock64();
    out[threadIdx.x] = s[0];
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
#define RUN(name, threads, ...) \
    name<<<1, threads>>>(__VA_ARGS__); name<<<1, threads>>>(__VA_ARGS__); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-28s threads=%4d  %.1f cycles/op\n", #name, threads, (double)h / ITERS);
    RUN(k_dfma, 32, out, cyc, 0.999, 1e-9)
    RUN(k_dfma, 512, out, cyc, 0.999, 1e-9)
    RUN(k_dmul, 32, out, cyc, 0.9999)
    RUN(k_iadd, 32, (int*)out, cyc, 3)
    RUN(k_shfl, 32, out, cyc)
    RUN(k_lds, 32, out, cyc)
    RUN(k_bar, 64, out, cyc)
    RUN(k_bar, 512, out, cyc)
    RUN(k_bar, 1024, out, cyc)
    RUN(k_pingpong, 64, out, cyc)
    RUN(k_membar, 32, out, cyc)
    return 0;
}
