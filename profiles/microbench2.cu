// Chain-latency micro-benchmark of the exchange owner loop variants (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 1024
__device__ __forceinline__ void sts_v4(int4* p, int4 v) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.volatile.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <int VAR>
__global__ void k_chain(double* out, long long* cyc, const double* kap, const double* inv) {
    __shared__ double sk[ITERS + 64];
    __shared__ double si[ITERS + 64];
    __shared__ int4 sw[ITERS + 64];
    for (int i = threadIdx.x; i < ITERS + 64; i += blockDim.x) { sk[i] = kap[i % 64]; si[i] = inv[i % 64]; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double A = 1.0 + lane * 1e-3, om = 1.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int st = 0; st < ITERS; ++st) {
        const int lane_o = st & 31;
        if (VAR >= 1) { if (lane >= lane_o) A = fma(sk[st], om, A); } else { A = fma(0.999, om, A); }
        double v = (VAR >= 1) ? A * si[st + 1] : A * 0.5;
        double nxt = __shfl_sync(0xffffffffu, v, lane_o);
        if (VAR >= 2) { if (!(nxt > 0x1.0p-400 && nxt < 0x1.0p400)) break; }
        om = nxt;
        if (VAR >= 3) {
            if (lane == lane_o) {
                int hi = __double2hiint(nxt);
                int ex = ((hi >> 20) & 0x7ff) - 1023;
                double m = __hiloint2double((hi & 0x800fffff) | (1023 << 20), __double2loint(nxt));
                sts_v4(&sw[st + 1], make_int4(__double2loint(m), __double2hiint(m), ex, st + 2));
            }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = A + om;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *out, *kap, *inv; long long* cyc; long long h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&kap, 64 * 8); cudaMalloc(&inv, 64 * 8);
    double hk[64], hi[64];
    for (int i = 0; i < 64; ++i) { hk[i] = 0.3 + 0.001 * i; hi[i] = 0.7; }
    cudaMemcpy(kap, hk, sizeof hk, cudaMemcpyHostToDevice); cudaMemcpy(inv, hi, sizeof hi, cudaMemcpyHostToDevice);
#define RUN(V, T) k_chain<V><<<1, T>>>(out, cyc, kap, inv); k_chain<V><<<1, T>>>(out, cyc, kap, inv); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("variant %d threads %4d: %.1f cycles/step\n", V, T, (double)h / ITERS);
    RUN(0, 32) RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(3, 512)
    return 0;
}
