// Device-side helpers: minimum image, warp/block reductions, Philox4x32-10 + Box-Muller.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace pimdb {

constexpr unsigned kFullMask = 0xffffffffu;

// Minimum image, reference src/common.cpp:41-43: dx -= L*floor(dx/L + 0.5).
// The division is replaced by a multiplication; when the argument of floor() lands within 1e-9 of an
// integer (e.g. lattice sites exactly L/2 apart) the exact division is redone so the image choice is the
// reference's.
__device__ __forceinline__ double min_image(double dx, double L, double invL) {
    double q = fma(dx, invL, 0.5);
    double fl = floor(q);
    if (fabs(q - rint(q)) < 1e-9) fl = floor(dx / L + 0.5);
    return dx - L * fl;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}

// Block-wide sum of K values per thread; result valid in thread 0. smem: K * 32 doubles.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) smem[k * 32 + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double t = lane < nw ? smem[k * 32 + lane] : 0.0;
            v[k] = warp_sum(t);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11). Counter-based: the noise of a degree of freedom is a pure function
// of (seed, draw index, bead-or-mode, axis, particle), independent of launch geometry and bead sharding.
struct Philox4 { uint32_t v[4]; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// Two independent standard normals for the particle pair (2q, 2q+1) of row `row` (= bead-or-mode * D + axis)
// at thermostat half-step `draw`:
//   (r0,r1,r2,r3) = Philox4x32-10(counter = (q, row, draw_lo, draw_hi), key = (seed_lo, seed_hi))
//   u1 = ((r0 | r1<<32) >> 11) + 1) * 2^-53  in (0,1],   u2 = ((r2 | r3<<32) >> 11) * 2^-53  in [0,1)
//   z_even = sqrt(-2 ln u1) cos(2 pi u2),  z_odd = sqrt(-2 ln u1) sin(2 pi u2)
__device__ __forceinline__ void gaussian_pair(uint32_t q, uint32_t row, unsigned long long draw,
                                              unsigned long long seed, double& z0, double& z1) {
    Philox4 r = philox4x32_10(q, row, (uint32_t)draw, (uint32_t)(draw >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
    unsigned long long a = ((unsigned long long)r.v[1] << 32) | r.v[0];
    unsigned long long b = ((unsigned long long)r.v[3] << 32) | r.v[2];
    double u1 = (double)((a >> 11) + 1ull) * 0x1.0p-53;
    double u2 = (double)(b >> 11) * 0x1.0p-53;
    double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    z0 = rad * cs;
    z1 = rad * sn;
}

}  // namespace pimdb
