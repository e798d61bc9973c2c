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

// Vector form for the hot pair loop: D multiplications + roundings and ONE rarely taken branch.
// rint(dx/L) equals floor(dx/L + 0.5) except at exact ties (|dx| an odd multiple of L/2); ties -- and anything within
// 1e-9 of one -- take the out-of-line exact path that evaluates the reference's expression literally.
// `negate`: the reference would have formed the separation with the opposite sign (x_lower - x_higher on a diagonal
// tile); that only matters at a tie, where mi(+L/2) = mi(-L/2) = -L/2, so it is handled in the exact path only.
static __device__ __noinline__ double min_image_exact(double d, double L, bool negate) {
    if (negate) d = -d;
    d -= L * floor(d / L + 0.5);
    return negate ? -d : d;
}
template <int D>
__device__ __forceinline__ void min_image_vec(double (&d)[D], double L, double invL, bool negate = false) {
    double n[D], worst = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const double q = d[c] * invL;
        n[c] = rint(q);
        worst = fmax(worst, fabs(q - n[c]));
    }
    if (worst > 0.5 - 1e-9) {
#pragma unroll
        for (int c = 0; c < D; ++c) d[c] = min_image_exact(d[c], L, negate);
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) d[c] = fma(-L, n[c], d[c]);
    }
}

// exp(t) for t <= 0 in the hot loops: no range branches, constants from the constant bank.
// t is clamped at -700 (e^-700 ~ 1e-304 is far below anything it is added to); n = rint(t log2 e),
// r = t - n ln2 (two FMAs, |r| <= 0.3466), degree-12 Taylor polynomial (remainder 1.7e-16), scaling by 2^n through
// the exponent field (n >= -1010 keeps the result normal).
static __constant__ double c_exp_poly[13] = {
    1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880, 1.0 / 3628800,
    1.0 / 39916800, 1.0 / 479001600};
__device__ __forceinline__ double exp_neg_fast(double t) {
    t = fmax(t, -700.0);
    const double n = rint(t * 1.4426950408889634);
    double r = fma(n, -6.93147180369123816490e-01, t);       // ln2 high part (fdlibm split)
    r = fma(n, -1.90821492927058770002e-10, r);              // ln2 low part
    // Estrin evaluation: the pair loop is bound by dependent-instruction latency (8.3 cycles per DFMA), so the
    // 12-deep Horner chain is folded into a depth-5 tree (same coefficients, 4 more multiplications)
    const double* c = c_exp_poly;
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a01 = fma(c[1], r, c[0]), a23 = fma(c[3], r, c[2]), a45 = fma(c[5], r, c[4]), a67 = fma(c[7], r, c[6]);
    const double a89 = fma(c[9], r, c[8]), aab = fma(c[11], r, c[10]);
    const double b0 = fma(a23, r2, a01), b1 = fma(a67, r2, a45), b2 = fma(aab, r2, a89);
    const double lo = fma(b1, r4, b0), hi = fma(c[12], r4, b2);
    const double p = fma(hi, r8, lo);
    const int ni = __double2int_rn(n);
    return __hiloint2double(__double2hiint(p) + (ni << 20), __double2loint(p));
}

// 1/sqrt(x) for x comfortably inside the float range: single-precision seed + two Newton steps in double
// (22 -> 44 -> 88 bits), no special-case branch.
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y = (double)rsqrtf((float)x);
    const double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    return y;
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
