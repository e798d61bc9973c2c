// Device-side helpers: minimum image, warp/block reductions, Philox4x32-10 + Box-Muller.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>

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

// Vector form for the hot pair loop: per axis one multiplication, one rounding and one FMA, plus ONE rarely taken
// branch decided by integer compares. rint(dx/L) equals floor(dx/L + 0.5) except at exact ties (|dx| an odd multiple
// of L/2); a result within 2^-19 (relative) of +-L/2 -- tested on the high words, `thr_hi` = high word of
// (L/2)(1 - 2^-19) -- takes the out-of-line exact path that evaluates the reference's expression literally.
// `negate`: the reference would have formed the separation with the opposite sign (x_lower - x_higher on a diagonal
// tile); that only matters at a tie, where mi(+L/2) = mi(-L/2) = -L/2, so it is handled in the exact path only.
static __device__ __noinline__ double min_image_exact(double d, double L, bool negate) {
    if (negate) d = -d;
    d -= L * floor(d / L + 0.5);
    return negate ? -d : d;
}
inline int min_image_tie_threshold(double L) {
    const double thr = 0.5 * L * (1.0 - 1.0 / 524288.0);
    unsigned long long u;
    memcpy(&u, &thr, 8);
    return (int)(u >> 32);
}
template <int D>
__device__ __forceinline__ void min_image_vec(double (&d)[D], double L, double invL, int thr_hi, bool negate = false) {
    double w[D];
    int worst = 0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        w[c] = fma(-L, rint(d[c] * invL), d[c]);
        worst = max(worst, __double2hiint(w[c]) & 0x7fffffff);
    }
    if (worst >= thr_hi) {
#pragma unroll
        for (int c = 0; c < D; ++c) d[c] = min_image_exact(d[c], L, negate);
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) d[c] = w[c];
    }
}

// exp() for non-positive arguments in the hot loops, no FP64 compares and no conversions, evaluated in base 2:
// n = round(t log2 e) by the 1.5*2^52 shift (the integer n is the low word of the shifted sum), f = t log2 e - n
// (|f| <= 1/2, FMAs against a two-part log2 e), 2^f by the degree-12 Taylor polynomial in f ln2 (remainder 1.7e-16)
// in even/odd Horner form, scaling by 2^n through the exponent field. Results below 2^-1020 return 0 (decided by an
// integer compare on the shifted sum, so -inf and huge arguments are safe; a NaN propagates).
// The constants whose low word is not zero travel as kernel parameters (ExpConsts inside the argument struct): the
// compiler then keeps them in uniform registers for the whole loop instead of re-materialising each one with two
// moves per use (that was a third of the pair loop's instructions).
struct ExpConsts {
    double l2e_hi, l2e_lo;
    double p[13];          // p[k] = ln2^k / k!   (p[0] = 1)
};
inline ExpConsts make_exp_consts() {
    ExpConsts k;
    k.l2e_hi = 1.4426950408889634;          // log2(e) rounded to double
    k.l2e_lo = 2.0355273740931033e-17;      // log2(e) - l2e_hi
    long double ln2 = 0.693147180559945309417232121458176568L, t = 1.0L;
    for (int i = 0; i <= 12; ++i) {
        k.p[i] = (double)t;
        t = t * ln2 / (long double)(i + 1);
    }
    return k;
}
constexpr double kExpShift = 6755399441055744.0;   // 1.5 * 2^52
// 2^f * 2^n from the shifted sum s = n + kExpShift and the reduced argument f
__device__ __forceinline__ double exp2_finish(double s, double f, const ExpConsts& k) {
    const double f2 = f * f;
    double pe = fma(k.p[12], f2, k.p[10]), po = fma(k.p[11], f2, k.p[9]);
    pe = fma(pe, f2, k.p[8]); po = fma(po, f2, k.p[7]);
    pe = fma(pe, f2, k.p[6]); po = fma(po, f2, k.p[5]);
    pe = fma(pe, f2, k.p[4]); po = fma(po, f2, k.p[3]);
    pe = fma(pe, f2, k.p[2]); po = fma(po, f2, k.p[1]);
    pe = fma(pe, f2, 1.0);
    const double p = fma(po, f, pe);
    // s >= kExpShift - 1020  <=>  n >= -1020 (positive doubles order like their bit patterns; -inf is negative)
    const bool under = __double_as_longlong(s) < 0x4337FFFFFFFFFC04LL;
    const int ni = __double2loint(s);
    return __hiloint2double(under ? 0 : __double2hiint(p) + (ni << 20), under ? 0 : __double2loint(p));
}
// exp(t), t <= 0
__device__ __forceinline__ double exp_neg_fast(double t, const ExpConsts& k) {
    const double s = fma(t, k.l2e_hi, kExpShift);
    const double n = s - kExpShift;
    double f = fma(t, k.l2e_hi, -n);
    f = fma(t, k.l2e_lo, f);
    return exp2_finish(s, f, k);
}
// exp(K2 x / log2 e) = 2^(K2 x) for K2 x <= 0 with the product formed inside the FMAs: two operations fewer on the
// pair loop. K2 = (rate) * log2(e) is rounded once, so the argument carries a relative error of 1.1e-16 -- the same
// size as the rounding of the reference's own product -alpha * x before it calls exp().
__device__ __forceinline__ double exp2_lin_fast(double x, double K2, const ExpConsts& k) {
    const double s = fma(x, K2, kExpShift);
    const double n = s - kExpShift;
    return exp2_finish(s, fma(x, K2, -n), k);
}

// 1/sqrt(x): the hardware's double-precision seed (MUFU.RSQ64H, ~2^-22) + one third-order step
// y (1 + e/2 + 3e^2/8), e = 1 - x y^2 (remainder ~ e^3/3 ~ 2^-67), no conversions and no special-case branch.
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

// In-kernel timeline (PIMDB_TIMELINE=1, profiling aid): every block stamps %globaltimer into its kernel's slot,
// [first start, last end] in ns -- the only way to see how the kernels of one captured step overlap on the device.
__device__ __forceinline__ unsigned long long gtimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tl_begin(unsigned long long* tl) {
    if (tl && threadIdx.x == 0) atomicMin(tl, gtimer_ns());
}
__device__ __forceinline__ void tl_end(unsigned long long* tl) {
    if (tl && threadIdx.x == 0) atomicMax(tl + 1, gtimer_ns());
}

// ---------------------------------------------------------------------------------------------------
// Peer-memory signalling (bead sharding, internal.cuh PeerMailbox): system-scope relaxed loads / stores of flag words
// that another GPU (or another process on this GPU) writes or reads, and BOUNDED waits on them. A wait that runs out
// raises `err_bit` in the host-mapped error word and a flag in device memory (`dead`), and returns false; once the flag
// is up every later wait returns at once, so a dead peer costs one time-out, not one per kernel of the remaining graph
// replays. (The flag that the waits consult lives in DEVICE memory: the error word is zero-copy host memory, and a read
// of it from every waiting thread of every block costs a PCIe round trip each -- 250 us per step at C3 when it was
// the one consulted.)
// (acquire: data read after a satisfied flag load is at least as new as the flag -- the peer released it with a
// system-scope fence before raising the flag. Cheaper than a fence.sys after the wait, which costs ~3.5 us here.)
__device__ __forceinline__ unsigned ld_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// wait until the counter at p has reached `want` (wrap-safe)
__device__ __forceinline__ bool wait_sys_u32_ge(const unsigned* p, unsigned want, unsigned long long timeout_ns, int* err,
                                                int err_bit, unsigned* dead) {
    if ((int)(ld_sys_u32(p) - want) >= 0) return true;
    if (*(volatile unsigned*)dead) return false;
    const unsigned long long t0 = gtimer_ns();
    unsigned spins = 0;
    while ((int)(ld_sys_u32(p) - want) < 0) {
        if ((++spins & 255u) == 0 && gtimer_ns() - t0 > timeout_ns) {
            *(volatile unsigned*)dead = 1u;
            atomicOr(err, err_bit);
            return false;
        }
    }
    return true;
}
// wait until the high half of the self-validating word at p equals `seq`; returns its low half
__device__ __forceinline__ unsigned wait_sys_word(const unsigned long long* p, unsigned seq, unsigned long long timeout_ns,
                                                  int* err, int err_bit, unsigned* dead) {
    unsigned long long w = ld_sys_u64(p);
    if ((unsigned)(w >> 32) == seq) return (unsigned)w;
    if (*(volatile unsigned*)dead) return 0u;
    const unsigned long long t0 = gtimer_ns();
    unsigned spins = 0;
    while ((unsigned)((w = ld_sys_u64(p)) >> 32) != seq) {
        if ((++spins & 255u) == 0 && gtimer_ns() - t0 > timeout_ns) {
            *(volatile unsigned*)dead = 1u;
            atomicOr(err, err_bit);
            return 0u;
        }
    }
    return (unsigned)w;
}

// Every block of a kernel that reads the halo slabs of a bead shard calls this first: the neighbours have raised the
// flags to the number of slices this rank itself has sent (all ranks push in lock-step program order).
__device__ __forceinline__ void peer_wait_halos(const unsigned int* halo_flag, const unsigned int* halo_seq,
                                                unsigned long long timeout_ns, int* err) {
    if (!halo_flag) return;
    if (threadIdx.x < 2) {
        wait_sys_u32_ge(&halo_flag[threadIdx.x], *halo_seq, timeout_ns, err, 8 /* kErrPeerTimeout */,
                        const_cast<unsigned*>(halo_seq) + 2 /* the handle's "peer is dead" flag, PeerDev::seq[3] */);
    }
    __syncthreads();
}

// Programmatic dependent launch (PTX griddepcontrol): the producer lets its dependents be scheduled early, the consumer
// blocks until the producer grid has completed and its writes are visible. A real dependency -- nothing to time out,
// correct under any serialisation of kernels (profilers, MPS, time slicing) -- that still lets the consumer's blocks take
// their SMs and run their prologue while the producer is running. No-ops without the launch attribute.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
