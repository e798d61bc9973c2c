// Bosonic exchange (Feldman-Hirshberg, O(N^2 + PN)) on the GPU: K4-K8 and the exchange estimators of K13.
//
// Reference: src/bosonic_exchange/quadratic_bosonic_exchange.cpp
//   evaluateCycleEnergies :34-61, evaluateVBn :73-99, evaluateVBackwards :101-128,
//   evaluateConnectionProbabilities :142-157, springForceLastBead :159-186, springForceFirstBead :188-215,
//   getDistinctProbability :222-229, getLongestProbability :238-240, primEstimator :250-279
// and src/bosonic_exchange/bosonic_exchange_base.cpp:30-64 (minimum-image bead separations).
//
// Design (DESIGN.md "exchange"):
//   1. Closed-form cycle energies. With d2(u,v) = |r^P_v - r^1_u|^2 (minimum image) and the prefix sum
//      A(w) = sum_{t<w} |r^1_{t+1} - r^P_t|^2 the reference's recurrence telescopes to
//          E^{[u..v]} = k/2 [ A(v) - A(u) + d2(u,v) ],      u <= v,
//      so E_kn (N(N+1)/2 doubles) is never stored.
//   2. All transcendental work is hoisted off the sequential chain. The Boltzmann factors
//          c(u,v) = exp(-beta E^{[u..v]})
//      depend on positions only; one fully parallel kernel evaluates them for all u <= v as *extended-range*
//      numbers (mantissa in [1,2) + 32-bit binary exponent), because beta*E easily exceeds the range of exp().
//   3. With W[m] = exp(-beta V[m]) and Wb[l] = exp(-beta Vb[l]) the two N-step recursions become linear
//      triangular recurrences
//          W[v+1] = 1/(v+1) sum_{j<=v} c(j,v) W[j],          Wb[l] = sum_{p>=l} c(l,p) Wb[p+1] / (p+1),
//      evaluated column-wise in extended-range arithmetic: when W[j] is known every row v >= j adds its term.
//      The dependency chain of a step is one multiply-add, one normalisation and one barrier -- no exp, no log,
//      no block reduction. Mathematically this is the reference's shifted log-sum-exp with the shift carried
//      exactly in the binary exponent, so it is robust for any positions the reference handles.
//      V[m] = -(ln W[m])/beta is recovered in parallel afterwards. Forward and backward recursions are
//      independent and run as two concurrent thread blocks; coefficient rows are prefetched with cp.async.
//   4. Connection probabilities are products of known extended-range numbers,
//          P(l->u) = W[u] c(u,l) Wb[l+1] / ((l+1) W[N]),
//      evaluated on the fly in the exterior-force kernel (one warp per particle and exterior bead); the N x N
//      matrix is only materialised when a caller asks for it (pimdb_exchange_get).
#include "internal.cuh"
#include "device_utils.cuh"

namespace pimdb {

// ---------------------------------------------------------------- extended-range positive numbers
struct Ext {
    double m;   // mantissa, in [1,2) when normalised (0 for an exact zero)
    int e;      // value = m * 2^e
};

constexpr int kExtZeroExp = -(1 << 29);
constexpr int kExchFastMaxN = 512;     // largest N served by the block-scaled tables and the fast recurrence

__device__ __forceinline__ double pow2i(int d) {   // 2^d for d in [-1022, 1023], 0 below
    return d < -1022 ? 0.0 : __hiloint2double((1023 + d) << 20, 0);
}

__device__ __forceinline__ Ext ext_normalize(double m, int e) {
    Ext r;
    if (!(m > 0.0)) { r.m = m; r.e = kExtZeroExp; return r; }   // zero (or NaN, which then propagates)
    const int hi = __double2hiint(m);
    const int ex = ((hi >> 20) & 0x7ff) - 1023;
    r.m = __hiloint2double((hi & 0x800fffff) | (1023 << 20), __double2loint(m));
    r.e = e + ex;
    return r;
}

// acc += a*b  (acc need not be normalised; its mantissa stays a moderate positive double)
__device__ __forceinline__ void ext_fma(double& am, int& ae, double m1, int e1, double m2, int e2) {
    const double tm = m1 * m2;
    const int te = e1 + e2;
    const int emax = max(ae, te);
    am = fma(am, pow2i(max(ae - emax, -1100)), tm * pow2i(max(te - emax, -1100)));
    ae = emax;
}

__device__ __forceinline__ double ext_to_double(double m, int e) {
    if (e < -1070) return 0.0;
    if (e < -1000) return (m * pow2i(e + 200)) * pow2i(-200);
    return m * pow2i(min(e, 1023));
}

// exp(-y) for y >= 0 of any magnitude, as an extended-range number. The reduction t = -y*log2(e) = n + r keeps r
// exact to ~2^-100 |y| by splitting log2(e) and using FMAs, so the relative error is that of exp2() itself.
__device__ __forceinline__ Ext ext_exp_neg(double y) {
    const double L2E_HI = 1.4426950408889634;          // log2(e) rounded to double
    const double L2E_LO = 2.0355273740931033e-17;      // log2(e) - L2E_HI
    double t = -y * L2E_HI;
    double n = rint(t);
    n = fmax(n, -1.0e9);
    double r = fma(-y, L2E_HI, -n);                    // exact product, one rounding
    r = fma(-y, L2E_LO, r);
    Ext o;
    o.m = exp2(r);                                     // in [2^-0.5, 2^0.5]
    o.e = (int)n;
    return ext_normalize(o.m, o.e);
}

__device__ __forceinline__ int4 ext_pack(double m, int e) { return make_int4(__double2loint(m), __double2hiint(m), e, 0); }
__device__ __forceinline__ double ext_m(const int4& c) { return __hiloint2double(c.y, c.x); }

// ---------------------------------------------------------------- arguments
struct ExArgs {
    const double *x1, *xP;     // bead 1 and bead P slices, [D][N]
    const double *x2, *xPm1;   // bead 2 (next of first) and bead P-1 (previous of last)
    double* A;                 // A[N] prefix sums
    double* Inv;               // Inv[i] = 1/i, i = 0..N (Inv[0] = 0): no FP64 division in the O(N^2) loops
    int4 *Cf, *Cb;             // Boltzmann factors, packed {mantissa lo, hi, binary exponent, 0}, N x N each:
                               //   Cf[j][v] = c(j,v) (v >= j);  Cb[p][l] = c(l,p) / (p+1) (l <= p, backward weight folded in)
    double *Kf, *Kb;           // the same factors block-scaled for the fast recurrence (N <= 512), or nullptr:
                               //   K[r][v] = C[r][v] * 2^-B[r/32][v] as a plain double (0 outside the triangle)
    int *Bf, *Bb;              //   B[rb][v] = largest binary exponent of C[32rb .. 32rb+31][v]
    double *Wm, *Wbm;          // W[0..N], Wb[0..N] mantissas
    int *We, *Wbe;             // ... exponents
    double *V, *Vb, *F;        // V[N+1], Vb[N+1], F[2][D][N]
    DevObs* obs;
    int* err;
    long long* dbg;            // optional profiling buffer (clock64 stamps per warp), nullptr in production
    int N, D, pbc, do_first, do_last;
    double k, beta, h, L, invL;   // h = beta*k/2
};

template <int D>
__device__ __forceinline__ double dist2(const ExArgs& a, const double* xa, int ia, const double* xb, int ib) {
    double r2 = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        double dx = xb[(size_t)c * a.N + ib] - xa[(size_t)c * a.N + ia];
        if (a.pbc) dx = min_image(dx, a.L, a.invL);
        r2 = fma(dx, dx, r2);
    }
    return r2;
}

// E^{[u..v]} for u <= v.
template <int D>
__device__ __forceinline__ double cycle_energy(const ExArgs& a, int u, int v) {
    return 0.5 * a.k * (a.A[v] - a.A[u] + dist2<D>(a, a.x1, u, a.xP, v));
}

// ---------------------------------------------------------------- 1. prefix sums A(w), one block
template <int D>
__global__ void __launch_bounds__(1024) k_exch_prefix(ExArgs a) {
    __shared__ double warp_tot[32];
    __shared__ double carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 32) warp_tot[tid] = 0.0;
    if (tid == 0) carry = 0.0;
    __syncthreads();
    // A[w] = sum_{t<w} link[t], link[t] = d2(P_t, 1_{t+1}); chunks of blockDim, inclusive scan per chunk
    for (int base = 0; base < a.N; base += blockDim.x) {
        const int w = base + tid;                      // produces A[w+1]
        double v = (w < a.N - 1) ? dist2<D>(a, a.xP, w, a.x1, w + 1) : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(kFullMask, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        if (warp == 0) {
            double t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                double u = __shfl_up_sync(kFullMask, t, o);
                if (lane >= o) t += u;
            }
            warp_tot[lane] = t;
        }
        __syncthreads();
        const double incl = carry + (warp > 0 ? warp_tot[warp - 1] : 0.0) + v;
        if (w + 1 < a.N) a.A[w + 1] = incl;
        __syncthreads();
        if (tid == blockDim.x - 1) carry = incl;
        __syncthreads();
    }
    if (tid == 0) a.A[0] = 0.0;
    for (int i = tid; i <= a.N; i += blockDim.x) a.Inv[i] = i > 0 ? 1.0 / (double)i : 0.0;
}

// ---------------------------------------------------------------- 2. Boltzmann factors, fully parallel
// element (r,s) of the N x N index square:  s >= r -> Cf[r][s] = c(r,s);  s <= r -> Cb[r][s] = c(s,r)
template <int D>
__global__ void __launch_bounds__(256) k_exch_coeff(ExArgs a) {
    const int N = a.N;
    const long long tot = (long long)N * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / N), s = (int)(i % N);
        const int u = min(r, s), v = max(r, s);
        const double y = a.h * (a.A[v] - a.A[u] + dist2<D>(a, a.x1, u, a.xP, v));
        const Ext c = ext_exp_neg(y < 0.0 ? 0.0 : y);   // (a NaN position stays NaN and is reported, like the reference)
        if (s >= r) a.Cf[i] = ext_pack(c.m, c.e);
        if (s <= r) {   // backward table: the 1/(p+1) weight of the sum (p = r) is folded in here, off the chain
            const Ext cb = ext_normalize(c.m * a.Inv[r + 1], c.e);
            a.Cb[i] = ext_pack(cb.m, cb.e);
        }
    }
}

// Tile version for N <= 512: one block per 32 x 32 tile of the index square also emits the block-scaled copy the
// fast recurrence consumes -- per (32-row block rb, column v) the largest binary exponent B and the factors as plain
// doubles relative to 2^B. A factor more than 2^1022 below its block's largest becomes 0; it multiplies values that
// the fast recurrence keeps within 2^+-400 of each other, so it could not have contributed.
template <int D>
__global__ void __launch_bounds__(1024) k_exch_coeff_tiles(ExArgs a) {
    __shared__ int s_ef[32][33], s_eb[32][33];     // [step within the tile][column], padded: conflict-free both ways
    __shared__ int s_mf[32], s_mb[32];
    __shared__ double sA[kExchFastMaxN + 1];       // the prefix sums A(w), recomputed by every tile (N <= 512: one chunk)
    __shared__ double warp_tot[32];
    const int N = a.N, nb = (N + 31) >> 5;
    const int rb = blockIdx.x / nb, sb = blockIdx.x % nb;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    {   // Same operations in the same order as k_exch_prefix (so A is bit-identical); fusing it here takes a 5 us
        // single-block kernel and a launch gap off the step's critical path. Block 0 also publishes A and 1/i.
        const int w = threadIdx.x;                     // produces A[w+1]
        double v = (w < N - 1) ? dist2<D>(a, a.xP, w, a.x1, w + 1) : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(kFullMask, v, o);
            if (tx >= o) v += t;
        }
        if (tx == 31) warp_tot[ty] = v;
        __syncthreads();
        if (ty == 0) {
            double t = warp_tot[tx];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                double u = __shfl_up_sync(kFullMask, t, o);
                if (tx >= o) t += u;
            }
            warp_tot[tx] = t;
        }
        __syncthreads();
        const double incl = 0.0 + (ty > 0 ? warp_tot[ty - 1] : 0.0) + v;
        if (w + 1 < N) sA[w + 1] = incl;
        if (w == 0) sA[0] = 0.0;
        if (blockIdx.x == 0) {
            if (w + 1 < N) a.A[w + 1] = incl;
            if (w == 0) a.A[0] = 0.0;
            if (w <= N) a.Inv[w] = w > 0 ? 1.0 / (double)w : 0.0;
        }
        __syncthreads();
    }
    const int r = rb * 32 + ty, sc = sb * 32 + tx;   // one element per thread: 4x the parallelism of a 256-thread tile
    double mf = 0.0, mb = 0.0;
    int ef = kExtZeroExp, eb = kExtZeroExp;
    const long long i = (long long)r * N + sc;
    if (r < N && sc < N) {
        const int u = min(r, sc), v = max(r, sc);
        const double y = a.h * (sA[v] - sA[u] + dist2<D>(a, a.x1, u, a.xP, v));
        const Ext c = ext_exp_neg(y < 0.0 ? 0.0 : y);   // (a NaN position stays NaN and is reported, like the reference)
        if (sc >= r) {
            a.Cf[i] = ext_pack(c.m, c.e);
            mf = c.m; ef = c.e;
        }
        if (sc <= r) {   // backward table: the 1/(p+1) weight of the sum (p = r) is folded in here, off the chain
            const Ext cb = ext_normalize(c.m * (1.0 / (double)(r + 1)), c.e);
            a.Cb[i] = ext_pack(cb.m, cb.e);
            mb = cb.m; eb = cb.e;
        }
    }
    s_ef[ty][tx] = ef;
    s_eb[ty][tx] = eb;
    __syncthreads();
    {   // warp ty reduces column ty over the 32 steps of the tile
        int xf = s_ef[tx][ty], xb = s_eb[tx][ty];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            xf = max(xf, __shfl_xor_sync(kFullMask, xf, o));
            xb = max(xb, __shfl_xor_sync(kFullMask, xb, o));
        }
        if (tx == 0) { s_mf[ty] = xf; s_mb[ty] = xb; }
    }
    __syncthreads();
    const int maxf = s_mf[tx], maxb = s_mb[tx];
    if (sc < N) {
        if (ty == 0) {
            a.Bf[rb * N + sc] = maxf;
            a.Bb[rb * N + sc] = maxb;
        }
        if (r < N) {
            // (NaN mantissas propagate; an exact zero has exponent kExtZeroExp and scales to 0)
            a.Kf[i] = mf * pow2i(max(ef - maxf, -1100));
            a.Kb[i] = mb * pow2i(max(eb - maxb, -1100));
        }
    }
}

// ---------------------------------------------------------------- 3. the two recurrences
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int K>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(K)); }

// One thread block evaluates one recurrence: FWD: W[1..N] from W[0] = 1;  !FWD: Wb[N-1..1] from Wb[N] = 1.
// Thread t owns rows t, t+nt, ... (R of them). Step s applies the newly known value (W[s] or Wb[N-s]) to every
// row that still needs it, then the thread whose row became complete publishes it and everybody meets at a barrier.
// The loop is issue-bound (16 warps x instructions per step), so everything is specialised at compile time and
// addresses advance by pointer increments.
// ST > 0: coefficient rows are staged through an ST-deep cp.async ring in shared memory (each thread copies and
//         later reads only its own elements, so cp.async.wait_group is the only synchronisation they need);
// ST == 0: large N, coefficients are read straight from global memory (R independent loads per thread and step).
// smem: ring int4[ST][R][nt] | sWm[N+2] | sInv[N+2] (reciprocals 1/i, so no division sits on the chain) | sWe[N+2]
template <bool FWD, int R, int ST>
__device__ __forceinline__ void recur_body(const ExArgs& a, double* smem_d) {
    static_assert(ST == 0 || (ST & (ST - 1)) == 0, "ring depth must be a power of two");
    const int tid = threadIdx.x, nt = blockDim.x, N = a.N;
    int4* ring = (int4*)smem_d;
    double* sWm = (double*)(ring + (size_t)ST * R * nt);
    double* sInv = sWm + (N + 2);
    int* sWe = (int*)(sInv + (N + 2));
    const int nsteps = FWD ? N : N - 1;     // forward: coefficient row j = s; backward: row p = N-1-s
    const int dstep = FWD ? N : -N;         // coefficient offset advance per step

    // row-validity of my R rows at coefficient row `row`: forward col in [row, N); backward col in [1, row]
    auto need = [&](int col, int row) { return FWD ? (col >= row && col < N) : (col >= 1 && col <= row); };

    const int4* gc = (FWD ? a.Cf : a.Cb) + (FWD ? 0 : (long long)(N - 1) * N) + tid;   // prefetch cursor
    int s_issue = 0;
    auto issue = [&]() {
        if (ST > 0) {
            if (s_issue < nsteps) {
                const int row = FWD ? s_issue : (N - 1 - s_issue);
                const int slot = s_issue & (ST > 0 ? ST - 1 : 0);
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (need(tid + r * nt, row)) cp_async16(&ring[(slot * R + r) * nt + tid], gc + r * nt);
            }
            cp_async_commit();
            ++s_issue;
            gc += dstep;
        }
    };

    double am[R];
    int ae[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { am[r] = 0.0; ae[r] = kExtZeroExp; }
    if (tid == 0) {
        sWm[FWD ? 0 : N] = 1.0;
        sWe[FWD ? 0 : N] = 0;
    }
    for (int i = tid; i <= N; i += nt) sInv[i] = i > 0 ? 1.0 / (double)i : 0.0;
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) issue();
    __syncthreads();

    // owner (thread, register slot) of the row that completes in the current step, tracked incrementally
    int own_t = FWD ? 0 : (N - 1) % nt;
    int own_r = FWD ? 0 : (N - 1) / nt;
    const int4* lc = (FWD ? a.Cf : a.Cb) + (FWD ? 0 : (long long)(N - 1) * N) + tid;   // direct-load cursor (ST == 0)
    const int warp = tid >> 5;

    for (int s = 0; s < nsteps; ++s) {
        const int row = FWD ? s : (N - 1 - s);
        // With one row per thread a warp has nothing left to do once all its rows are complete (forward: rows < s,
        // backward: rows > p): it leaves the loop and the per-step barrier shrinks with it.
        int bar_count = nt;
        if (R == 1) {
            if (FWD) {
                if (32 * warp + 31 < s) break;
                bar_count = nt - 32 * (s >> 5);
            } else {
                if (32 * warp > row) break;
                bar_count = 32 * ((row >> 5) + 1);
            }
        }
        double cm[R];
        int ce[R];
        if (ST > 0) {
            cp_async_wait<(ST > 1 ? ST - 2 : 0)>();    // this thread's copies for step s have landed
            const int slot = s & (ST > 0 ? ST - 1 : 0);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int4 c = ring[(slot * R + r) * nt + tid];
                cm[r] = ext_m(c);
                ce[r] = c.z;
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool nd = need(tid + r * nt, row);
                const int4 c = nd ? __ldg(lc + r * nt) : make_int4(0, 0, 0, 0);
                cm[r] = ext_m(c);
                ce[r] = c.z;
            }
            lc += dstep;
        }
        const int src = FWD ? s : (N - s);         // index of the known value: W[j] or Wb[p+1]
        const double wm = sWm[src];                 // (the backward 1/(p+1) weight is already inside Cb)
        const int we = sWe[src];
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (need(tid + r * nt, row)) ext_fma(am[r], ae[r], cm[r], ce[r], wm, we);
        // the row that just became complete: forward row v == j (-> W[j+1]); backward row l == p (-> Wb[p])
        if (tid == own_t && (FWD || row >= 1)) {
            double fm = am[0];
            int fe = ae[0];
#pragma unroll
            for (int r = 1; r < R; ++r) if (r == own_r) { fm = am[r]; fe = ae[r]; }
            if (FWD) fm *= sInv[row + 1];
            const Ext w = ext_normalize(fm, fe);
            sWm[FWD ? row + 1 : row] = w.m;
            sWe[FWD ? row + 1 : row] = w.e;
        }
        if (FWD) { if (++own_t == nt) { own_t = 0; ++own_r; } }
        else { if (--own_t < 0) { own_t = nt - 1; --own_r; } }
        if (ST > 1) issue();
        asm volatile("bar.sync 1, %0;" ::"r"(bar_count) : "memory");
        if (ST == 1) issue();
    }
    if (ST > 0) cp_async_wait<0>();
    __syncthreads();

    // V = -(ln W)/beta in parallel; publish W for the force kernel
    const double LN2 = 0.6931471805599453;
    double* Wm_g = FWD ? a.Wm : a.Wbm;
    int* We_g = FWD ? a.We : a.Wbe;
    double* V_g = FWD ? a.V : a.Vb;
    for (int i = (FWD ? 0 : 1) + tid; i <= N; i += nt) {
        const double m = sWm[i];
        const int e = sWe[i];
        Wm_g[i] = m;
        We_g[i] = e;
        const double val = -(log(m) + (double)e * LN2) / a.beta;
        if (!isfinite(val)) atomicOr(a.err, FWD ? kErrOverflowFwd : kErrOverflowBwd);
        V_g[i] = (i == (FWD ? 0 : N)) ? 0.0 : val;
    }
}

template <int R, int ST>
__global__ void __launch_bounds__(1024) k_exch_recur(ExArgs a) {
    extern __shared__ __align__(16) double smem_d[];
    if (blockIdx.x == 0) recur_body<true, R, ST>(a, smem_d);
    else recur_body<false, R, ST>(a, smem_d);
}

// ---------------------------------------------------------------- 3b. warp-decoupled recurrence (N <= 1024)
// Same arithmetic as recur_body<., 1, .>, different synchronisation. The barrier version forces all warps through
// every step in lock-step, so a step costs the serial latency of a whole warp's instruction stream (~480 cycles
// measured). Here only ONE warp is on the dependency chain at any time:
//   * "value #s" in step order is W[s] (forward) or Wb[N-s] (backward); step s turns it into value #s+1, owned by
//     the thread of row r(s) = s (forward) / N-1-s (backward). 32 consecutive steps are owned by one warp.
//   * inside its 32 owner steps a warp hands the new value from lane to lane with shuffles (no shared-memory round
//     trip, no barrier: ~100 cycles per step, DFMA latency 8.3 and shuffle ~25 cycles measured, profiles/microbench.cu);
//   * every value is also published to shared memory as ONE 16-byte entry {mantissa, exponent, tag = s+1}; data and
//     flag travel in the same 128-bit store, so no fence sits on the chain. Every other warp consumes published
//     values at its own pace -- four columns per poll, so it is faster than the owner and never holds it up -- and
//     takes over as owner when its rows come up (one ~160-cycle shared-memory hand-off per 32 steps).
//   Dependencies are acyclic (a warp only ever waits for values owned by earlier warps), so there is no deadlock.
// smem: entries int4[N+2] | coefficient ring int4[ST][nt] | sInv[N+2]
__device__ __forceinline__ int4 lds_volatile_v4(const int4* p) {
    int4 r;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts_volatile_v4(int4* p, int4 v) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.volatile.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <bool FWD, int ST>
__device__ __forceinline__ void recur_decoupled(const ExArgs& a, double* smem_d) {
    static_assert(ST >= 8 && (ST & (ST - 1)) == 0, "ring depth must be a power of two >= 8");
    constexpr int NB = 4;                            // columns a consumer applies per poll (2 was measured: slower hand-offs)
    int tid;   // read %tid.x once into a register (the compiler otherwise re-reads the special register inside the chain loop)
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int nt = blockDim.x, N = a.N, lane = tid & 31, warp = tid >> 5;
    int4* sW = (int4*)smem_d;                        // {m.lo, m.hi, e, tag}
    int4* ring = sW + (N + 2);
    double* sInv = (double*)(ring + (size_t)ST * nt);
    const int nsteps = FWD ? N : N - 1;
    const int dstep = FWD ? N : -N;
    const int v = tid;                                              // my row
    const bool row_ok = FWD ? (v < N) : (v >= 1 && v < N);
    // steps in which my warp owns the completing row
    const int own_lo = FWD ? 32 * warp : max(0, N - 1 - (32 * warp + 31));
    const int own_hi = min(nsteps - 1, FWD ? 32 * warp + 31 : N - 1 - 32 * warp);

    // does my row take part in step s ?  forward: v >= s; backward: v <= N-1-s  -- both are "s <= last_need", one
    // integer compare against a per-thread constant (nothing is re-derived from %tid inside the chain loops)
    const int last_need = row_ok ? (FWD ? v : N - 1 - v) : -1;
    auto need = [&](int s) { return s <= last_need; };
    auto idx_of = [&](int s) { return FWD ? s : N - s; };           // where value #s lives
    const int4* gc = (FWD ? a.Cf : a.Cb) + (FWD ? 0 : (long long)(N - 1) * N) + tid;
    int s_issue = 0;
    auto issue = [&]() {
        if (s_issue < nsteps && need(s_issue))
            cp_async16(&ring[(s_issue & (ST - 1)) * nt + tid], gc);
        cp_async_commit();
        ++s_issue;
        gc += dstep;
    };
    auto wait_value = [&](int s) {                                  // spin until value #s is published
        int4 w;
        do { w = lds_volatile_v4(&sW[idx_of(s)]); } while (w.w != s + 1);
        return w;
    };
    // (Backing far consumers off with __nanosleep between polls was measured and does not help: the owner's step
    // time is not set by the pollers.)

    for (int i = tid; i <= N + 1; i += nt) sW[i] = make_int4(0, 0, 0, 0);
    for (int i = tid; i <= N; i += nt) sInv[i] = i > 0 ? 1.0 / (double)i : 0.0;
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) issue();
    __syncthreads();
    if (tid == 0) sts_volatile_v4(&sW[idx_of(0)], make_int4(__double2loint(1.0), __double2hiint(1.0), 0, 1));

    double am = 0.0;
    int ae = kExtZeroExp;
    int s = 0;
    const long long t_begin = a.dbg ? clock64() : 0;
    // ---- consumer phase: columns owned by earlier warps; every row of my warp takes part in all of them
    while (s < own_lo) {
        if (own_lo - s >= NB) {
            wait_value(s + NB - 1);                                 // values are published in order
            cp_async_wait<ST - 1 - NB>();
            double cm[NB], wm[NB];
            int ce[NB], we[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const int4 w = sW[idx_of(s + k)];
                wm[k] = __hiloint2double(w.y, w.x);
                we[k] = w.z;
                const int4 c = ring[((s + k) & (ST - 1)) * nt + tid];
                cm[k] = ext_m(c);
                ce[k] = c.z;
            }
            if (row_ok) {
#pragma unroll
                for (int k = 0; k < NB; ++k) ext_fma(am, ae, cm[k], ce[k], wm[k], we[k]);
            }
#pragma unroll
            for (int k = 0; k < NB; ++k) issue();
            s += NB;
        } else {
            const int4 w = wait_value(s);
            cp_async_wait<ST - 1 - NB>();
            const int4 c = ring[(s & (ST - 1)) * nt + tid];
            if (row_ok) ext_fma(am, ae, ext_m(c), c.z, __hiloint2double(w.y, w.x), w.z);
            issue();
            ++s;
        }
    }
    const long long t_own0 = a.dbg ? clock64() : 0;
    // ---- owner phase: the completing rows are my warp's; the chain runs lane to lane through shuffles
    if (s <= own_hi) {
        auto row_of = [&](int st) { return FWD ? st : (N - 1 - st); };
        const int4 w0 = wait_value(s);
        double wm = __hiloint2double(w0.y, w0.x);
        int we = w0.z;
#pragma unroll 1
        for (; s <= own_hi; ++s) {
            cp_async_wait<ST - 1 - NB>();
            if (need(s)) {
                const int4 c = ring[(s & (ST - 1)) * nt + tid];
                ext_fma(am, ae, ext_m(c), c.z, wm, we);
            }
            const int lane_o = row_of(s) & 31;
            const Ext fin = ext_normalize(FWD ? am * sInv[s + 1] : am, ae);
            wm = __shfl_sync(kFullMask, fin.m, lane_o);
            we = __shfl_sync(kFullMask, fin.e, lane_o);
            if (lane == lane_o)
                sts_volatile_v4(&sW[idx_of(s + 1)], make_int4(__double2loint(fin.m), __double2hiint(fin.m), fin.e, s + 2));
            issue();
        }
    }
    if (a.dbg && lane == 0) {
        long long* o = a.dbg + ((FWD ? 0 : 32) + warp) * 3;
        o[0] = t_own0 - t_begin;            // cycles spent as a consumer (incl. waiting)
        o[1] = clock64() - t_own0;          // cycles spent as the owner
        o[2] = t_begin;
    }
    cp_async_wait<0>();
    __syncthreads();

    // V = -(ln W)/beta in parallel; publish W for the force kernel
    const double LN2 = 0.6931471805599453;
    double* Wm_g = FWD ? a.Wm : a.Wbm;
    int* We_g = FWD ? a.We : a.Wbe;
    double* V_g = FWD ? a.V : a.Vb;
    for (int i = (FWD ? 0 : 1) + tid; i <= N; i += nt) {
        const int4 w = sW[i];
        const Ext wn = ext_normalize(__hiloint2double(w.y, w.x), w.z);   // fast-path entries are not normalised
        const double m = wn.m;
        const int e = wn.e;
        Wm_g[i] = m;
        We_g[i] = e;
        const double val = -(log(m) + (double)e * LN2) / a.beta;
        if (!isfinite(val)) atomicOr(a.err, FWD ? kErrOverflowFwd : kErrOverflowBwd);
        V_g[i] = (i == (FWD ? 0 : N)) ? 0.0 : val;
    }
}

// (am, ae) += acc * 2^e for a plain non-negative double acc; out of line: it runs once per 32 columns
static __device__ __noinline__ void ext_fold(double& am, int& ae, double acc, int e) {
    if (acc > 0.0 || acc != acc) {
        const Ext t = ext_normalize(acc, e);
        const int emax = max(ae, t.e);
        am = fma(am, pow2i(max(ae - emax, -1100)), t.m * pow2i(max(t.e - emax, -1100)));
        ae = emax;
    }
}

// ---------------------------------------------------------------- 3c. fast warp-decoupled recurrence (N <= 512)
// Same protocol as recur_decoupled (tagged 16-byte entries, one owner warp on the chain, everybody else consuming),
// with both sides of the work made cheap enough that the chain itself is what is left:
//   * consumers read the BLOCK-SCALED factors K (plain doubles, k_exch_coeff_tiles) through an 8-byte cp.async ring and
//     accumulate a 32-column phase as a plain dot product, acc += K * omega -- one FP64 instruction per column instead
//     of an extended-range multiply-add (~25 instructions). Published entries of a fast phase share one binary exponent
//     E, so the partial sum is folded into the extended-range accumulator once per phase (and whenever an entry's
//     exponent differs, which is what entries of the exact fallback path do -- then every column is folded
//     separately, still correct). This matters twice: a consumer shares its scheduler with the owner (in-order issue,
//     4 warps each), and the NEXT owner cannot start before it has applied every earlier column.
//   * the owner phase is the block-scaled loop described below (operands prefetched, publish software-pipelined).
// smem: entries int4[N+2] | factor ring double[48][512] | sInv[N+2]
template <bool FWD>
__device__ __forceinline__ void recur_fast(const ExArgs& a, double* smem_d) {
    constexpr int ST = 48;                           // ring depth: 12 groups of 4 columns per thread
    constexpr int AHEAD = 11;                        // groups in flight ahead of the consumer (global latency ~1000+ cycles
                                                     // against ~100 cycles per column: 12 columns ahead was measured too few)
    constexpr int RS = 512;                          // row stride of the ring: a compile-time constant, so every
                                                     // shared-memory address in the loops is base + immediate
    int tid;   // read %tid.x once into a register (the compiler otherwise re-reads the special register inside the chain loop)
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int nt = blockDim.x, N = a.N, lane = tid & 31, warp = tid >> 5;
    int4* sW = (int4*)smem_d;                        // {omega.lo, omega.hi, e, tag}
    double* ring = (double*)(sW + (N + 2));
    double* sInv = ring + (size_t)ST * RS;
    const int nsteps = FWD ? N : N - 1;
    const int dstep = FWD ? N : -N;
    const int v = tid;                                              // my row
    const bool row_ok = FWD ? (v < N) : (v >= 1 && v < N);
    // steps in which my warp owns the completing row; they use the factor rows r = 32 warp .. 32 warp + 31
    const int own_lo = FWD ? 32 * warp : max(0, N - 1 - (32 * warp + 31));
    const int own_hi = min(nsteps - 1, FWD ? 32 * warp + 31 : N - 1 - 32 * warp);
    const int last_need = row_ok ? (FWD ? v : N - 1 - v) : -1;
    auto need = [&](int s) { return s <= last_need; };
    auto idx_of = [&](int s) { return FWD ? s : N - s; };           // where value #s lives
    auto row_of = [&](int st) { return FWD ? st : (N - 1 - st); };  // factor row used by step st
    const double* Kg = FWD ? a.Kf : a.Kb;
    const int* Bg = FWD ? a.Bf : a.Bb;
    // Factor ring, private per thread: column s of my row lives in slot (s + off) % 48, with `off` chosen per warp so
    // that the 32 columns of the warp's OWN phase sit in slots 0..34 without wrapping -- the owner loop then walks a
    // plain pointer. Columns travel in aligned groups of four (one cp.async group each), AHEAD groups ahead of the
    // consumer; the copies run on through the own phase (s <= own_hi), whose factors are read from the same ring.
    // A group may reach up to three columns past own_hi (or past the table: the allocation is padded) -- never read.
    const int off = (ST * 1024 - (own_lo & ~3)) % ST;
    double* const ring_me = ring + tid;
    auto slot_ptr = [&](int s) { return ring_me + ((s + off) % ST) * RS; };
    const double* gk = Kg + (FWD ? 0 : (long long)(N - 1) * N) + tid;
    int s_issue = 0;
    double* ip = slot_ptr(0);                                       // slot of column s_issue
    auto issue4 = [&]() {
        if (s_issue <= own_hi && row_ok) {
#pragma unroll
            for (int k = 0; k < 4; ++k) cp_async8(ip + k * RS, gk + (long long)k * dstep);
        }
        cp_async_commit();
        s_issue += 4;
        gk += 4 * (long long)dstep;
        ip += 4 * RS;
        if (ip >= ring_me + ST * RS) ip -= ST * RS;
    };
    auto wait_value = [&](int s) {                                  // spin until value #s is published
        int4 w;                     // (a __nanosleep(40 / 200) back-off between polls was measured: no gain / 1.5% slower)
        do { w = lds_volatile_v4(&sW[idx_of(s)]); } while (w.w != s + 1);
        return w;
    };

    if (!row_ok) {                             // rows outside the recurrence never copy: their ring slots stay 0
#pragma unroll 4
        for (int k = 0; k < ST; ++k) ring_me[k * RS] = 0.0;
    }
    for (int i = tid; i <= N + 1; i += nt) sW[i] = make_int4(0, 0, 0, 0);
    for (int i = tid; i <= N; i += nt) sInv[i] = i > 0 ? 1.0 / (double)i : 0.0;
#pragma unroll 1
    for (int g = 0; g < AHEAD; ++g) issue4();
    // my own block's scale: the owner phase works on factors K (relative to 2^Bown) and folds 2^Bown into the weight
    const int Bown = row_ok ? Bg[warp * N + v] : kExtZeroExp;
    __syncthreads();
    if (tid == 0) sts_volatile_v4(&sW[idx_of(0)], make_int4(__double2loint(1.0), __double2hiint(1.0), 0, 1));

    double am = 0.0;                                 // extended-range accumulator of my row (normalised)
    int ae = kExtZeroExp;
    double acc = 0.0;                                // plain partial sum of the current run of columns, scale 2^(Bq + Eph)
    int Eph = 0, Bq = 0;
    auto fold = [&]() {                              // am * 2^ae += acc * 2^(Bq + Eph)
        ext_fold(am, ae, acc, Bq + Eph);
        acc = 0.0;
    };
    auto apply = [&](const int4& w, double kk) {     // one column the careful way (warp-uniform branch: w is broadcast)
        if (w.z != Eph) { fold(); Eph = w.z; }
        acc = fma(kk, __hiloint2double(w.y, w.x), acc);
    };
    int s = 0;
    double* rp = slot_ptr(0);                        // slot of column s
    auto advance = [&](int n) {
        rp += n * RS;
        if (rp >= ring_me + ST * RS) rp -= ST * RS;
    };
    const long long t_begin = a.dbg ? clock64() : 0;
    // ---- consumer phases: columns owned by earlier warps, one 32-row factor block q at a time
    // invariant: issued groups = floor(s / 4) + AHEAD, so "all but the AHEAD-1 newest groups" covers the group of column s
    int Bnext = (own_lo > 0 && row_ok) ? Bg[(row_of(0) >> 5) * N + v] : 0;
    while (s < own_lo) {
        const int q = row_of(s) >> 5;
        const int s_end = min(own_lo, FWD ? (q + 1) * 32 : N - q * 32);   // one past this block's last step
        Bq = Bnext;
        if (s_end < own_lo && row_ok) Bnext = Bg[(row_of(s_end) >> 5) * N + v];
        while (s < s_end) {
            if ((s & 3) == 0 && s + 4 <= s_end) {                   // a whole group: four columns per poll
                wait_value(s + 3);                                   // values are published in order
                cp_async_wait<AHEAD - 1>();
                const int4* wp = &sW[idx_of(s)];
                int4 w[4];
                double kk[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    w[k] = wp[FWD ? k : -k];
                    kk[k] = rp[k * RS];
                }
                if (((w[0].z ^ Eph) | (w[1].z ^ Eph) | (w[2].z ^ Eph) | (w[3].z ^ Eph)) == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc = fma(kk[k], __hiloint2double(w[k].y, w[k].x), acc);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) apply(w[k], kk[k]);
                }
                issue4();
                s += 4;
                advance(4);
            } else {                                                 // ragged head / tail of a block (N % 4 != 0)
                const int4 w = wait_value(s);
                cp_async_wait<AHEAD - 1>();
                apply(w, *rp);
                if ((s & 3) == 3) issue4();
                ++s;
                advance(1);
            }
        }
        fold();                                                      // the next block has its own scale
    }
    const long long t_own0 = a.dbg ? clock64() : 0;
    long long t_loop0 = t_own0;
    // ---- owner phase: the completing rows are my warp's; the chain runs lane to lane through shuffles
    if (s <= own_hi) {
        int s_resume = s;                 // first step whose result is NOT yet published
        // Block-scaled fast path. Inside one 32-step owner phase every W is written as omega * 2^E with a common
        // binary exponent E and a plain double omega, and every row keeps its sum in units of 2^(E + Bown) (Bown: the
        // scale of the row's own factor block), so a step is ONE fma + ONE multiply + ONE shuffle (no exponent
        // alignment, no normalisation on the chain). It is exact as long as every omega of the phase stays within
        // 2^+-400 of the phase's first value and no accumulator starts above 2^600 on that scale: whatever a plain
        // double then flushes to zero is < 2^-622 and negligible against the result. The moment a value leaves
        // that window the phase is redone from the untouched extended-range accumulators (below); nothing out of
        // range has been published by then. The loop is deliberately NOT unrolled: it runs once per warp, and
        // straight-line code executed once is bound by instruction fetch.
        // A published entry of such a phase is {omega (plain double), E, tag}: no normalisation work in the
        // loop; entries are therefore NOT normalised in general (consumers treat them as mantissa * 2^E with a
        // mantissa anywhere in 2^+-400), and each phase renormalises its first value so E does not drift.
        const int4 w0 = wait_value(s);
        const Ext n0 = ext_normalize(__hiloint2double(w0.y, w0.x), w0.z);
        const int E = n0.e;
        double om = n0.m;
        const int d = ae - E - Bown;
        double A = (am == 0.0) ? 0.0 : ext_to_double(am, min(d, 600));
        bool exact = __any_sync(kFullMask, (am != 0.0) && (d > 600)) || !(om > 0.0);
        // every group that holds an own-phase column has been issued (AHEAD * 4 >= 35); make sure they have landed
        cp_async_wait<AHEAD - 9>();
        if (a.dbg) t_loop0 = clock64();
        if (!exact) {
            // Measured on B200 (profiles/microbench2.cu): fma + mul + shuffle = 46 cycles per step; shared-memory
            // operands loaded inside the step add 65, a compare-and-branch range check 55, a divergent publish 91.
            // Hence: the factor is prefetched one step ahead, the weight 2^Bown / (v+1) is a per-lane constant (only
            // the completing lane's product is used), the range check is an integer test on the exponent bits folded
            // into a predicate (no branch; once it fails nothing more is published and the phase is redone
            // below). The loop is software-pipelined by one step: iteration st issues the chain's fma and
            // multiply first and only then checks and publishes the value the PREVIOUS iteration produced (every
            // lane holds it after the broadcast shuffle), so the integer work and the store sit in the shadow of
            // the FP64 latency. Every instruction here also costs one issue slot per co-resident consumer warp
            // (in-order issue, 4 warps per scheduler), so the body is kept to ~25 instructions.
            const double* kp = rp;                                       // own-phase slots do not wrap (see `off`)
            double kcur = *kp;
            const double myw = (FWD ? sInv[min(v + 1, N)] : 1.0) * pow2i(max(Bown, -1100));
            bool okall = true;
            int lane_o = row_of(s) & 31;
            int4* wp = &sW[idx_of(s)];
#pragma unroll 1
            for (int st = s; st <= own_hi; ++st) {
                kp += RS;
                const double knext = *kp;
                A = fma(kcur, om, A);                                     // K is 0 where a row takes no part
                const double t = A * myw;
                {   // value #st (st == s: re-stores the hand-off value, renormalised -- same number, same tag)
                    const unsigned hi = (unsigned)__double2hiint(om);
                    okall = okall && (hi - (623u << 20) < (800u << 20));   // positive, within 2^+-400, not NaN/inf/0
                    if (okall && lane == 0) sts_volatile_v4(wp, make_int4(__double2loint(om), (int)hi, E, st + 1));
                    s_resume = okall ? st : s_resume;
                }
                om = __shfl_sync(kFullMask, t, lane_o);
                lane_o = (lane_o + (FWD ? 1 : -1)) & 31;
                wp += FWD ? 1 : -1;
                kcur = knext;
            }
            {   // the phase's last value, #(own_hi + 1)
                const unsigned hi = (unsigned)__double2hiint(om);
                okall = okall && (hi - (623u << 20) < (800u << 20));
                if (okall && lane == 0) sts_volatile_v4(wp, make_int4(__double2loint(om), (int)hi, E, own_hi + 2));
                s_resume = okall ? own_hi + 1 : s_resume;
            }
            exact = !okall;
        }
        if (exact) {
            // Extended-range fallback: restarts the phase from the pre-phase accumulators, re-applies the already
            // published columns without publishing them again (factors straight from global memory: rare), and
            // continues from the first unpublished step; its entries are normalised {mantissa, exponent}.
            const int4* Cg = FWD ? a.Cf : a.Cb;
            double wm = __hiloint2double(w0.y, w0.x);
            int we = w0.z;
#pragma unroll 1
            for (; s <= own_hi; ++s) {
                if (s > own_lo && s <= s_resume) {
                    const int4 w = wait_value(s);
                    wm = __hiloint2double(w.y, w.x);
                    we = w.z;
                }
                if (need(s)) {
                    const int4 c = __ldg(&Cg[(long long)row_of(s) * N + v]);
                    ext_fma(am, ae, ext_m(c), c.z, wm, we);
                }
                if (s >= s_resume) {
                    const int lane_o = row_of(s) & 31;
                    const Ext fin = ext_normalize(FWD ? am * sInv[s + 1] : am, ae);
                    wm = __shfl_sync(kFullMask, fin.m, lane_o);
                    we = __shfl_sync(kFullMask, fin.e, lane_o);
                    if (lane == lane_o)
                        sts_volatile_v4(&sW[idx_of(s + 1)], make_int4(__double2loint(fin.m), __double2hiint(fin.m), fin.e, s + 2));
                }
            }
        }
    }
    if (a.dbg && lane == 0) {
        long long* o = a.dbg + ((FWD ? 0 : 32) + warp) * 3;
        o[0] = t_own0 - t_begin;            // cycles spent as a consumer (incl. waiting)
        o[1] = clock64() - t_own0;          // cycles spent as the owner
        o[2] = t_loop0 - t_own0;            // ... of which: hand-off wait + owner prologue
    }
    cp_async_wait<0>();
    __syncthreads();

    // V = -(ln W)/beta in parallel; publish W (normalised) for the force kernel
    const double LN2 = 0.6931471805599453;
    double* Wm_g = FWD ? a.Wm : a.Wbm;
    int* We_g = FWD ? a.We : a.Wbe;
    double* V_g = FWD ? a.V : a.Vb;
    for (int i = (FWD ? 0 : 1) + tid; i <= N; i += nt) {
        const int4 w = sW[i];
        const Ext wn = ext_normalize(__hiloint2double(w.y, w.x), w.z);
        Wm_g[i] = wn.m;
        We_g[i] = wn.e;
        const double val = -(log(wn.m) + (double)wn.e * LN2) / a.beta;
        if (!isfinite(val)) atomicOr(a.err, FWD ? kErrOverflowFwd : kErrOverflowBwd);
        V_g[i] = (i == (FWD ? 0 : N)) ? 0.0 : val;
    }
}

__global__ void __launch_bounds__(512) k_exch_recur_fast(ExArgs a) {
    extern __shared__ __align__(16) double smem_d[];
    if (blockIdx.x == 0) recur_fast<true>(a, smem_d);
    else recur_fast<false>(a, smem_d);
}

template <int ST>
__global__ void __launch_bounds__(1024) k_exch_recur_dec(ExArgs a) {
    extern __shared__ __align__(16) double smem_d[];
    if (blockIdx.x == 0) recur_decoupled<true, ST>(a, smem_d);
    else recur_decoupled<false, ST>(a, smem_d);
}

// ---------------------------------------------------------------- 4. exterior spring forces (K7 + K8)
// one block of kFT threads per (exterior bead, particle l): the u-sum is latency-bound (a handful of dependent
// global loads per term), so it is spread over 4 warps and the partials are combined in a fixed order
constexpr int kFT = 128;
template <int D>
__device__ __forceinline__ void block_reduce_store(double (&acc)[D], double* sred) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = warp_sum(acc[c]);
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) sred[warp * D + c] = acc[c];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double t = 0.0;
            for (int k = 0; k < kFT / 32; ++k) t += sred[k * D + c];
            acc[c] = t;
        }
    }
}

template <int D>
__global__ void __launch_bounds__(kFT) k_exch_forces(ExArgs a) {
    __shared__ double sred[(kFT / 32) * D];
    const int lane = threadIdx.x;          // position in the u-stride of this block
    const int w = blockIdx.x;
    const int N = a.N;
    const int which = w / N, l = w % N;   // 0: first bead, 1: last bead
    if ((which == 0 && !a.do_first) || (which == 1 && !a.do_last)) return;
    const double iWN = 1.0 / a.Wm[N];
    const int eWN = a.We[N];
    double acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.0;

    if (which == 0) {
        // f_l = k [ sum_{u=max(0,l-1)}^{N-1} P(u->l) mi(r^P_u - r^1_l) + mi(r^2_l - r^1_l) ]
        const double wl = a.Wm[l] * iWN;
        const int el = a.We[l] - eWN;
#pragma unroll 2
        for (int u = max(0, l - 1) + lane; u < N; u += kFT) {
            double pr;
            if (u == l - 1) {
                pr = 1.0 - ext_to_double(wl * a.Wbm[l], el + a.Wbe[l]);
            } else {
                const size_t ci = (size_t)l * N + u;
                const int4 c = __ldg(&a.Cf[ci]);
                pr = ext_to_double(wl * ext_m(c) * a.Wbm[u + 1] * a.Inv[u + 1], el + c.z + a.Wbe[u + 1]);
            }
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dx = a.xP[(size_t)c * N + u] - a.x1[(size_t)c * N + l];
                if (a.pbc) dx = min_image(dx, a.L, a.invL);
                acc[c] = fma(pr, dx, acc[c]);
            }
        }
        block_reduce_store<D>(acc, sred);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dx = a.x2[(size_t)c * N + l] - a.x1[(size_t)c * N + l];
                if (a.pbc) dx = min_image(dx, a.L, a.invL);
                a.F[(size_t)c * N + l] = (acc[c] + dx) * a.k;
            }
        }
    } else {
        // f_l = k [ sum_{u=0}^{min(l+1,N-1)} P(l->u) mi(r^1_u - r^P_l) + mi(r^{P-1}_l - r^P_l) ]
        const double wb = a.Wbm[l + 1] * iWN;
        const int eb = a.Wbe[l + 1] - eWN;
        const int uend = min(l + 1, N - 1);
#pragma unroll 2
        for (int u = lane; u <= uend; u += kFT) {
            double pr;
            if (u == l + 1) {
                pr = 1.0 - ext_to_double(wb * a.Wm[l + 1], eb + a.We[l + 1]);
            } else {
                const size_t ci = (size_t)l * N + u;
                const int4 c = __ldg(&a.Cb[ci]);                       // = c(u,l) / (l+1)
                pr = ext_to_double(wb * ext_m(c) * a.Wm[u], eb + c.z + a.We[u]);
            }
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dx = a.x1[(size_t)c * N + u] - a.xP[(size_t)c * N + l];
                if (a.pbc) dx = min_image(dx, a.L, a.invL);
                acc[c] = fma(pr, dx, acc[c]);
            }
        }
        block_reduce_store<D>(acc, sred);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dx = a.xPm1[(size_t)c * N + l] - a.xP[(size_t)c * N + l];
                if (a.pbc) dx = min_image(dx, a.L, a.invL);
                a.F[(size_t)(D + c) * N + l] = (acc[c] + dx) * a.k;
            }
        }
    }
    if (w == 0 && lane == 0) a.Vb[0] = a.V[N];   // V_backwards[0] = V[N] (quadratic_bosonic_exchange.cpp:127)
}

// ---------------------------------------------------------------- estimators (bead-0 owner only), one block
// e[m] = sum_{j<m} w(m,j) (e[j] - E^{[j..m-1]}),  w(m,j) = c(j,m-1) W[j] / (m W[m])   (primEstimator :250-279)
// Column-wise; the weights are extended-range products (no exp), the chain per step is one FMA + one barrier.
template <int D, int R>
__global__ void __launch_bounds__(1024) k_exch_estimators(ExArgs a) {
    extern __shared__ double se[];   // e[0..N]
    __shared__ double red[32];
    const int tid = threadIdx.x, nt = blockDim.x, N = a.N;
    double acc[R], iwm[R];
    int iwe[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        acc[r] = 0.0;
        const int v = tid + r * nt;
        iwm[r] = v < N ? a.Inv[v + 1] / a.Wm[v + 1] : 0.0;
        iwe[r] = v < N ? -a.We[v + 1] : 0;
    }
    if (tid == 0) se[0] = 0.0;
    __syncthreads();
    for (int j = 0; j < N; ++j) {
        const double ej = se[j], wjm = a.Wm[j];
        const int wje = a.We[j];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int v = tid + r * nt;
            if (v >= j && v < N) {
                const size_t ci = (size_t)j * N + v;
                const int4 c = __ldg(&a.Cf[ci]);
                const double wgt = ext_to_double(ext_m(c) * wjm * iwm[r], c.z + wje + iwe[r]);
                acc[r] = fma(wgt, ej - cycle_energy<D>(a, j, v), acc[r]);
            }
        }
        const int owner = j % nt, rr = j / nt;
        if (tid == owner) {
            double val = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) if (r == rr) val = acc[r];
            se[j + 1] = val;
        }
        __syncthreads();
    }
    // sum_m E^{[m..m]} (prob_dist) in a fixed order
    double part[1] = {0.0};
    for (int m = tid; m < N; m += nt) part[0] += 0.5 * a.k * dist2<D>(a, a.x1, m, a.xP, m);
    block_sum<1>(part, red);
    if (tid == 0) {
        a.obs->prim_est = se[N];
        a.obs->v_n = a.V[N];
        a.obs->e_diag_sum = part[0];
        a.obs->e_full = cycle_energy<D>(a, 0, N - 1);
    }
}

// ---------------------------------------------------------------- on-demand tables (tests / debugging)
template <int D>
__global__ void k_exch_table_E(ExArgs a, double* out) {
    const long long tot = (long long)a.N * (a.N + 1) / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        // serial order of the reference: index = m(m+1)/2 - k, m = v+1 in [1,N], k = v-u+1 in [1,m]
        // rows m occupy [m(m-1)/2, m(m+1)/2): offset within the row = m - k = u
        long long m = (long long)((sqrt(8.0 * (double)i + 1.0) - 1.0) * 0.5) + 1;
        while (m * (m - 1) / 2 > i) --m;
        while (m * (m + 1) / 2 <= i) ++m;
        int u = (int)(i - m * (m - 1) / 2);
        out[i] = cycle_energy<D>(a, u, (int)m - 1);
    }
}

__global__ void k_exch_table_prob(ExArgs a, double* out) {
    const int N = a.N;
    const long long tot = (long long)N * N;
    const double iWN = 1.0 / a.Wm[N];
    const int eWN = a.We[N];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i / N), u = (int)(i % N);
        double pr = 0.0;
        if (u == l + 1) pr = 1.0 - ext_to_double(a.Wm[l + 1] * a.Wbm[l + 1] * iWN, a.We[l + 1] + a.Wbe[l + 1] - eWN);
        else if (u <= l)
            pr = ext_to_double(a.Wm[u] * ext_m(a.Cb[i]) * a.Wbm[l + 1] * iWN, a.We[u] + a.Cb[i].z + a.Wbe[l + 1] - eWN);
        out[i] = pr;
    }
}

// ---------------------------------------------------------------- host side
static ExArgs make_args(Sim* s) {
    ExArgs a;
    const size_t S = s->S;
    // slab index (with halo offset): owned bead j -> x + (j+1) S
    if (s->has_first) {
        a.x1 = s->x + 1 * S;
        a.xP = s->all_local ? s->x + (size_t)s->Ploc * S : s->x;       // halo before first = bead P-1
        a.x2 = s->x + 2 * S;                                           // next of first (owned or trailing halo)
    } else {
        a.x1 = s->x + (size_t)(s->Ploc + 1) * S;                       // trailing halo = bead 0
        a.xP = s->x + (size_t)s->Ploc * S;
        a.x2 = nullptr;
    }
    a.xPm1 = s->has_last ? s->x + (size_t)(s->Ploc - 1) * S : nullptr; // previous of last (owned or leading halo)
    const size_t NN = (size_t)s->N * s->N;
    a.A = s->exA;
    a.Inv = s->exA + s->N;
    a.Cf = s->exC; a.Cb = s->exC + NN;
    a.Kf = s->exK; a.Kb = s->exK ? s->exK + NN + 512 : nullptr;
    a.Bf = s->exB; a.Bb = s->exB ? s->exB + (size_t)((s->N + 31) / 32) * s->N : nullptr;
    a.Wm = s->exWm; a.Wbm = s->exWm + (s->N + 1);
    a.We = s->exWe; a.Wbe = s->exWe + (s->N + 1);
    a.V = s->exV; a.Vb = s->exVb; a.F = s->exF;
    a.obs = s->obs_d; a.err = s->err_d; a.dbg = s->dbg_buf;
    a.N = s->N; a.D = s->D; a.pbc = s->cfg.pbc;
    a.do_first = s->has_first; a.do_last = s->has_last;
    a.k = s->kspring; a.beta = s->exch_beta; a.h = 0.5 * s->exch_beta * s->kspring;
    a.L = s->L; a.invL = 1.0 / s->L;
    return a;
}

static int rows_per_thread(int N, int& nt) {
    nt = ((N + 31) / 32) * 32;
    if (nt > 1024) nt = 1024;
    int r = (N + nt - 1) / nt;
    int R = 1;
    while (R < r) R <<= 1;
    return R;
}

template <int R, int ST>
static int launch_recur(Sim* s, const ExArgs& a, cudaStream_t st, int nt) {
    const size_t smem = (size_t)ST * R * nt * sizeof(int4) + (size_t)(s->N + 2) * (2 * sizeof(double) + sizeof(int)) + 16;
    if (smem > 220 * 1024) {
        s->err = "natoms too large for the single-block exchange recursion of this build";
        return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(k_exch_recur<R, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_exch_recur<R, ST><<<2, nt, smem, st>>>(a);
    return PIMDB_OK;
}

static int run_recursion(Sim* s, const ExArgs& a, cudaStream_t st) {
    int nt;
    const int R = rows_per_thread(s->N, nt);
    if (R == 1 && !getenv("PIMDB_EXCH_BARRIER")) {           // N <= 1024: warp-decoupled kernel
        // coefficient ring: 16 rows in flight up to 512 rows per block, 8 beyond (shared-memory budget)
        const int ST = nt <= 512 ? 16 : 8;
        const size_t smem = (size_t)(s->N + 2) * (sizeof(int4) + sizeof(double)) + (size_t)ST * nt * sizeof(int4) + 16;
        if (nt <= 512 && a.Kf && !getenv("PIMDB_EXCH_NOFAST")) {
            // fast kernel: block-scaled factors, plain-double consumers and owner phase, 48-column factor ring
            const size_t smem_fast = (size_t)(s->N + 2) * (sizeof(int4) + sizeof(double)) + (size_t)48 * 512 * sizeof(double) + 16;
            if (smem_fast > 48 * 1024)
                cudaFuncSetAttribute(k_exch_recur_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast);
            k_exch_recur_fast<<<2, nt, smem_fast, st>>>(a);
        } else if (ST == 16) {
            if (smem > 48 * 1024)
                cudaFuncSetAttribute(k_exch_recur_dec<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_exch_recur_dec<16><<<2, nt, smem, st>>>(a);
        } else {
            cudaFuncSetAttribute(k_exch_recur_dec<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_exch_recur_dec<8><<<2, nt, smem, st>>>(a);
        }
        return PIMDB_OK;
    }
    switch (R) {
        case 1: return launch_recur<1, 8>(s, a, st, nt);    // N <= 1024 (barrier variant, PIMDB_EXCH_BARRIER=1)
        case 2: return launch_recur<2, 4>(s, a, st, nt);    // N <= 2048
        case 4: return launch_recur<4, 0>(s, a, st, nt);    // N <= 4096: direct global loads
        case 8: return launch_recur<8, 0>(s, a, st, nt);    // N <= 8192
        case 16: return launch_recur<16, 0>(s, a, st, nt);  // N <= 16384
        default:
            s->err = "natoms too large for the single-block exchange recursion (max 16384)";
            return PIMDB_ERR_INVALID_ARGUMENT;
    }
}

// part 0: prefix sums + Boltzmann factors (fully parallel, a few microseconds);
// part 1: the two recurrences (2 thread blocks, latency-bound) + the exterior forces.
// They are separate so the caller can start part 1 on a side stream *before* it floods the GPU with pair tiles.
template <int D>
static int exchange_impl(Sim* s, cudaStream_t st, int part) {
    ExArgs a = make_args(s);
    if (part == 0) {
        if (a.Kf) {      // N <= 512: the tiles recompute the prefix sums themselves (one launch)
            const int nb = (s->N + 31) / 32;
            k_exch_coeff_tiles<D><<<nb * nb, 1024, 0, st>>>(a);
            s->launches += 1;
        } else {
            k_exch_prefix<D><<<1, 1024, 0, st>>>(a);
            k_exch_coeff<D><<<grid_for((size_t)s->N * s->N, 256, 16 * kNumSM), 256, 0, st>>>(a);
            s->launches += 2;
        }
    } else {
        int rc = run_recursion(s, a, st);
        if (rc != PIMDB_OK) return rc;
        k_exch_forces<D><<<2 * s->N, kFT, 0, st>>>(a);
        s->launches += 2;
    }
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_exchange_part(Sim* s, cudaStream_t st, int part) {
    if (s->D == 1) return exchange_impl<1>(s, st, part);
    if (s->D == 2) return exchange_impl<2>(s, st, part);
    return exchange_impl<3>(s, st, part);
}

int launch_exchange(Sim* s, cudaStream_t st) {
    int rc = launch_exchange_part(s, st, 0);
    if (rc != PIMDB_OK) return rc;
    return launch_exchange_part(s, st, 1);
}

template <int D>
static int estimators_impl(Sim* s) {
    ExArgs a = make_args(s);
    int nt;
    const int R = rows_per_thread(s->N, nt);
    const size_t smem = (size_t)(s->N + 1) * sizeof(double);
#define PIMDB_EST(RR)                                                                                           \
    case RR:                                                                                                    \
        if (smem > 48 * 1024)                                                                                   \
            cudaFuncSetAttribute(k_exch_estimators<D, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        k_exch_estimators<D, RR><<<1, nt, smem, s->stream>>>(a);                                                \
        break;
    switch (R) {
        PIMDB_EST(1) PIMDB_EST(2) PIMDB_EST(4) PIMDB_EST(8) PIMDB_EST(16)
        default:
            s->err = "natoms too large for the exchange estimator kernel (max 16384)";
            return PIMDB_ERR_INVALID_ARGUMENT;
    }
#undef PIMDB_EST
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_exchange_estimators(Sim* s) {
    if (s->D == 1) return estimators_impl<1>(s);
    if (s->D == 2) return estimators_impl<2>(s);
    return estimators_impl<3>(s);
}

template <int D>
static int tables_impl(Sim* s, int table) {
    ExArgs a = make_args(s);
    const size_t n = table == PIMDB_EXCH_E ? (size_t)s->N * (s->N + 1) / 2 : (size_t)s->N * s->N;
    if (table == PIMDB_EXCH_E) k_exch_table_E<D><<<grid_for(n, 256), 256, 0, s->stream>>>(a, s->exTab);
    else k_exch_table_prob<<<grid_for(n, 256), 256, 0, s->stream>>>(a, s->exTab);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_exchange_tables(Sim* s, int table) {
    if (s->D == 1) return tables_impl<1>(s, table);
    if (s->D == 2) return tables_impl<2>(s, table);
    return tables_impl<3>(s, table);
}

}  // namespace pimdb
