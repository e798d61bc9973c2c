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
//      i.e. triangular linear systems: no exp, no log, no reduction on the chain. Mathematically this is the
//      reference's shifted log-sum-exp with the shift carried exactly in the binary exponent, so it is robust for
//      any positions the reference handles. They are solved 32 unknowns at a time (N <= 2048): the 32 x 32 diagonal
//      blocks are inverted ahead of the chain (positions only), the owner warp of a block turns what the earlier
//      values contribute to its rows into its 32 new values with one matrix-vector product, and every later row
//      applies the block as a plain dot product with block-scaled factor tiles (TMA bulk copies). One recurrence is
//      one thread-block cluster; new values are pushed into the other blocks' shared memory (section 3e). Larger N
//      falls back to scalar column-wise kernels in extended-range arithmetic (sections 3, 3b).
//      V[m] = -(ln W[m])/beta is recovered as the values are published.
//   4. Connection probabilities are products of known extended-range numbers,
//          P(l->u) = W[u] c(u,l) Wb[l+1] / ((l+1) W[N]),
//      evaluated on the fly in the exterior-force kernel (one warp per particle and exterior bead); the N x N
//      matrix is only materialised when a caller asks for it (pimdb_exchange_get).
#include <algorithm>

#include "internal.cuh"
#include "device_utils.cuh"

namespace pimdb {

// ---------------------------------------------------------------- extended-range positive numbers
struct Ext {
    double m;   // mantissa, in [1,2) when normalised (0 for an exact zero)
    int e;      // value = m * 2^e
};

constexpr int kExtZeroExp = -(1 << 29);
constexpr int kExchFastMaxN = 512;     // largest N for which every factor tile recomputes the prefix sums itself
constexpr int kExchBlockedMaxN = 8192; // largest N served by the block-scaled tiles and the blocked (cluster) recurrence

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
    double *Kf, *Kb;           // the same factors block-scaled for the blocked recurrence (N <= 512), or nullptr:
                               //   K[r][v] = C[r][v] * 2^-B[r/32][v] as a plain double (0 outside the triangle), stored as
                               //   32 x 32 tiles: K[(r/32 * nb + v/32) * 1024 + (r%32) * 32 + v%32], nb = ceil(N/32)
    int *Bf, *Bb;              //   B[rb][v] = largest binary exponent of C[32rb .. 32rb+31][v]
    double *Gf, *Gb;           // inverses of the 32 x 32 diagonal blocks of the two triangular systems (blocked recurrence),
                               //   G[q][k''][lane]: row k of block q lives in the lane that owns its particle row
    double *Hf, *Hb;           //   H[q][lane] = (G kin)[k]: the block's response to the value handed over by the previous block
    int *Gokf, *Gokb;          //   1 when every structural entry of G[q] is a normal double within the fast window
    int* sync;                 // device counters that let the recurrence kernel start before the factor tiles are done:
                               //   [0] tiles finished (ticket), [1] generations of tables completed, [2], [3] generations
                               //   consumed by the forward / backward recurrence block
    unsigned long long *tl0, *tl1, *tl2;   // timeline slots of the tile / recurrence / exterior-force kernels (or nullptr)
    int *statf, *statb;        //   how the last run solved each block, in step order: 1 = G rho, 2 = exact sequential steps
    double *Wm, *Wbm;          // W[0..N], Wb[0..N] mantissas
    int *We, *Wbe;             // ... exponents
    double *V, *Vb, *F;        // V[N+1], Vb[N+1], F[2][D][N]
    DevObs* obs;
    int* err;
    long long* dbg;            // optional profiling buffer (clock64 stamps per warp), nullptr in production
    int N, D, pbc, do_first, do_last;
    double k, beta, h, L, invL;   // h = beta*k/2
    // bead shard over peer memory: the other exterior bead arrives in a halo slab written by a ring neighbour; the
    // first kernels of the chain wait (bounded) for it (device_utils.cuh peer_wait_halos), nullptr otherwise
    const unsigned int* halo_flag; const unsigned int* halo_seq; unsigned long long timeout_ns;
    int tiles_diag_only;          // k_exch_coeff_tiles: one block per DIAGONAL tile (the others come from k_exch_offdiag_tiles)
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

// The Boltzmann factor the 16-byte tables of round 1 held at [r][v], recomputed from the positions (only the rarely taken
// exact paths of the blocked recurrences need extended-range factors; everything on the fast paths reads the block-scaled
// tiles): forward c(r, v), v >= r; backward c(v, r) / (r + 1), v <= r -- the same operations in the same order as the
// tile kernel, so the value is bit-identical to what that kernel scaled into its tile.
template <bool FWD>
__device__ __noinline__ int4 factor_exact(const ExArgs& a, int r, int v);

// ---------------------------------------------------------------- 1. prefix sums A(w), one block
template <int D>
__global__ void __launch_bounds__(1024) k_exch_prefix(ExArgs a) {
    __shared__ double warp_tot[32];
    __shared__ double carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    peer_wait_halos(a.halo_flag, a.halo_seq, a.timeout_ns, a.err);
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

template <bool FWD>
__device__ __noinline__ int4 factor_exact(const ExArgs& a, int r, int v) {
    const int u = FWD ? r : v, w = FWD ? v : r;        // u <= w
    double d2 = 0.0;
    for (int c = 0; c < a.D; ++c) {                    // dist2<D>(a, x1, u, xP, w) with the dimension at run time
        double dx = a.xP[(size_t)c * a.N + w] - a.x1[(size_t)c * a.N + u];
        if (a.pbc) dx = min_image(dx, a.L, a.invL);
        d2 = fma(dx, dx, d2);
    }
    const double y = a.h * (a.A[w] - a.A[u] + d2);
    Ext c = ext_exp_neg(y < 0.0 ? 0.0 : y);
    if (!FWD) c = ext_normalize(c.m * (1.0 / (double)(r + 1)), c.e);
    return ext_pack(c.m, c.e);
}

// A block-scaled tile entry is the factor divided by 2^B, flushed to 0 when it lies more than 2^1022 below the largest factor
// of its 32 rows in that column. Next to FAST blocks that loses nothing (their W stay within 2^+-400 of each other); once a
// recurrence had to solve a block exactly the weights may span more, and whoever forms products W c Wb afterwards (exterior
// forces, estimators, the probability table) recomputes the factors instead of reading the tiles. Called by every thread
// of a block.
__device__ __forceinline__ bool any_exact_block(const ExArgs& a) {
    return *(volatile const int*)&a.sync[0] != 0;      // raised by the recurrence kernels, reset by the next tile kernel
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

// ---------------------------------------------------------------- inverses of the diagonal blocks (blocked recurrence)
// Done by the diagonal tiles of k_exch_coeff_tiles. In step order (k = 0 .. n-1 within block q) the recurrence reads
//     u_k = mu_k (rho_k + sum_{k'<k} T[k][k'] u_k'),   rho = what the earlier values contribute,
// so  u = G rho  with  G = (I - diag(mu) T)^-1 diag(mu).  Every entry of T and mu is >= 0, hence G is a sum of products
// of non-negative numbers: no cancellation, and an entry that stays a normal double is accurate to a few ulp. G is
// built by recursive halving, G = [[G11, 0], [G22 T21 G11, G22]]: the four 8 x 8 diagonal blocks by substitution (28
// dependent FMAs), then two levels of small matrix products -- ~80 dependent FMAs instead of 496. Both directions at
// once (d = 0 forward, 1 backward). The block's response to the value handed over by the previous block (factor row
// of step 0, `kin`) is folded into a vector h = G kin, so the owner needs no factor of its own block at run time:
// u = G A + omega_0 h.
// Called by all 1024 threads of a diagonal tile's block; `sX` is 2 x 16 x 17 doubles of scratch shared memory.
__device__ __forceinline__ void diag_block_inverse(const ExArgs& a, int q, double kf, double kb, int maxf, int maxb,
                                                   double (*sX)[16][17]) {
    __shared__ double sT[2][32][33], sG[2][32][33];
    __shared__ double sMu[2][32], sKin[2][32];
    const int N = a.N;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    // forward: step k completes row 32q + k (tile column k) and uses factor row 32q + k (tile row k);
    // backward: steps run downwards from `top` = min(N-1, 32q+31): step k completes row top - k, factor row top - k.
    const int tt = min(N - 1, 32 * q + 31) - 32 * q;
    const int nf = min(32, N - 32 * q), nbk = tt + 1 - (q == 0 ? 1 : 0);          // backward: row 0 is never computed
    for (int i = tid; i < 2 * 32 * 33; i += 1024) { (&sT[0][0][0])[i] = 0.0; (&sG[0][0][0])[i] = 0.0; }
    if (tid < 64) { (&sMu[0][0])[tid] = 0.0; (&sKin[0][0])[tid] = 0.0; }
    __syncthreads();
    // T[k][k'] = K[factor row of step k'+1][row of step k], k' < k < n
    if (ty >= 1 && ty <= tx && tx < nf) sT[0][tx][ty - 1] = kf;                     // forward: tile (k'+1, k)
    if (ty == 0 && tx < nf) {
        sKin[0][tx] = kf;
        sMu[0][tx] = (1.0 / (double)(32 * q + tx + 1)) * pow2i(max(maxf, -1100));
    }
    if (tx <= tt && ty >= tx && ty <= tt - 1 && tt - tx < nbk) sT[1][tt - tx][tt - ty - 1] = kb;   // backward: tile (top-k'-1, top-k)
    if (ty == tt && tx <= tt && tt - tx < nbk) sKin[1][tt - tx] = kb;
    if (ty == 1 && tx <= tt && tt - tx < nbk) sMu[1][tt - tx] = pow2i(max(maxb, -1100));
    __syncthreads();
    if (tid < 64) {   // level 0: the 8 x 8 diagonal blocks, lane = (block b, column j)
        const int d = tid >> 5, b = (tid & 31) >> 3, j = tid & 7;
        double g[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double sum = (k == j) ? 1.0 : 0.0;
#pragma unroll
            for (int kp = 0; kp < k; ++kp) sum = fma(sT[d][8 * b + k][8 * b + kp], g[kp], sum);
            g[k] = (k < j) ? 0.0 : sMu[d][8 * b + k] * sum;
            sG[d][8 * b + k][8 * b + j] = g[k];
        }
    }
    __syncthreads();
    {   // level 1: 16 x 16 blocks out of 8 x 8 ones; 2 directions x 2 pairs x 64 elements
        const int d = tid >> 7, pr = (tid >> 6) & 1, i = (tid >> 3) & 7, j = tid & 7;
        const int r0 = 16 * pr, r1 = r0 + 8;
        if (tid < 256) {
            double x = 0.0;
#pragma unroll
            for (int m = 0; m < 8; ++m) x = fma(sT[d][r1 + i][r0 + m], sG[d][r0 + m][r0 + j], x);
            sX[d][8 * pr + i][j] = x;
        }
        __syncthreads();
        if (tid < 256) {
            double x = 0.0;
#pragma unroll
            for (int m = 0; m < 8; ++m) x = fma(sG[d][r1 + i][r1 + m], sX[d][8 * pr + m][j], x);
            sG[d][r1 + i][r0 + j] = x;
        }
    }
    __syncthreads();
    {   // level 2: the 32 x 32 block; 2 directions x 256 elements
        const int d = tid >> 8, i = (tid >> 4) & 15, j = tid & 15;
        if (tid < 512) {
            double x0 = 0.0, x1 = 0.0;
#pragma unroll
            for (int m = 0; m < 16; m += 2) {
                x0 = fma(sT[d][16 + i][m], sG[d][m][j], x0);
                x1 = fma(sT[d][16 + i][m + 1], sG[d][m + 1][j], x1);
            }
            sX[d][i][j] = x0 + x1;
        }
        __syncthreads();
        if (tid < 512) {
            double x0 = 0.0, x1 = 0.0;
#pragma unroll
            for (int m = 0; m < 16; m += 2) {
                x0 = fma(sG[d][16 + i][16 + m], sX[d][m][j], x0);
                x1 = fma(sG[d][16 + i][16 + m + 1], sX[d][m + 1][j], x1);
            }
            sG[d][16 + i][j] = x0 + x1;
        }
    }
    __syncthreads();
    // validity: every structural entry (k'' <= k < n) and h_k must be a normal positive double in [2^-700, 2^300]
    auto in_window = [](double g) { return (unsigned)__double2hiint(g) - (323u << 20) < (1000u << 20); };
    bool okf = !(tx <= ty && ty < nf) || in_window(sG[0][ty][tx]);
    bool okb = !(tx <= ty && ty < nbk) || in_window(sG[1][ty][tx]);
    if (tid < 64) {   // h = G kin, lane = row k; the result goes to the lane that owns row k in the recurrence
        const int d = tid >> 5, k = tid & 31;
        double h0 = 0.0, h1 = 0.0;
#pragma unroll 8
        for (int c = 0; c < 32; c += 2) {
            h0 = fma(sG[d][k][c], sKin[d][c], h0);
            h1 = fma(sG[d][k][c + 1], sKin[d][c + 1], h1);
        }
        const double h = h0 + h1;
        const int n = d == 0 ? nf : nbk;
        if (k < n && !in_window(h)) { if (d == 0) okf = false; else okb = false; }
        const int ln = d == 0 ? k : (tt - k) & 31;
        (d == 0 ? a.Hf : a.Hb)[q * 32 + ln] = h;
    }
    okf = __syncthreads_and(okf);
    okb = __syncthreads_and(okb);
    {   // G[q][k''][lane]: row k lives in the lane that owns its particle row (rows k >= n: the lanes without a row, zeros)
        const int c = ty, ln = tx;
        a.Gf[(size_t)q * 1024 + c * 32 + ln] = sG[0][ln][c];
        a.Gb[(size_t)q * 1024 + c * 32 + ln] = sG[1][(tt - ln) & 31][c];
    }
    if (tid == 0) { a.Gokf[q] = okf ? 1 : 0; a.Gokb[q] = okb ? 1 : 0; }
}

// Tile version for N <= 512: one block per 32 x 32 tile of the index square also emits the block-scaled copy the
// fast recurrence consumes -- per (32-row block rb, column v) the largest binary exponent B and the factors as plain
// doubles relative to 2^B. A factor more than 2^1022 below its block's largest becomes 0; it multiplies values that
// the fast recurrence keeps within 2^+-400 of each other, so it could not have contributed.
template <int D>
// (2 blocks per SM, 32 registers: all 256 tiles of N = 512 are dispatched in ONE wave. The block scheduler does not
// start the pair-tile grid, launched right behind, before every block of this grid has been dispatched -- with one block
// per SM the pair tiles started 6 us late, measured with the in-kernel timeline.)
__global__ void __launch_bounds__(1024, 2) k_exch_coeff_tiles(ExArgs a) {
    __shared__ __align__(16) int s_e2[2][32][33];  // [step within the tile][column], padded: conflict-free both ways
    int (*s_ef)[33] = s_e2[0], (*s_eb)[33] = s_e2[1];
    __shared__ int s_mf[32], s_mb[32];
    __shared__ double sA[kExchFastMaxN + 1];       // the prefix sums A(w), recomputed by every tile (N <= 512: one chunk)
    __shared__ double warp_tot[32];
    grid_dependency_wait();      // (a captured step launches this grid early behind the integrator kernel that moves the beads)
    tl_begin(a.tl0);
    grid_launch_dependents();    // the recurrence kernel may take its SMs now; it waits for this grid before it reads the tiles
    if (blockIdx.x == 0 && threadIdx.x == 0) a.sync[0] = 0;   // "a block of these tables was solved exactly": raised by the recurrences
    peer_wait_halos(a.halo_flag, a.halo_seq, a.timeout_ns, a.err);
    const int N = a.N, nb = (N + 31) >> 5;
    const int rb = a.tiles_diag_only ? blockIdx.x : blockIdx.x / nb, sb = a.tiles_diag_only ? blockIdx.x : blockIdx.x % nb;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const bool own_prefix = N <= kExchFastMaxN;      // beyond that k_exch_prefix has run before this kernel
    if (own_prefix) {
        // Same operations in the same order as k_exch_prefix (so A is bit-identical); fusing it here takes a 5 us
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
    if (r < N && sc < N) {
        const int u = min(r, sc), v = max(r, sc);
        const double Av = own_prefix ? sA[v] : a.A[v], Au = own_prefix ? sA[u] : a.A[u];
        const double y = a.h * (Av - Au + dist2<D>(a, a.x1, u, a.xP, v));
        const Ext c = ext_exp_neg(y < 0.0 ? 0.0 : y);   // (a NaN position stays NaN and is reported, like the reference)
        // (the factors live on only as block-scaled tiles + block exponents: 8 bytes instead of 24 per entry, which is what
        // bounds this kernel from N ~ 2000 on; the exact paths of the recurrences recompute what they need, factor_exact)
        // (a.Cf: up to N = 512 the 16-byte entries are kept as well -- 8 MB, L2 resident -- because the exterior forces read a
        // contiguous row of them 1.7 us faster than the same row spread over 16 tiles)
        if (sc >= r) {
            mf = c.m; ef = c.e;
            if (a.Cf) a.Cf[(long long)r * N + sc] = ext_pack(c.m, c.e);
        }
        if (sc <= r) {   // backward: the 1/(p+1) weight of the sum (p = r) is folded in here, off the chain
            const Ext cb = ext_normalize(c.m * (1.0 / (double)(r + 1)), c.e);
            mb = cb.m; eb = cb.e;
            if (a.Cb) a.Cb[(long long)r * N + sc] = ext_pack(cb.m, cb.e);
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
    // (NaN mantissas propagate; an exact zero has exponent kExtZeroExp and scales to 0)
    const double kf = mf * pow2i(max(ef - maxf, -1100)), kb = mb * pow2i(max(eb - maxb, -1100));
    if (sc < N && ty == 0) {
        a.Bf[rb * N + sc] = maxf;
        a.Bb[rb * N + sc] = maxb;
    }
    // tiled layout: the 32 x 32 tile (factor-row block rb, accumulating-row block sb) is one contiguous 8 KB piece
    // [factor row][accumulating row], so that a warp of the recurrence fetches it with two bulk copies
    {
        // (tiles on the wrong side of the diagonal are never fetched -- the forward recurrence reads tiles with sb >= rb,
        // the backward one tiles with sb <= rb -- and stay as allocated: zero)
        const size_t t = ((size_t)rb * nb + sb) * 1024 + ty * 32 + tx;
        if (sb >= rb) a.Kf[t] = kf;
        if (sb <= rb) a.Kb[t] = kb;
    }
    if (rb == sb) {
        __syncthreads();                             // s_e2 is free now: it becomes the scratch of the inversion
        static_assert(sizeof(s_e2) >= 2 * 16 * 17 * sizeof(double), "scratch too small");
        diag_block_inverse(a, rb, kf, kb, maxf, maxb, reinterpret_cast<double (*)[16][17]>(&s_e2[0][0][0]));
    }
    tl_end(a.tl0);
}

// Off-diagonal factor tiles for N > 512, where the tile grid is many waves long and what bounds it is the latency of one
// tile times the tiles in flight per SM (measured at N = 8192: 65536 blocks of 1024 threads, two per SM, 3.3 us each =
// 730 us for 0.55 GB of output). Here a tile is a block of 256 threads, four rows per thread -- four independent exp()
// chains per thread and up to eight tiles in flight per SM. A tile above the diagonal only feeds the forward recurrence,
// one below it only the backward one; the diagonal tiles, which also invert their block, stay with k_exch_coeff_tiles.
template <int D>
__global__ void __launch_bounds__(256) k_exch_offdiag_tiles(ExArgs a) {
    __shared__ int s_max[8][32];
    const int N = a.N, nb = (N + 31) >> 5;
    const int rb = blockIdx.x / nb, sb = blockIdx.x % nb;
    if (rb == sb) return;
    const bool fwd = sb > rb;
    const int tx = threadIdx.x & 31, tq = threadIdx.x >> 5;
    const int sc = sb * 32 + tx;
    double m[4];
    int e[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = rb * 32 + tq * 4 + k;
        m[k] = 0.0; e[k] = kExtZeroExp;
        if (r < N && sc < N) {
            const int u = min(r, sc), v = max(r, sc);
            const double y = a.h * (a.A[v] - a.A[u] + dist2<D>(a, a.x1, u, a.xP, v));
            Ext c = ext_exp_neg(y < 0.0 ? 0.0 : y);
            if (!fwd) c = ext_normalize(c.m * (1.0 / (double)(r + 1)), c.e);
            m[k] = c.m; e[k] = c.e;
        }
    }
    s_max[tq][tx] = max(max(e[0], e[1]), max(e[2], e[3]));
    __syncthreads();
    int mx = s_max[0][tx];
#pragma unroll
    for (int q = 1; q < 8; ++q) mx = max(mx, s_max[q][tx]);
    double* K = (fwd ? a.Kf : a.Kb) + ((size_t)rb * nb + sb) * 1024;
#pragma unroll
    for (int k = 0; k < 4; ++k) K[(tq * 4 + k) * 32 + tx] = m[k] * pow2i(max(e[k] - mx, -1100));
    if (tq == 0 && sc < N) (fwd ? a.Bf : a.Bb)[rb * N + sc] = mx;
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

// (am, ae) += acc * 2^e for a plain non-negative double acc; out of line: it runs once per 32 columns
static __device__ __noinline__ void ext_fold(double& am, int& ae, double acc, int e) {
    if (acc > 0.0 || acc != acc) {
        const Ext t = ext_normalize(acc, e);
        const int emax = max(ae, t.e);
        am = fma(am, pow2i(max(ae - emax, -1100)), t.m * pow2i(max(t.e - emax, -1100)));
        ae = emax;
    }
}

// ---------------------------------------------------------------- 3d. blocked recurrence: the scheme (kernels in 3e / 3f)
// The two recurrences are triangular linear systems. The scalar kernel (recur_body) walks them one unknown at a time: N dependent
// steps of ~100-200 cycles. Here the system is solved 32 unknowns at a time: k_exch_coeff_tiles has already inverted
// every 32 x 32 diagonal block (G and h, positions only -- off the chain), so the owner warp of block q turns "what
// the earlier values contribute to my 32 rows" (A, one plain double per lane) into its 32 new values with ONE 32 x 32
// matrix-vector product, u = G A + omega_0 h: lane k keeps row k of G in registers, A travels through 256 bytes of
// shared memory. The dependency chain is N/32 block steps instead of N scalar steps.
//   * value #s in step order is W[s] (forward) / Wb[N-s] (backward); block q covers the steps [lo_q, hi_q] whose
//     completing rows are 32q .. 32q+31, consumes the values #lo_q .. #hi_q and produces #lo_q+1 .. #hi_q+1. A value
//     is stored under the factor row it multiplies (r = s forward, N-1-s backward), so block q's 32 entries are
//     sOm[32q .. 32q+31] in either direction.
//   * the owner publishes its block as plain doubles omega relative to ONE binary exponent E_q (sOm / sEx), hands its
//     last value to the next owner, fences, and raises the block's flag; every later warp then applies the
//     block's 32 columns to its own rows as a plain dot product with the block-scaled factor tile K[q][warp] and folds
//     it into its extended-range accumulator once.
//   * factor tiles are contiguous 8 KB pieces (k_exch_coeff_tiles), fetched by TMA bulk copies (cp.async.bulk +
//     mbarrier transaction counts) as 4 KB half tiles into a private 3-slot ring per warp: one copy instruction per
//     16 columns instead of 16 LDGSTS (8 cycles each at the SM's load/store unit -- that, not the arithmetic, bounded
//     the per-thread cp.async ring of the scalar kernel at ~4000 cycles per block, measured).
//   * fast-path validity, checked where it is cheap: G's structural entries and h are normal doubles in
//     [2^-700, 2^300] (flag from the tile kernel), the accumulators entering the block are <= 2^600 on the block's
//     scale, every new omega is a normal double in [2^-700, 2^300]. All terms are non-negative, so whatever underflows
//     on the way is at least 2^-322 below the result it was added to. If any check fails the block is redone exactly:
//     sequential extended-range steps from the untouched accumulators (the recur_body arithmetic), flag value 2,
//     and the consumers apply such a block column by column from the extended-range table.
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    } while (!ok);
}
// one 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

// ---------------------------------------------------------------- 3e. blocked recurrence on a thread-block cluster
// With all consumer warps of a direction on ONE SM (round 1's single-block kernel, since removed) every factor crosses
// that SM's shared-memory pipe twice (TMA write + LDS read, ~2200 cycles per block with 15 consumers; measured with
// clock64 stamps) and the owner's own shared-memory traffic queues behind it, so a block step costs ~2700 cycles although
// the chain itself needs ~900. Here one recurrence is a CLUSTER of 8 thread blocks (one or two warps each, 8 SMs): each block keeps
// the rings, accumulators and G rows of its own row blocks, and a full copy of the published values. The owner of
// row block q writes its 32 new values straight into the shared memory of every block that still needs them
// (st.shared::cluster over the SM-to-SM network, ~215 cycles), orders them with one cluster-scope fence and raises
// the block's flag in each of those copies; consumers poll their LOCAL flag. Per SM the factor traffic drops 8-fold,
// so the chain is what is left. Small blocks also fit wherever a pair-tile block retires.
// Arithmetic, validity checks and the exact fallback are those described in 3d.
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned map_to_cta(const void* smem_ptr, unsigned cta) {
    unsigned a = (unsigned)__cvta_generic_to_shared(smem_ptr), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_cluster_f64(unsigned addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_s32(unsigned addr, int v) {
    asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cluster_s32(const int* p) {
    int r;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.acquire.cluster.shared::cta.s32 %0, [%1];" : "=r"(r) : "r"(a) : "memory");
    return r;
}

constexpr int kClusterSize = 8;

__device__ __forceinline__ void st_cluster_b64(unsigned addr, unsigned long long v) {
    asm volatile("st.relaxed.cluster.shared::cluster.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long lds_volatile_b64(const void* p) {
    unsigned long long r;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.b64 %0, [%1];" : "=l"(r) : "r"(a) : "memory");
    return r;
}

// Publication protocol (no fence on the chain): every published word is 8 bytes, naturally aligned -- single-copy atomic
// -- and carries its own validity: an omega of a fast block is a strictly positive double (0 = not there yet), the
// block word is {mode, E} with mode != 0. A consumer waits for the block word, then for the block's omegas (one lane
// per entry, one vote), and needs no ordering between different words. Only an EXACT block (per-value exponents in
// separate words) orders its stores with a cluster-scope fence before it raises its block word.
template <bool FWD>
__device__ __forceinline__ void recur_cluster(const ExArgs& a, double* smem_d) {
    constexpr int SLOTS = 3, HALF = 512;
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int nt = blockDim.x, N = a.N, lane = tid & 31, lw = tid >> 5, wpc = nt >> 5;
    const int nb = (N + 31) >> 5;
    const int crank = (int)cluster_ctarank();
    const int warp = crank * wpc + lw;               // = the row block this warp owns (if < nb)
    const bool has_block = warp < nb;
    const int nsteps = FWD ? N : N - 1;
    const int npos = nsteps > 0 ? nb : 0;
    const int nb2 = (nb + 3) & ~1;
    double* ring = smem_d;
    double* sOm = ring + (size_t)wpc * SLOTS * HALF;                 // value table, by factor row; 0 = not published
    double* sRho = sOm + 32 * nb;
    double* sHandOm = sRho + 32 * wpc;                               // hand-off value per position; 0 = not published
    unsigned long long* sWord = reinterpret_cast<unsigned long long*>(sHandOm + nb2);   // block word per position
    unsigned long long* sBar = sWord + nb2;
    int* sEx = reinterpret_cast<int*>(sBar + wpc * SLOTS + (wpc & 1));   // per-value exponents (exact blocks only)
    int* sHandE = sEx + 32 * nb;

    auto g_lo = [&](int q) { return FWD ? 32 * q : max(0, N - 32 * (q + 1)); };
    auto g_hi = [&](int q) { return min(nsteps - 1, FWD ? 32 * q + 31 : N - 1 - 32 * q); };
    auto row_of = [&](int st) { return FWD ? st : (N - 1 - st); };
    const int v = 32 * warp + lane;                                 // my row
    const bool row_ok = has_block && (FWD ? (v < N) : (v >= 1 && v < N));
    const int own_lo = has_block ? g_lo(warp) : 0, own_hi = has_block ? g_hi(warp) : -1, n_own = own_hi - own_lo + 1;
    const int mypos = FWD ? warp : nb - 1 - warp;
    const int last_need = row_ok ? (FWD ? v : N - 1 - v) : -1;
    auto need = [&](int s) { return s <= last_need; };
    const int* Bg = FWD ? a.Bf : a.Bb;
    const double* Ktile = (FWD ? a.Kf : a.Kb) + (size_t)warp * 1024;
    const size_t tile_stride = (size_t)nb * 1024;
    const int ntile_half = has_block ? 2 * min(mypos, npos) : 0;
    double* const ring_w = ring + (size_t)lw * SLOTS * HALF;
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring_w);
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(sBar + lw * SLOTS);
    auto issue_half = [&](int i) {
        const int pos = i >> 1, q = FWD ? pos : nb - 1 - pos, slot = i % SLOTS;
        bulk_load(ring_s + slot * HALF * 8, Ktile + (size_t)q * tile_stride + (i & 1) * HALF, HALF * 8, bar_s + slot * 8);
    };
    // the blocks that still need the values of my row block: those holding a later position (my own included)
    const int dst_lo = FWD ? crank : 0, dst_hi = FWD ? kClusterSize - 1 : crank;

    for (int i = tid; i < 32 * nb; i += nt) { sOm[i] = 0.0; sEx[i] = 0; }
    for (int i = tid; i < nb2; i += nt) { sWord[i] = 0ull; sHandOm[i] = 0.0; sHandE[i] = 0; }
    __syncthreads();
    if (tid == 0) {
        sHandOm[0] = 1.0;
    }
    grid_dependency_wait();     // the tiles / block inverses of k_exch_coeff_tiles (programmatic stream serialisation, see above)
    __syncthreads();
    if (lane == 0 && has_block) {
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) mbar_init(bar_s + k * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
        for (int i = 0; i < SLOTS && i < ntile_half; ++i) issue_half(i);
    }
    double Grow[32];
    double hrow = 0.0;
    int Gok = 0, Bown = kExtZeroExp;
    if (has_block) {
        const double* Gg = (FWD ? a.Gf : a.Gb) + (size_t)warp * 1024 + lane;
#pragma unroll
        for (int k = 0; k < 32; ++k) Grow[k] = Gg[k * 32];
        hrow = (FWD ? a.Hf : a.Hb)[warp * 32 + lane];
        Gok = (FWD ? a.Gokf : a.Gokb)[warp];
        if (row_ok) Bown = Bg[warp * N + v];
    } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) Grow[k] = 0.0;
    }
    cluster_sync_all();                              // every copy is initialised before anybody writes into it

    double am = 0.0;
    int ae = kExtZeroExp;
    long long* stamp = (a.dbg && lane == 0 && has_block && nb <= 16) ? a.dbg + 192 + ((FWD ? 0 : 16) + warp) * 64 : nullptr;
    if (stamp) stamp[5] = clock64();                                  // start (after the cluster barrier)
    int Eprev = 0, mode_prev = 1;                                     // block word of the block before mine
    // ---- consumer phases
    int Bnext = (has_block && mypos > 0 && row_ok) ? Bg[(FWD ? 0 : nb - 1) * N + v] : 0;
#pragma unroll 1
    for (int pos = 0; has_block && pos < mypos && pos < npos; ++pos) {
        const int q = FWD ? pos : nb - 1 - pos;
        const int Bq = Bnext;
        if (pos + 1 < mypos && row_ok) Bnext = Bg[(FWD ? pos + 1 : nb - 2 - pos) * N + v];
        unsigned long long word;
        do { word = lds_volatile_b64(&sWord[pos]); } while (word == 0ull);
        const int fl = (int)(word >> 32), Eq = (int)(unsigned)word;
        Eprev = Eq; mode_prev = fl;
        const double* om = sOm + 32 * q;
        if (fl == 1) {   // lane k vouches for entry k of the block (entries without a value stay 0 and multiply K = 0)
            const int r = 32 * q + lane;
            const bool expect = r >= (FWD ? 0 : 1) && r < N;
            while (!__all_sync(kFullMask, !expect || lds_volatile_b64(&om[lane]) != 0ull)) {}
        } else {
            asm volatile("fence.acq_rel.cluster;" ::: "memory");
        }
        if (stamp && pos + 1 == mypos) stamp[0] = clock64();          // the previous block's values are here
        const int i0 = 2 * pos;
        mbar_wait(bar_s + (i0 % SLOTS) * 8, (unsigned)(i0 / SLOTS) & 1u);
        mbar_wait(bar_s + ((i0 + 1) % SLOTS) * 8, (unsigned)((i0 + 1) / SLOTS) & 1u);
        if (fl == 1) {
            double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double* kp = ring_w + ((i0 + h) % SLOTS) * HALF + lane;
                const double* op = om + 16 * h;
#pragma unroll
                for (int c = 0; c < 16; c += 4) {
                    const double2 w01 = *reinterpret_cast<const double2*>(op + c);
                    const double2 w23 = *reinterpret_cast<const double2*>(op + c + 2);
                    acc0 = fma(kp[(c + 0) * 32], w01.x, acc0);
                    acc1 = fma(kp[(c + 1) * 32], w01.y, acc1);
                    acc2 = fma(kp[(c + 2) * 32], w23.x, acc2);
                    acc3 = fma(kp[(c + 3) * 32], w23.y, acc3);
                }
            }
            ext_fold(am, ae, (acc0 + acc1) + (acc2 + acc3), Bq + Eq);
        } else {
            const int s0 = g_lo(q), s1 = g_hi(q);
#pragma unroll 1
            for (int s = s0; s <= s1; ++s) {
                if (need(s)) {
                    const int r = row_of(s);
                    const int4 c = factor_exact<FWD>(a, r, v);
                    ext_fma(am, ae, ext_m(c), c.z, sOm[r], sEx[r]);
                }
            }
        }
        __syncwarp();
        if (lane == 0 && i0 + SLOTS < ntile_half) {   // both slots are free again (the warp has read them: __syncwarp above)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_half(i0 + SLOTS);
            if (i0 + 1 + SLOTS < ntile_half) issue_half(i0 + 1 + SLOTS);
        }
    }
    if (!row_ok) { am = 0.0; ae = kExtZeroExp; }
    if (stamp) stamp[1] = clock64();                                  // every earlier block applied
    // ---- owner phase
    if (n_own > 0) {
        // value #s goes to every block that needs it: s <= own_hi into the value table under its factor row, the
        // block's last value into the hand-off slot of the next owner
        auto publish = [&](int s, double m) {
            const void* pm = (s == own_hi + 1) ? (const void*)&sHandOm[mypos + 1] : (const void*)&sOm[row_of(s)];
            for (int d = dst_lo; d <= dst_hi; ++d) st_cluster_b64(map_to_cta(pm, (unsigned)d), (unsigned long long)__double_as_longlong(m));
        };
        auto publish_exp = [&](int s, int e) {                        // exact blocks: per-value exponent
            const void* pe = (s == own_hi + 1) ? (const void*)&sHandE[mypos + 1] : (const void*)&sEx[row_of(s)];
            for (int d = dst_lo; d <= dst_hi; ++d) st_cluster_s32(map_to_cta(pe, (unsigned)d), e);
        };
        auto publish_word = [&](int mode, int E) {
            const unsigned long long w = ((unsigned long long)(unsigned)mode << 32) | (unsigned long long)(unsigned)E;
            for (int d = dst_lo; d <= dst_hi; ++d) st_cluster_b64(map_to_cta(&sWord[mypos], (unsigned)d), w);
        };
        // the same value for the force kernel and the caller: W / Wb (normalised) and V = -(ln W)/beta
        auto emit = [&](int s, double m, int e) {
            const Ext wn = ext_normalize(m, e);
            const int i = FWD ? s : N - s;
            (FWD ? a.Wm : a.Wbm)[i] = wn.m;
            (FWD ? a.We : a.Wbe)[i] = wn.e;
            const double val = -(log(wn.m) + (double)wn.e * 0.6931471805599453) / a.beta;
            if (!isfinite(val)) atomicOr(a.err, FWD ? kErrOverflowFwd : kErrOverflowBwd);
            (FWD ? a.V : a.Vb)[i] = (s == 0) ? 0.0 : val;
        };
        // the value handed over by the previous owner (its exponent: the block word's E, or its own word after an exact block)
        unsigned long long hbits;
        do { hbits = lds_volatile_b64(&sHandOm[mypos]); } while (hbits == 0ull);
        const double hm = __longlong_as_double((long long)hbits);
        const int he = mypos == 0 ? 0 : (mode_prev == 1 ? Eprev : sHandE[mypos]);
        const Ext n0 = ext_normalize(hm, he);
        const int E = n0.e;
        const double om0 = n0.m;
        const int d = ae - E - Bown;
        const double A = (am == 0.0) ? 0.0 : ext_to_double(am, min(d, 600));
        bool exact = __any_sync(kFullMask, (am != 0.0) && (d > 600)) || !(om0 > 0.0) || !Gok;
        const int tt = FWD ? 0 : row_of(own_lo) & 31;
        const int kslot = FWD ? lane : (tt - lane) & 31;
        int mode = 2;
        if (!exact) {
            double* rho_w = sRho + lw * 32;
            rho_w[kslot] = A;
            __syncwarp();
            double u0 = om0 * hrow, u1 = 0.0, u2 = 0.0, u3 = 0.0;
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
                const double2 r01 = *reinterpret_cast<const double2*>(rho_w + k);
                const double2 r23 = *reinterpret_cast<const double2*>(rho_w + k + 2);
                u0 = fma(Grow[k], r01.x, u0);
                u1 = fma(Grow[k + 1], r01.y, u1);
                u2 = fma(Grow[k + 2], r23.x, u2);
                u3 = fma(Grow[k + 3], r23.y, u3);
            }
            const double u = (u0 + u1) + (u2 + u3);
            const bool mine = kslot < n_own;
            const unsigned hi = (unsigned)__double2hiint(u);
            const bool okall = __all_sync(kFullMask, !mine || (hi - (323u << 20) < (1000u << 20)));
            if (stamp) stamp[2] = clock64();                          // new values computed
            if (okall) {
                // destination by destination, the next owner's block first: it is the only one in a hurry
                const void* pm = !mine ? nullptr : (kslot == n_own - 1) ? (const void*)&sHandOm[mypos + 1]
                                                                       : (const void*)&sOm[row_of(own_lo + kslot + 1)];
                const unsigned long long ub = (unsigned long long)__double_as_longlong(u);
                const unsigned long long ob = (unsigned long long)__double_as_longlong(om0);
                const unsigned long long wb = (1ull << 32) | (unsigned long long)(unsigned)E;
                for (int k = 0; k <= dst_hi - dst_lo; ++k) {
                    const unsigned dd = (unsigned)(FWD ? dst_lo + k : dst_hi - k);
                    if (mine) st_cluster_b64(map_to_cta(pm, dd), ub);
                    if (lane == 0) {
                        st_cluster_b64(map_to_cta(&sOm[row_of(own_lo)], dd), ob);
                        st_cluster_b64(map_to_cta(&sWord[mypos], dd), wb);
                    }
                    if (stamp && k == 0) stamp[3] = clock64();        // the next owner's copy is on its way
                }
                if (stamp) stamp[4] = clock64();                      // all remote stores issued
                mode = 1;
                if (mine) emit(own_lo + kslot + 1, u, E);
                if (lane == 0 && mypos == 0) emit(0, om0, E);
            } else {
                exact = true;
            }
        }
        if (exact) {
            double wm = n0.m;
            int we = n0.e;
            if (lane == 0) {
                publish(own_lo, wm);
                publish_exp(own_lo, we);
                if (mypos == 0) emit(0, wm, we);
            }
#pragma unroll 1
            for (int st = own_lo; st <= own_hi; ++st) {
                if (need(st)) {
                    const int4 c = factor_exact<FWD>(a, row_of(st), v);
                    ext_fma(am, ae, ext_m(c), c.z, wm, we);
                }
                const int lane_o = row_of(st) & 31;
                const Ext fin = ext_normalize(FWD ? am * a.Inv[st + 1] : am, ae);
                wm = __shfl_sync(kFullMask, fin.m, lane_o);
                we = __shfl_sync(kFullMask, fin.e, lane_o);
                if (lane == lane_o) {
                    publish_exp(st + 1, fin.e);
                    publish(st + 1, fin.m);                           // (a zero here is caught as a non-finite V below)
                    emit(st + 1, fin.m, fin.e);
                }
            }
            __syncwarp();
            if (lane == 0) {
                asm volatile("fence.acq_rel.cluster;" ::: "memory");
                publish_word(2, 0);
            }
        }
        if (lane == 0) { (FWD ? a.statf : a.statb)[mypos] = mode; if (mode == 2) atomicOr(&a.sync[0], 1); }
    }
    if (has_block && n_own <= 0 && lane == 0) {
        (FWD ? a.statf : a.statb)[mypos] = 0;
        if (nsteps == 0 && warp == 0) {                 // backward recurrence of a single particle: Wb[1] = 1
            a.Wbm[N] = 1.0; a.Wbe[N] = 0; a.Vb[N] = 0.0;
        }
    }
    cluster_sync_all();                              // nobody leaves while its shared memory may still be written
    if (crank == 0 && tid == 0) a.sync[FWD ? 2 : 3] += 1;
}

// ---------------------------------------------------------------- 3f. the same, several row blocks per warp (N <= 8192)
// Beyond 64 row blocks a direction still runs on ONE cluster of 8 thread blocks x 8 warps, and warp g owns the row
// blocks g, g + 64, g + 128, ... (up to 4): consecutive blocks belong to consecutive warps, so the chain hops from
// warp to warp exactly as before, and every warp keeps consuming for the blocks it has not solved yet. Per published
// block a warp applies up to 4 factor tiles (one per block it still owns), the one of its NEXT own block first. The
// row of G of the next own block is fetched right after the previous own block has been solved (64 chain steps of
// slack). Shared memory per thread block: 96 KB of tile rings + 64 KB for the value table of N = 8192.
constexpr int kMultiM = 4;        // row blocks per warp
constexpr int kMultiWpc = 8;      // warps per thread block (8 blocks per cluster: 64 owner warps per direction)

template <bool FWD>
__device__ __forceinline__ void recur_cluster_multi(const ExArgs& a, double* smem_d) {
    constexpr int SLOTS = 3, HALF = 512, WPC = kMultiWpc, NW = kClusterSize * kMultiWpc, M = kMultiM;
    constexpr int kSpinMax = 1 << 26;            // every wait is bounded: a protocol error becomes an error code, not a hang
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int nt = blockDim.x, N = a.N, lane = tid & 31, lw = tid >> 5;
    const int nb = (N + 31) >> 5;
    const int crank = (int)cluster_ctarank();
    const int gwarp = crank * WPC + lw;
    const int nsteps = FWD ? N : N - 1;
    const int npos = nb;                                             // (N > 2048: every block owns steps)
    const int nb2 = (nb + 3) & ~1;
    double* ring = smem_d;
    double* sOm = ring + (size_t)WPC * SLOTS * HALF;
    double* sRho = sOm + 32 * nb;
    double* sHandOm = sRho + 32 * WPC;
    unsigned long long* sWord = reinterpret_cast<unsigned long long*>(sHandOm + nb2);
    unsigned long long* sBar = sWord + nb2;
    int* sEx = reinterpret_cast<int*>(sBar + WPC * SLOTS + (WPC & 1));
    int* sHandE = sEx + 32 * nb;

    auto g_lo = [&](int q) { return FWD ? 32 * q : max(0, N - 32 * (q + 1)); };
    auto g_hi = [&](int q) { return min(nsteps - 1, FWD ? 32 * q + 31 : N - 1 - 32 * q); };
    auto row_of = [&](int st) { return FWD ? st : (N - 1 - st); };
    auto pos_of = [&](int q) { return FWD ? q : nb - 1 - q; };
    auto blk_at = [&](int pos) { return FWD ? pos : nb - 1 - pos; };
    const int* Bg = FWD ? a.Bf : a.Bb;
    const double* Kall = FWD ? a.Kf : a.Kb;

    // my row blocks q_j = gwarp + j NW and my row in each of them
    auto q_of = [&](int j) { return gwarp + j * NW; };
    auto own_valid = [&](int j) { return q_of(j) < nb; };
    auto row_v = [&](int j) { return 32 * q_of(j) + lane; };
    auto row_ok = [&](int j) { const int v = row_v(j); return own_valid(j) && (FWD ? v < N : (v >= 1 && v < N)); };

    // Consumption order: my row blocks one after the other in step order, each taking ALL earlier positions -- the
    // positions older than my previous own block arrive as a burst from the table (they were published long ago), the
    // rest as they are published. Only one accumulator is live, and the warp that is next on the chain has exactly one
    // tile to apply between the previous block's publication and its own owner phase. The ring runs SLOTS half tiles
    // ahead of the same enumeration.
    struct Cur { int k, pos; };      // k-th of my blocks in step order, source position
    auto jth = [&](int k) { return FWD ? k : M - 1 - k; };          // forward: j ascending, backward: descending
    auto cur_norm = [&](Cur& c) {
        while (c.k < M) {
            const int j = jth(c.k);
            if (own_valid(j) && c.pos < pos_of(q_of(j))) return;
            ++c.k; c.pos = 0;
        }
    };
    auto cur_next = [&](Cur& c) { ++c.pos; cur_norm(c); };
    double* const ring_w = ring + (size_t)lw * SLOTS * HALF;
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring_w);
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(sBar + lw * SLOTS);
    Cur prod{0, 0};                  // next tile to fetch
    int prod_half = 0, n_issued = 0;
    auto issue_one = [&]() {         // fetch the next half tile into slot n_issued % SLOTS (every lane runs the cursor, lane 0 copies)
        if (prod.k >= M) return false;
        if (lane == 0) {
            const int qs = blk_at(prod.pos), qd = q_of(jth(prod.k)), slot = n_issued % SLOTS;
            bulk_load(ring_s + slot * HALF * 8, Kall + ((size_t)qs * nb + qd) * 1024 + prod_half * HALF, HALF * 8, bar_s + slot * 8);
        }
        ++n_issued;
        if (++prod_half == 2) { prod_half = 0; cur_next(prod); }
        return true;
    };

    for (int i = tid; i < 32 * nb; i += nt) { sOm[i] = 0.0; sEx[i] = 0; }
    for (int i = tid; i < nb2; i += nt) { sWord[i] = 0ull; sHandOm[i] = 0.0; sHandE[i] = 0; }
    __syncthreads();
    if (tid == 0) {
        sHandOm[0] = 1.0;
    }
    grid_dependency_wait();     // the tiles / block inverses of k_exch_coeff_tiles (programmatic stream serialisation, see above)
    __syncthreads();
    cur_norm(prod);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) mbar_init(bar_s + k * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncwarp();
    for (int k = 0; k < SLOTS; ++k) if (!issue_one()) break;
    // G row / h / flags of my first own block in step order
    double Grow[32];
    double hrow = 0.0;
    int Gok = 0, Bown = kExtZeroExp, jown = -1;
    auto load_own = [&](int j) {     // prepare the owner phase of block q_j
        jown = j;
        const int q = q_of(j);
        const double* Gg = (FWD ? a.Gf : a.Gb) + (size_t)q * 1024 + lane;
#pragma unroll
        for (int k = 0; k < 32; ++k) Grow[k] = Gg[k * 32];
        hrow = (FWD ? a.Hf : a.Hb)[q * 32 + lane];
        Gok = (FWD ? a.Gokf : a.Gokb)[q];
        Bown = row_ok(j) ? Bg[(size_t)q * N + row_v(j)] : kExtZeroExp;
    };
#pragma unroll
    for (int k = 0; k < 32; ++k) Grow[k] = 0.0;
    cluster_sync_all();

    int n_consumed = 0;              // half tiles taken out of the ring so far
#pragma unroll 1
    for (int kk = 0; kk < M; ++kk) {
        const int j = jth(kk);
        if (!own_valid(j)) continue;
        const int q = q_of(j), pos = pos_of(q);                      // my block and its position on the chain
        load_own(j);                                                 // row of G, h, flags (latency hidden by the loop below)
        const int v = row_v(j);
        const bool rok = row_ok(j);
        double am = 0.0;
        int ae = kExtZeroExp;
        // ---- everything the earlier positions contribute to my rows
        int Bnext = (pos > 0 && rok) ? Bg[(size_t)blk_at(0) * N + v] : kExtZeroExp;
#pragma unroll 1
        for (int sp = 0; sp < pos; ++sp) {
            const int qs = blk_at(sp);
            const int Bq = Bnext;
            if (sp + 1 < pos && rok) Bnext = Bg[(size_t)blk_at(sp + 1) * N + v];
            unsigned long long word;
            { int spins = 0; do { word = lds_volatile_b64(&sWord[sp]); } while (word == 0ull && ++spins < kSpinMax);
              if (word == 0ull) { atomicOr(a.err, kErrSyncTimeout); word = 2ull << 32; } }
            const int smode = (int)(word >> 32), Eq = (int)(unsigned)word;
            const double* om = sOm + 32 * qs;
            if (smode == 1) {
                const int r = 32 * qs + lane;
                const bool expect = r >= (FWD ? 0 : 1) && r < N;
                int spins = 0;
                while (!__all_sync(kFullMask, !expect || lds_volatile_b64(&om[lane]) != 0ull) && ++spins < kSpinMax) {}
            } else {
                asm volatile("fence.acq_rel.cluster;" ::: "memory");
            }
            const int i0 = n_consumed;
            mbar_wait(bar_s + (i0 % SLOTS) * 8, (unsigned)(i0 / SLOTS) & 1u);
            mbar_wait(bar_s + ((i0 + 1) % SLOTS) * 8, (unsigned)((i0 + 1) / SLOTS) & 1u);
            if (smode == 1) {
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double* kp = ring_w + ((i0 + h) % SLOTS) * HALF + lane;
                    const double* op = om + 16 * h;
#pragma unroll
                    for (int c = 0; c < 16; c += 4) {
                        const double2 w01 = *reinterpret_cast<const double2*>(op + c);
                        const double2 w23 = *reinterpret_cast<const double2*>(op + c + 2);
                        acc0 = fma(kp[(c + 0) * 32], w01.x, acc0);
                        acc1 = fma(kp[(c + 1) * 32], w01.y, acc1);
                        acc2 = fma(kp[(c + 2) * 32], w23.x, acc2);
                        acc3 = fma(kp[(c + 3) * 32], w23.y, acc3);
                    }
                }
                if (rok) ext_fold(am, ae, (acc0 + acc1) + (acc2 + acc3), Bq + Eq);
            } else {
                const int last_need_c = rok ? (FWD ? v : N - 1 - v) : -1;
                const int s0 = g_lo(qs), s1 = g_hi(qs);
#pragma unroll 1
                for (int s_ = s0; s_ <= s1; ++s_) {
                    if (s_ <= last_need_c) {
                        const int r = row_of(s_);
                        const int4 c = factor_exact<FWD>(a, r, v);
                        ext_fma(am, ae, ext_m(c), c.z, sOm[r], sEx[r]);
                    }
                }
            }
            n_consumed += 2;
            __syncwarp();
            if (lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_one();
            issue_one();
        }
        int mode = 0, Eq = 0;
        const int own_lo = g_lo(q), own_hi = g_hi(q), n_own = own_hi - own_lo + 1;
        {
            // ---- owner phase of block q (= my block jown)
            double amo = rok ? am : 0.0;
            int aeo = rok ? ae : kExtZeroExp;
            const int last_need = rok ? (FWD ? v : N - 1 - v) : -1;
            auto dest = [&](int k) {             // destination blocks, the next owner's first
                const int nxt = ((FWD ? gwarp + 1 : gwarp + NW - 1) % NW) / WPC;
                return (unsigned)((nxt + k) % kClusterSize);
            };
            auto emit = [&](int s_, double m, int e) {
                const Ext wn = ext_normalize(m, e);
                const int i = FWD ? s_ : N - s_;
                (FWD ? a.Wm : a.Wbm)[i] = wn.m;
                (FWD ? a.We : a.Wbe)[i] = wn.e;
                const double val = -(log(wn.m) + (double)wn.e * 0.6931471805599453) / a.beta;
                if (!isfinite(val)) atomicOr(a.err, FWD ? kErrOverflowFwd : kErrOverflowBwd);
                (FWD ? a.V : a.Vb)[i] = (s_ == 0) ? 0.0 : val;
            };
            unsigned long long hbits;
            { int spins = 0; do { hbits = lds_volatile_b64(&sHandOm[pos]); } while (hbits == 0ull && ++spins < kSpinMax);
              if (hbits == 0ull) atomicOr(a.err, kErrSyncTimeout); }
            const double hm = __longlong_as_double((long long)hbits);
            int he = 0;
            if (pos > 0) {
                const unsigned long long wprev = lds_volatile_b64(&sWord[pos - 1]);   // (seen non-zero as a consumer)
                he = ((int)(wprev >> 32) == 1) ? (int)(unsigned)wprev : sHandE[pos];
            }
            const Ext n0 = ext_normalize(hm, he);
            const int E = n0.e;
            const double om0 = n0.m;
            const int d = aeo - E - Bown;
            const double A = (amo == 0.0) ? 0.0 : ext_to_double(amo, min(d, 600));
            bool exact = __any_sync(kFullMask, (amo != 0.0) && (d > 600)) || !(om0 > 0.0) || !Gok;
            const int tt = FWD ? 0 : row_of(own_lo) & 31;
            const int kslot = FWD ? lane : (tt - lane) & 31;
            mode = 2;
            if (!exact) {
                double* rho_w = sRho + lw * 32;
                rho_w[kslot] = A;
                __syncwarp();
                double u0 = om0 * hrow, u1 = 0.0, u2 = 0.0, u3 = 0.0;
#pragma unroll
                for (int k = 0; k < 32; k += 4) {
                    const double2 r01 = *reinterpret_cast<const double2*>(rho_w + k);
                    const double2 r23 = *reinterpret_cast<const double2*>(rho_w + k + 2);
                    u0 = fma(Grow[k], r01.x, u0);
                    u1 = fma(Grow[k + 1], r01.y, u1);
                    u2 = fma(Grow[k + 2], r23.x, u2);
                    u3 = fma(Grow[k + 3], r23.y, u3);
                }
                const double u = (u0 + u1) + (u2 + u3);
                const bool have = kslot < n_own;
                const unsigned hi = (unsigned)__double2hiint(u);
                const bool okall = __all_sync(kFullMask, !have || (hi - (323u << 20) < (1000u << 20)));
                if (okall) {
                    const void* pm = !have ? nullptr : (kslot == n_own - 1) ? (const void*)&sHandOm[pos + 1]
                                                                           : (const void*)&sOm[row_of(own_lo + kslot + 1)];
                    const unsigned long long ub = (unsigned long long)__double_as_longlong(u);
                    const unsigned long long ob = (unsigned long long)__double_as_longlong(om0);
                    const unsigned long long wb = (1ull << 32) | (unsigned long long)(unsigned)E;
                    for (int k = 0; k < kClusterSize; ++k) {
                        const unsigned dd = dest(k);
                        if (have) st_cluster_b64(map_to_cta(pm, dd), ub);
                        if (lane == 0) {
                            st_cluster_b64(map_to_cta(&sOm[row_of(own_lo)], dd), ob);
                            st_cluster_b64(map_to_cta(&sWord[pos], dd), wb);
                        }
                    }
                    mode = 1; Eq = E;
                    if (have) emit(own_lo + kslot + 1, u, E);
                    if (lane == 0 && pos == 0) emit(0, om0, E);
                } else {
                    exact = true;
                }
            }
            if (exact) {
                double wm = n0.m;
                int we = n0.e;
                auto push = [&](int s_, double m, int e) {
                    const bool last = s_ == own_hi + 1;
                    const void* pm = last ? (const void*)&sHandOm[pos + 1] : (const void*)&sOm[row_of(s_)];
                    const void* pe = last ? (const void*)&sHandE[pos + 1] : (const void*)&sEx[row_of(s_)];
                    for (int k = 0; k < kClusterSize; ++k) {
                        st_cluster_s32(map_to_cta(pe, dest(k)), e);
                        st_cluster_b64(map_to_cta(pm, dest(k)), (unsigned long long)__double_as_longlong(m));
                    }
                };
                if (lane == 0) { push(own_lo, wm, we); if (pos == 0) emit(0, wm, we); }
#pragma unroll 1
                for (int st = own_lo; st <= own_hi; ++st) {
                    if (st <= last_need) {
                        const int4 c = factor_exact<FWD>(a, row_of(st), v);
                        ext_fma(amo, aeo, ext_m(c), c.z, wm, we);
                    }
                    const int lane_o = row_of(st) & 31;
                    const Ext fin = ext_normalize(FWD ? amo * a.Inv[st + 1] : amo, aeo);
                    wm = __shfl_sync(kFullMask, fin.m, lane_o);
                    we = __shfl_sync(kFullMask, fin.e, lane_o);
                    if (lane == lane_o) { push(st + 1, fin.m, fin.e); emit(st + 1, fin.m, fin.e); }
                }
                __syncwarp();
                if (lane == 0) {
                    asm volatile("fence.acq_rel.cluster;" ::: "memory");
                    const unsigned long long wb = 2ull << 32;
                    for (int k = 0; k < kClusterSize; ++k) st_cluster_b64(map_to_cta(&sWord[pos], dest(k)), wb);
                }
            }
            if (lane == 0) { (FWD ? a.statf : a.statb)[pos] = mode; if (mode == 2) atomicOr(&a.sync[0], 1); }
        }
        (void)Eq;
    }
    cluster_sync_all();
    if (crank == 0 && tid == 0) a.sync[FWD ? 2 : 3] += 1;
}

__global__ void __launch_bounds__(256, 1) k_exch_recur_cluster(ExArgs a) {
    extern __shared__ __align__(16) double smem_d[];
    tl_begin(a.tl1);
    grid_launch_dependents();   // whatever follows on this stream with the programmatic attribute (the pair tiles of a captured
                                // step) may be scheduled now: these few blocks hold their SMs before that grid floods the GPU
    if (blockIdx.x < kClusterSize) recur_cluster<true>(a, smem_d);
    else recur_cluster<false>(a, smem_d);
    tl_end(a.tl1);
}

__global__ void __launch_bounds__(256, 1) k_exch_recur_cluster_multi(ExArgs a) {
    extern __shared__ __align__(16) double smem_d[];
    tl_begin(a.tl1);
    grid_launch_dependents();
    if (blockIdx.x < kClusterSize) recur_cluster_multi<true>(a, smem_d);
    else recur_cluster_multi<false>(a, smem_d);
    tl_end(a.tl1);
}

// ---------------------------------------------------------------- 4. exterior spring forces (K7 + K8)
// The kernel runs beside the pair tiles, and every block it needs has to wait for a pair-tile block to retire
// (in-kernel timeline: 1024 blocks of a block-per-particle version took 22 us to trickle in). So: 128 blocks of 4
// warps, one WARP per (exterior bead, particle l) and two such tasks per warp at N = 512. What a term needs besides
// its Boltzmann factor -- the other recurrence's weights and one bead slice -- is staged in shared memory once per
// block (even blocks serve the first bead, odd blocks the last), and a task issues all of its factor loads (16 B each,
// L2) before it uses the first one, so a task costs one L2 round trip instead of one per term.
#ifndef PIMDB_EXCH_FW
#define PIMDB_EXCH_FW 4
#endif
constexpr int kFW = PIMDB_EXCH_FW;           // warps per block
constexpr int kFU = 16;                      // terms per lane held in flight (covers N <= 512)
// STAGE: 1 = weights, exponents and the bead slice in shared memory (N <= ~5000), 2 = weights and exponents only (the
// slice is read from global memory; N <= 8192), 0 = nothing staged and no chunk skipping (larger N)
// (At most 160 registers: 160 x 128 = 20480 was exactly what ONE retiring pair-tile block of round 1 (80 registers x 256
// threads) freed; at 165 the kernel had to wait for two neighbouring slots. The round-2 pair-tile blocks are smaller (94
// registers x 128 threads), and smaller blocks of this kernel were measured -- 2 and 1 warps, 10.3 and 13.1 us against 7.4 --
// so the shape stays: what delays its start now is the pair-tile grid still dispatching, not the size of a freed slot.)
// KSRC: the factors come from the block-scaled tiles (N > 512) instead of the 16-byte tables
template <int D, int STAGE, bool KSRC, bool EXACT>
__device__ __forceinline__ void exch_forces_body(const ExArgs& a, double* fsm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = a.N;
    const int which = blockIdx.x & 1;         // 0: first bead, 1: last bead
    const int nblk = gridDim.x >> 1, blk = blockIdx.x >> 1;
    // per-u operands: weight mantissa gm[u], exponent ge[u], slice xo[c][u]
    //   first bead (sum over u >= l-1): g = Wb[u+1] / (u+1), xo = bead P;   last bead (sum over u <= l+1): g = W[u], xo = bead 1
    double* gm = fsm;
    int* ge = reinterpret_cast<int*>(fsm + N + 1);
    double* xo = fsm + N + 1 + ((N + 2) >> 1);
    const double* xo_g = which == 0 ? a.xP : a.x1;
    if (STAGE) {
        for (int u = threadIdx.x; u < N; u += blockDim.x) {
            if (which == 0) { gm[u] = a.Wbm[u + 1] * a.Inv[u + 1]; ge[u] = a.Wbe[u + 1]; }
            else { gm[u] = a.Wm[u]; ge[u] = a.We[u]; }
            if (STAGE == 1) {
#pragma unroll
                for (int c = 0; c < D; ++c) xo[c * N + u] = xo_g[(size_t)c * N + u];
            }
        }
        __syncthreads();
    }
    auto g_of = [&](int u, int& e) {
        if (STAGE) { e = ge[u]; return gm[u]; }
        if (which == 0) { e = a.Wbe[u + 1]; return a.Wbm[u + 1] * a.Inv[u + 1]; }
        e = a.We[u];
        return a.Wm[u];
    };
    auto xo_of = [&](int c, int u) { return STAGE == 1 ? xo[c * N + u] : xo_g[(size_t)c * N + u]; };
    const bool active = which == 0 ? a.do_first : a.do_last;
    const double iWN = 1.0 / a.Wm[N];
    const int eWN = a.We[N];
    const int4* Ctab = which == 0 ? a.Cf : a.Cb;             // (STAGE == 0 only: N beyond the block-scaled tiles)
    const double* Ktab = which == 0 ? a.Kf : a.Kb;           // block-scaled factor tiles, 32 x 32, [row block][column block]
    const int nbk = (N + 31) >> 5;
    constexpr bool exact = EXACT;   // recompute the factors instead of reading tiles (see any_exact_block)
    for (int l = blk * kFW + warp; active && l < N; l += nblk * kFW) {
        // first bead: f_l = k [ sum_{u=max(0,l-1)}^{N-1} P(u->l) mi(r^P_u - r^1_l) + mi(r^2_l - r^1_l) ]
        //             P(u->l) = W[l] c(l,u) Wb[u+1] / ((u+1) W[N]),  P(l-1->l) = 1 - W[l] Wb[l] / W[N]
        // last bead:  f_l = k [ sum_{u=0}^{min(l+1,N-1)} P(l->u) mi(r^1_u - r^P_l) + mi(r^{P-1}_l - r^P_l) ]
        //             P(l->u) = W[u] [c(u,l)/(l+1)] Wb[l+1] / W[N],  P(l->l+1) = 1 - W[l+1] Wb[l+1] / W[N]
        const double wl = (which == 0 ? a.Wm[l] : a.Wbm[l + 1]) * iWN;
        const int el = (which == 0 ? a.We[l] : a.Wbe[l + 1]) - eWN;
        const int ulo = which == 0 ? max(0, l - 1) : 0, uhi = which == 0 ? N - 1 : min(l + 1, N - 1);
        const int special = which == 0 ? l - 1 : l + 1;               // the neighbour term (1 - ...), no factor
        const double* xs = which == 0 ? a.x1 : a.xP;
        double xl[D], acc[D];
#pragma unroll
        for (int c = 0; c < D; ++c) { xl[c] = xs[(size_t)c * N + l]; acc[c] = 0.0; }
        const int4* Crow = KSRC ? nullptr : Ctab + (size_t)l * N;
        // row l of the factor matrix inside its row block of tiles: element u sits at Krow[(u >> 5) * 1024 + (u & 31)], and
        // equals the factor divided by 2^B[l / 32][u] (exactly: a power-of-two scaling)
        const double* Krow = KSRC ? Ktab + (size_t)(l >> 5) * nbk * 1024 + (l & 31) * 32 : nullptr;
        // Upper bound of a term's binary exponent from the block-scaled tables: B[l/32][u] >= exponent of the factor
        // (l, u), so  e(term) <= el + B + e(g_u) + 3.  Below -1080 the connection probability is exactly 0 (the same
        // cut ext_to_double applies); a 32-wide chunk whose lanes are all below it is skipped before its factors are
        // loaded -- in a cold liquid that is everything but the chunks next to l.
        const int* Brow = STAGE ? (which == 0 ? a.Bf : a.Bb) + (size_t)(l >> 5) * N : nullptr;
        for (int ub = ulo; ub <= uhi; ub += 32 * kFU) {     // warp-uniform trip count (votes inside)
            const int u0 = ub + lane;
            int4 cv[KSRC ? 1 : kFU];
            double kv[KSRC ? kFU : 1];
            int bexp[STAGE ? kFU : 1];
            unsigned live = ~0u;                  // bit k: chunk k of this batch may hold a non-zero probability
            if (STAGE) {
#pragma unroll
                for (int k = 0; k < kFU; ++k) {
                    const int u = u0 + 32 * k;
                    bexp[k] = (u <= uhi) ? __ldg(Brow + u) : kExtZeroExp;
                }
                live = 0u;
#pragma unroll
                for (int k = 0; k < kFU; ++k) {
                    const int u = u0 + 32 * k;
                    const bool maybe = (u <= uhi) && (u == special || el + bexp[k] + ge[u] >= -1080);
                    if (__any_sync(kFullMask, maybe)) live |= 1u << k;
                }
            }
            // (all loads of the batch are issued before the first is used; the rarely taken recompute path sits outside that loop)
            if (!KSRC || !exact) {
#pragma unroll
                for (int k = 0; k < kFU; ++k) {
                    const int u = u0 + 32 * k;
                    const bool want = ((live >> k) & 1u) && (u <= uhi && u != special);
                    if (KSRC) kv[k] = want ? __ldg(Krow + (size_t)(u >> 5) * 1024 + (u & 31)) : 0.0;
                    else cv[k] = want ? __ldg(Crow + u) : make_int4(0, 0, 0, kExtZeroExp);
                }
            } else {
#pragma unroll
                for (int k = 0; k < kFU; ++k) {
                    const int u = u0 + 32 * k;
                    const bool want = ((live >> k) & 1u) && (u <= uhi && u != special);
                    kv[k] = 0.0;
                    if (want) {
                        const int4 c = which == 0 ? factor_exact<true>(a, l, u) : factor_exact<false>(a, l, u);
                        kv[k] = ext_m(c); bexp[k] = c.z;         // the factor itself: mantissa and its own exponent
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kFU; ++k) {
                const int u = u0 + 32 * k;
                if (((live >> k) & 1u) && u <= uhi) {
                    double pr;
                    if (u == special) {
                        pr = which == 0 ? 1.0 - ext_to_double(wl * a.Wbm[l], el + a.Wbe[l])
                                        : 1.0 - ext_to_double(wl * a.Wm[l + 1], el + a.We[l + 1]);
                    } else {
                        int e;
                        const double g = g_of(u, e);
                        if (KSRC) pr = ext_to_double(wl * kv[k] * g, el + bexp[k] + e);
                        else pr = ext_to_double(wl * ext_m(cv[k]) * g, el + cv[k].z + e);
                    }
                    if (pr != 0.0) {   // most connection probabilities underflow to exactly 0: skip their separations
#pragma unroll
                        for (int c = 0; c < D; ++c) {
                            double dx = xo_of(c, u) - xl[c];
                            if (a.pbc) dx = min_image(dx, a.L, a.invL);
                            acc[c] = fma(pr, dx, acc[c]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = warp_sum(acc[c]);
        if (lane == 0) {
            const double* xn = which == 0 ? a.x2 : a.xPm1;            // the interior neighbour bead
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dx = xn[(size_t)c * N + l] - xl[c];
                if (a.pbc) dx = min_image(dx, a.L, a.invL);
                a.F[(size_t)(which * D + c) * N + l] = (acc[c] + dx) * a.k;
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.Vb[0] = a.V[N];   // V_backwards[0] = V[N] (quadratic_bosonic_exchange.cpp:127)
}

// Two kernels, launched one behind the other: the common one reads the factor tiles and returns at once when a block of the
// current tables was solved exactly; the other recomputes every factor and returns at once when none was. (One kernel with
// both paths -- inlined or behind a call -- slowed the common path from 11 to 14-18 us at C3.)
template <int D, int STAGE, bool KSRC, bool EXACT>
__global__ void __maxnreg__(160) k_exch_forces(ExArgs a) {   // (blocks of 32 * kFW threads; __maxnreg__ excludes __launch_bounds__)
    extern __shared__ __align__(16) double fsm[];
    if (KSRC) {
        if (EXACT) grid_dependency_wait();                   // (launched early behind the common kernel)
        else grid_launch_dependents();
        if (any_exact_block(a) != EXACT) return;
    }
    tl_begin(a.tl2);
    exch_forces_body<D, STAGE, KSRC, EXACT>(a, fsm);
    tl_end(a.tl2);
}

// ---------------------------------------------------------------- estimators (bead-0 owner only), one block
// e[m] = sum_{j<m} w(m,j) (e[j] - E^{[j..m-1]}),  w(m,j) = c(j,m-1) W[j] / (m W[m])   (primEstimator :250-279)
// Column-wise; the weights are extended-range products (no exp), the chain per step is one FMA + one barrier.
template <int D, int R>
__global__ void __launch_bounds__(1024) k_exch_estimators(ExArgs a) {
    extern __shared__ double se[];   // e[0..N]
    __shared__ double red[32];
    const int tid = threadIdx.x, nt = blockDim.x, N = a.N;
    const bool exact = (a.Kf && !a.Cf) ? any_exact_block(a) : false;
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
                double cm;
                int ce;
                if (exact) {
                    const int4 c = factor_exact<true>(a, j, v);
                    cm = ext_m(c); ce = c.z;
                } else if (a.Kf && !a.Cf) {   // block-scaled tile entry (j, v) and its block exponent
                    cm = __ldg(&a.Kf[((size_t)(j >> 5) * ((N + 31) >> 5) + (v >> 5)) * 1024 + (j & 31) * 32 + (v & 31)]);
                    ce = __ldg(&a.Bf[(size_t)(j >> 5) * N + v]);
                } else {
                    const int4 c = __ldg(&a.Cf[(size_t)j * N + v]);
                    cm = ext_m(c); ce = c.z;
                }
                const double wgt = ext_to_double(cm * wjm * iwm[r], ce + wje + iwe[r]);
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
    const bool exact = (a.Kb && !a.Cb) ? any_exact_block(a) : false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i / N), u = (int)(i % N);
        double pr = 0.0;
        if (u == l + 1) pr = 1.0 - ext_to_double(a.Wm[l + 1] * a.Wbm[l + 1] * iWN, a.We[l + 1] + a.Wbe[l + 1] - eWN);
        else if (u <= l) {
            double cm;
            int ce;
            if (exact) {
                const int4 c = factor_exact<false>(a, l, u);
                cm = ext_m(c); ce = c.z;
            } else if (a.Kb && !a.Cb) {
                cm = a.Kb[((size_t)(l >> 5) * ((N + 31) >> 5) + (u >> 5)) * 1024 + (l & 31) * 32 + (u & 31)];
                ce = a.Bb[(size_t)(l >> 5) * N + u];
            } else {
                cm = ext_m(a.Cb[i]); ce = a.Cb[i].z;
            }
            pr = ext_to_double(a.Wm[u] * cm * a.Wbm[l + 1] * iWN, a.We[u] + ce + a.Wbe[l + 1] - eWN);
        }
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
    a.Cf = s->exC; a.Cb = s->exC ? s->exC + NN : nullptr;
    const size_t nbk = (size_t)((s->N + 31) / 32);
    // (PIMDB_EXCH_NOBLOCKED=1, read when the handle is created: the scalar recurrences on the 16-byte tables as a cross-check)
    a.Kf = s->exK; a.Kb = s->exK ? s->exK + nbk * nbk * 1024 : nullptr;
    a.Bf = s->exB; a.Bb = s->exB ? s->exB + (size_t)((s->N + 31) / 32) * s->N : nullptr;
    a.Gf = s->exG; a.Gb = s->exG ? s->exG + nbk * 1024 : nullptr;
    a.Hf = s->exG ? s->exG + 2 * nbk * 1024 : nullptr; a.Hb = s->exG ? s->exG + 2 * nbk * 1024 + nbk * 32 : nullptr;
    a.Gokf = s->exGok; a.Gokb = s->exGok ? s->exGok + (s->N + 31) / 32 : nullptr;
    a.sync = s->exSync;
    a.tl0 = a.tl1 = a.tl2 = nullptr;
    a.statf = s->exGok ? s->exGok + 2 * ((s->N + 31) / 32) : nullptr;
    a.statb = s->exGok ? s->exGok + 3 * ((s->N + 31) / 32) : nullptr;
    a.Wm = s->exWm; a.Wbm = s->exWm + (s->N + 1);
    a.We = s->exWe; a.Wbe = s->exWe + (s->N + 1);
    a.V = s->exV; a.Vb = s->exVb; a.F = s->exF;
    a.obs = s->obs_d; a.err = s->err_d; a.dbg = s->dbg_buf;
    a.N = s->N; a.D = s->D; a.pbc = s->cfg.pbc;
    a.do_first = s->has_first; a.do_last = s->has_last;
    a.k = s->kspring; a.beta = s->exch_beta; a.h = 0.5 * s->exch_beta * s->kspring;
    a.L = s->L; a.invL = 1.0 / s->L;
    a.halo_flag = s->peer_on ? s->peer.mine->halo_flag : nullptr;
    a.halo_seq = s->peer_on ? s->peer.seq + 1 : nullptr;
    a.timeout_ns = s->peer.timeout_ns;
    a.tiles_diag_only = 0;
    return a;
}

// One launch helper for the kernels of the exchange chain. Every launch carries an explicit PRIORITY attribute (the side
// stream's priority is not what decides which pending thread block gets a freed SM slot once the launch is a node of a
// captured graph: measured, the exterior-force kernel sat behind the whole queue of pair-tile blocks), optionally a cluster
// dimension and programmatic stream serialisation.
template <typename K>
static void launch_chain(Sim* s, K kernel, const ExArgs& a, int grid, int block, size_t smem, cudaStream_t st, int cluster, bool pdl) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(block); lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[3];
    int n = 0;
    at[n].id = cudaLaunchAttributePriority; at[n].val.priority = s->prio_hi; ++n;
    if (cluster > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1; ++n;
    }
    if (pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1; ++n;
    }
    lc.attrs = at;
    lc.numAttrs = n;
    cudaLaunchKernelEx(&lc, kernel, a);
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
    launch_chain(s, k_exch_recur<R, ST>, a, 2, nt, smem, st, 1, false);
    return PIMDB_OK;
}

static int run_recursion(Sim* s, const ExArgs& a, cudaStream_t st) {
    int nt;
    const int R = rows_per_thread(s->N, nt);
    const int nblk = (s->N + 31) / 32;                       // 32-row blocks
    const bool blocked_ok = a.Kf != nullptr;   // (the scalar cross-check path is chosen when the handle is created: no tiles then)
    const bool pdl = s->pdl_recur;     // launched right behind k_exch_coeff_tiles on the same stream (api.cu enqueue_forces)
    if (blocked_ok && nblk > 8 * kClusterSize) {
        // more than 64 row blocks (2048 < N <= 8192): the same cluster, up to 4 row blocks per warp
        const int nb = nblk, nb2 = (nb + 3) & ~1, wpc = kMultiWpc;
        const size_t smem_cl = sizeof(double) * ((size_t)wpc * 3 * 512 + 32 * nb + 32 * wpc + nb2) + 8 * ((size_t)nb2 + wpc * 3 + (wpc & 1))
                               + sizeof(int) * ((size_t)32 * nb + nb2) + 16;
        cudaFuncSetAttribute(k_exch_recur_cluster_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cl);
        launch_chain(s, k_exch_recur_cluster_multi, a, 2 * kClusterSize, 32 * wpc, smem_cl, st, kClusterSize, pdl);
        return PIMDB_OK;
    }
    if (blocked_ok) {
        // blocked recurrence on two clusters of 8 thread blocks (forward, backward), 1..8 warps each (N <= 2048)
        const int nb = nblk, nb2 = (nb + 3) & ~1, wpc = (nb + kClusterSize - 1) / kClusterSize;
        const size_t smem_cl = sizeof(double) * ((size_t)wpc * 3 * 512 + 32 * nb + 32 * wpc + nb2) + 8 * ((size_t)nb2 + wpc * 3 + (wpc & 1))
                               + sizeof(int) * ((size_t)32 * nb + nb2) + 16;
        // On a bead shard the recurrence blocks ask for (nearly) all the shared memory of their SMs, so that no pair-tile block
        // fits beside them: the recurrence is a latency chain, and sharing its SM's issue slots with 20 pair-tile warps stretches
        // it from 17 to 26 us at N = 512. A shard has fewer pair tiles than the exchange chain is long, so the 16 SMs are not
        // missed (measured on 2 GPUs at C3: 70.8 -> 68.2 us per step); a handle that owns every bead is balanced between the two
        // arms and loses what the recurrence gains (62.1 -> 63.9 us), so it does not. PIMDB_RECUR_SMEM_KB overrides (0 = off).
        size_t smem_req = smem_cl;
        size_t pad_kb = (s->peer_on && !s->peer_shares_gpu) ? 218 : 0;    // (shards sharing one GPU would starve each other's clusters)
        if (const char* e = getenv("PIMDB_RECUR_SMEM_KB")) pad_kb = (size_t)std::max(0, atoi(e));
        smem_req = std::max(smem_cl, std::min<size_t>(pad_kb, 227) * 1024);
        if (smem_req > 48 * 1024)
            cudaFuncSetAttribute(k_exch_recur_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_req);
        launch_chain(s, k_exch_recur_cluster, a, 2 * kClusterSize, 32 * wpc, smem_req, st, kClusterSize, pdl);
        return PIMDB_OK;
    }
    // scalar recurrence, one unknown per step: the cross-check path (PIMDB_EXCH_NOBLOCKED=1) and N > 8192
    switch (R) {
        case 1: return launch_recur<1, 8>(s, a, st, nt);    // N <= 1024
        case 2: return launch_recur<2, 4>(s, a, st, nt);    // N <= 2048
        case 4: return launch_recur<4, 0>(s, a, st, nt);    // N <= 4096: direct global loads
        case 8: return launch_recur<8, 0>(s, a, st, nt);    // N <= 8192
        case 16: return launch_recur<16, 0>(s, a, st, nt);  // N <= 16384
        default:
            s->err = "natoms too large for the single-block exchange recursion (max 16384)";
            return PIMDB_ERR_INVALID_ARGUMENT;
    }
}

// part 0: prefix sums + Boltzmann factors / factor tiles + block inverses (fully parallel);
// part 1: the two recurrences (latency-bound) + the exterior forces; part 2: the recurrences only; part 3: the forces only.
// They are separate so that the caller can place them on different streams around the pair tiles (api.cu enqueue_forces).
template <int D>
static int exchange_impl(Sim* s, cudaStream_t st, int part) {
    ExArgs a = make_args(s);
    if (part == 0) a.tl0 = tl_slot(s, 2);
    if (part == 1 || part == 2) a.tl1 = tl_slot(s, 3);
    if (part == 1 || part == 3) a.tl2 = tl_slot(s, 4);
    if (part == 0) {
        if (a.Kf) {      // N <= 8192: factor tiles + diagonal-block inverses; up to N = 512 the tiles recompute the prefix sums
            const int nb = (s->N + 31) / 32;
            if (s->N > kExchFastMaxN) {
                launch_chain(s, k_exch_prefix<D>, a, 1, 1024, 0, st, 1, false);
                s->launches += 1;
            }
            if (s->N > kExchFastMaxN && !getenv("PIMDB_EXCH_TILES1024")) {
                // many waves of tiles: the off-diagonal ones four rows per thread, then one block per diagonal tile (+ inverse)
                launch_chain(s, k_exch_offdiag_tiles<D>, a, nb * nb, 256, 0, st, 1, false);
                s->launches += 1;
                a.tiles_diag_only = 1;
                s->pdl_next = false;
                launch_chain(s, k_exch_coeff_tiles<D>, a, nb, 1024, 0, st, 1, false);
                a.tiles_diag_only = 0;
            } else {
                const bool early = s->pdl_next && s->N <= kExchFastMaxN;   // (beyond that the prefix kernel sits in front)
                s->pdl_next = false;
                launch_chain(s, k_exch_coeff_tiles<D>, a, nb * nb, 1024, 0, st, 1, early);
            }
            s->launches += 1;
        } else {
            launch_chain(s, k_exch_prefix<D>, a, 1, 1024, 0, st, 1, false);
            launch_chain(s, k_exch_coeff<D>, a, grid_for((size_t)s->N * s->N, 256, 16 * kNumSM), 256, 0, st, 1, false);
            s->launches += 2;
        }
    } else {
        if (part != 3) {
            int rc = run_recursion(s, a, st);
            if (rc != PIMDB_OK) return rc;
            s->launches += 1;
        }
        if (part != 2) {
            // tasks (particles) per warp: a task is a chain of ~4 dependent L2 round trips, so one per warp while the grid
            // still fits the machine a few times over (PIMDB_EXCH_FTASKS overrides, for timing)
            static const int ftasks = getenv("PIMDB_EXCH_FTASKS") ? std::max(1, atoi(getenv("PIMDB_EXCH_FTASKS"))) : 1;
            const int per_kind = std::max(1, std::min((s->N + ftasks * kFW - 1) / (ftasks * kFW), 4 * kNumSM));
            const size_t smem_w = sizeof(double) * ((size_t)s->N + 1 + ((s->N + 2) >> 1));            // weights + exponents
            const size_t smem_full = smem_w + sizeof(double) * (size_t)D * s->N;                         // + the bead slice
            auto forces = [&](auto common, auto recompute, size_t smem, bool two) {
                if (smem > 48 * 1024) cudaFuncSetAttribute(common, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                launch_chain(s, common, a, 2 * per_kind, 32 * kFW, smem, st, 1, false);
                if (two) {   // tile-sourced factors: the recompute twin rides behind and returns at once unless a block went exact
                    if (smem > 48 * 1024) cudaFuncSetAttribute(recompute, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    launch_chain(s, recompute, a, 2 * per_kind, 32 * kFW, smem, st, 1, true);
                    s->launches += 1;
                }
            };
            if (a.Kf && a.Cf && smem_full <= 200 * 1024)          // N <= 512: staged operands, factors from the 16-byte tables
                forces(k_exch_forces<D, 1, false, false>, k_exch_forces<D, 1, false, false>, smem_full, false);
            else if (a.Kf && smem_full <= 200 * 1024)
                forces(k_exch_forces<D, 1, true, false>, k_exch_forces<D, 1, true, true>, smem_full, true);
            else if (a.Kf && smem_w <= 200 * 1024)
                forces(k_exch_forces<D, 2, true, false>, k_exch_forces<D, 2, true, true>, smem_w, true);
            else
                forces(k_exch_forces<D, 0, false, false>, k_exch_forces<D, 0, false, false>, 0, false);
            s->launches += 1;
        }
    }
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_exchange_part(Sim* s, cudaStream_t st, int part) {
    if (s->factorial) return part == 0 ? launch_factorial_exchange(s, st) : PIMDB_OK;   // (one short chain of four small kernels)
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
