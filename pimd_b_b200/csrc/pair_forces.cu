// Pair forces (K1) and the pair part of the energy estimators (K13), FP64, sm_100a.
//
// Replaces the i<j loop of Simulation::updatePhysicalForces (reference src/simulation.cpp:432-454) with
// getSeparation / applyMinimumImage (src/simulation.cpp:499-512, src/common.cpp:41-43) and
// AzizPotential / DipolePotential / HarmonicPotential ::gradV, ::V (src/potentials/*.cpp), and the second
// pair loop of EnergyObservable::calculatePotential (src/observables/energy.cpp:76-94).
//
// Decomposition: the N particles of a bead are cut into T = ceil(N/32) tiles. One WARP owns one unordered
// tile pair (I<=J) of one bead -- or, when the launch would otherwise fill the GPU fewer than ~6 times over, one
// half / quarter of its 32 rotations (`split`; C3 has 8704 tile pairs for 3552 resident warps: 2.45 waves, so a
// third of the machine idles through the tail unless the items are finer). Lane l holds particle 32I+l ("i") in
// registers; the j tile and its reaction-force accumulators live in shared memory, and the 32 rotations
// (l, (l+t)%32) enumerate the 32x32 tile: in a rotation every lane touches a different j, so the shared-memory
// reads and read-modify-writes are conflict-free and need no atomics. Every unordered pair is evaluated exactly once
// (Newton's third law). Diagonal tiles run the rotations t=1..16 only (t=16 on half the lanes). The warps that
// share a tile pair combine their partial sums through shared memory in a fixed order, then ONE of them writes the
// two 32-particle partial force vectors to a scratch slab S[bead][tile K][other tile M][axis][lane]; the assemble
// kernel sums the T partials of every particle in fixed order M=0..T-1 (deterministic, bit-reproducible) and adds
// the external and spring forces.
#include "internal.cuh"
#include "device_utils.cuh"

namespace pimdb {

#ifndef PIMDB_PAIR_MINBLOCKS
#define PIMDB_PAIR_MINBLOCKS 3
#endif
constexpr int kPairWarps = 8;   // warps per block

struct PairArgs {
    const double* x;      // first bead of this launch, slab stride S
    double* scratch;      // [nb][T][T][D][32]
    double* obs_part;     // [items * split][2] (V, virial) when OBS
    const ushort2* tile_ij;
    const ushort4* dual_ijj;   // work items of the two-rows-per-warp kernel: tiles (I, I+1) against tile J >= I + 2, all three full
    int N, T, TP, nb;
    int split;            // warps per tile pair: 1, 2 or 4
    size_t S;
    double L, invL, rc, par;
    int tie_hi;           // min_image_tie_threshold(L)
    ExpConsts ek;
    // Aziz constants in the combinations the loop uses (uniform registers, see ExpConsts):
    //   az_k2 = -alpha log2(e) / rm, az_drm = D rm, az_ca = -(eps/rm) A alpha,
    //   az_h{0,1,2} = (eps/rm) rm^{7,9,11} {6 C6, 8 C8, 10 C10}
    double az_k2, az_drm, az_ca, az_h0, az_h1, az_h2;
    double az_far2;       // (x_far rm)^2: beyond x_far the repulsive term A e^{-alpha x} is dropped (see pair_rotation)
    unsigned long long* tl;   // timeline slot (profiling aid) or nullptr
};

// dV/dr / r (so that grad V = g * r_vec) and optionally V, from r^2.
template <int POT, bool WANT_V>
__device__ __forceinline__ double pair_eval(double r2, const PairArgs& a, double& v) {
    const double par = a.par;
    if (POT == PIMDB_POT_HARMONIC) {            // reference src/potentials/harmonic.cpp:3-23 (par = m w^2)
        if (WANT_V) v = 0.5 * par * r2;
        return par;
    } else if (POT == PIMDB_POT_DIPOLE) {       // reference src/potentials/dipole.cpp:5-45 (par = strength)
        double ir = rsqrt_fast(r2);
        double ir2 = ir * ir;
        double ir3 = ir2 * ir;
        if (WANT_V) v = par * ir3;
        return -3.0 * par * ir3 * ir2;
    } else {                                    // Aziz HFDHE2, reference src/potentials/aziz.cpp:18-106
        // with x = r / rm: V = eps [A e^{-alpha x} - F(x) sum_k C_k x^-k], F = exp(-(D/x - 1)^2) for x < D, else 1
        const double ir = rsqrt_fast(r2);
        const double r = r2 * ir;
        const double e1 = exp2_lin_fast(r, a.az_k2, a.ek);            // e^{-alpha x}
        if (r >= a.az_drm) {                    // x >= D, the common case: no damping (include/potentials/aziz.h:25-34)
            // g = (eps/rm) [-A alpha e1 + x^-7 (6 C6 + 8 C8 x^-2 + 10 C10 x^-4)] / r, in powers of u = 1/r^2
            const double u = ir * ir, u2 = u * u;
            const double P = (u2 * u2) * fma(fma(a.az_h2, u, a.az_h1), u, a.az_h0);
            if (WANT_V) {
                const double ix2 = (kAzRm * kAzRm) * u;
                v = kAzEps * fma(kAzA, e1, -(ix2 * ix2 * ix2) * fma(fma(kAzC10, ix2, kAzC8), ix2, kAzC6));
            }
            return fma(a.az_ca, e1 * ir, P);
        }
        const double xs = r * (1.0 / kAzRm);
        double w, disp_f = 0.0;                 // w = dV/dx / eps
        if (xs > kEps && xs < 0.01) {           // hard-core branch: repulsion only (aziz.cpp:39-42, 78-79)
            w = -kAzA * kAzAlpha * e1;
        } else {
            const double ix = kAzRm * ir;
            const double ix2 = ix * ix, ix6 = ix2 * ix2 * ix2;
            const double q = kAzD * ix - 1.0;
            const double fdamp = exp_neg_fast(-q * q, a.ek);
            const double dfdamp = 2.0 * kAzD * ix2 * q * fdamp;
            const double disp = ix6 * fma(fma(kAzC10, ix2, kAzC8), ix2, kAzC6);
            const double ddisp = (ix6 * ix) * fma(fma(10.0 * kAzC10, ix2, 8.0 * kAzC8), ix2, 6.0 * kAzC6);
            w = fma(-kAzA * kAzAlpha, e1, ddisp * fdamp - disp * dfdamp);
            disp_f = disp * fdamp;
        }
        if (WANT_V) v = kAzEps * fma(kAzA, e1, -disp_f);
        return w * (ir * (kAzEps / kAzRm));
    }
}

// One rotation of the 32x32 tile: lane l meets the j particle at slot (l+t)%32 of the shared j tile.
// MASKED = false: every lane pair is a valid, distinct pair (off-diagonal tile, both tiles full).
template <int D, int POT, bool PBC, bool CUT, bool OBS, bool MASKED>
__device__ __forceinline__ void pair_rotation(const PairArgs& a, int t, int lane, bool diag, bool vi, int jbase,
                                              const double (&xi)[D], double (&fi)[D], double& vsum, double& virsum,
                                              const double* sx, double* sf) {
    const int src = (lane + t) & 31;
    double d[D], xo[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        xo[c] = sx[c * 32 + src];
        d[c] = xi[c] - xo[c];
    }
    // The reference always forms x_lower_index - x_higher_index before the minimum image
    // (src/simulation.cpp:433-437, 503); mi(+L/2) = mi(-L/2) = -L/2, so the order matters for particles
    // exactly half a box apart (perfect lattices). Only diagonal tiles can see the higher index first.
    if (PBC) min_image_vec<D>(d, a.L, a.invL, a.tie_hi, MASKED && diag && (src < lane));
    double r2 = d[0] * d[0];
#pragma unroll
    for (int c = 1; c < D; ++c) r2 = fma(d[c], d[c], r2);
    bool active = true;
    if (MASKED) active = vi && (jbase + src < a.N) && !(diag && t == 16 && lane >= 16);
    if (CUT) active = active && (sqrt(r2) < a.rc);       // strict '<' (src/simulation.cpp:444)
    if (MASKED || CUT) {
        if (!active) r2 = 1.0;                            // keep the arithmetic finite on masked lanes
    }
    double v = 0.0;
    double g;
    // Aziz, force only: when ALL 32 pairs of the rotation are beyond x_far = 3.6 the repulsive term is below 1e-11 of
    // the pair's own (dispersion) force -- and that force below 1e-3 of a near-neighbour force -- so the rotation
    // evaluates the dispersion term alone: one reciprocal instead of rsqrt + exp, 27 FP64 instructions instead of 49.
    // A warp-uniform decision (no divergence); how often it applies depends on how well particle order follows space
    // (64 % of the rotations for the lattice-ordered start of the He-4 workloads, PIMDB_PAIR_NOFAR=1 disables it).
    if (POT == PIMDB_POT_AZIZ && !OBS && __all_sync(kFullMask, r2 > a.az_far2)) {
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
        y = fma(y, fma(-r2, y, 1.0), y);
        const double u = fma(y, fma(-r2, y, 1.0), y);
        const double u2 = u * u;
        g = (u2 * u2) * fma(fma(a.az_h2, u, a.az_h1), u, a.az_h0);
    } else {
        g = pair_eval<POT, OBS>(r2, a, v);
    }
    if (MASKED || CUT) {
        if (!active) { g = 0.0; v = 0.0; }
    }
    // force on i: -g d;  reaction on j: +g d  (every lane owns a distinct j slot in this rotation)
#pragma unroll
    for (int c = 0; c < D; ++c) {
        fi[c] = fma(-g, d[c], fi[c]);
        sf[c * 32 + src] = fma(g, d[c], sf[c * 32 + src]);
    }
    if (OBS) {
        // energy.cpp:88-90: virial -= x_first . f_on_first, "first" = the lower particle index of the pair
        vsum += v;
        const bool i_first = !diag || (src > lane);
        double dot = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) dot += (i_first ? xi[c] : -xo[c]) * (g * d[c]);
        virsum += dot;
    }
    __syncwarp();
}

template <int D, int POT, bool PBC, bool CUT, bool OBS>
__global__ void __launch_bounds__(32 * kPairWarps, PIMDB_PAIR_MINBLOCKS) k_pair_tiles(PairArgs a) {
    __shared__ double s_x[kPairWarps][D * 32], s_f[kPairWarps][D * 32];
    grid_launch_dependents();   // a launch chained behind this one (the next slice of the pair tiles) may be scheduled as soon as
                                // every block of this grid has started
    tl_begin(a.tl);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sp = a.split;
    const long long gw = (long long)blockIdx.x * kPairWarps + warp;
    const long long item = gw / sp;
    const int part = (int)(gw - item * sp);
    const bool live = item < (long long)a.nb * a.TP;      // no early exit: every warp reaches the block barrier
    const int bl = live ? (int)(item / a.TP) : 0;
    const ushort2 ij = a.tile_ij[live ? item % a.TP : 0];
    const int I = ij.x, J = ij.y;
    const bool diag = (I == J);
    const double* xb = a.x + (size_t)bl * a.S;
    const int pi = I * kTile + lane, pj = J * kTile + lane;
    const bool vi = pi < a.N, vj = pj < a.N;
    double* sx = s_x[warp];
    double* sf = s_f[warp];

    double xi[D], fi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        xi[c] = vi ? xb[(size_t)c * a.N + pi] : 0.0;
        sx[c * 32 + lane] = vj ? xb[(size_t)c * a.N + pj] : 0.0;
        sf[c * 32 + lane] = 0.0;
        fi[c] = 0.0;
    }
    __syncwarp();
    double vsum = 0.0, virsum = 0.0;
    const int jbase = J * kTile;

    if (live) {
        // my share of the rotations: t = 0..31 (off-diagonal) or 1..16 (diagonal), cut into `sp` equal runs
        const int nrot = (diag ? 16 : 32) / sp;
        const int tb = (diag ? 1 : 0) + part * nrot;
        if (!diag && (J + 1) * kTile <= a.N) {   // I < J, so tile I is full as well
#pragma unroll 1
            for (int t = tb; t < tb + nrot; ++t)
                pair_rotation<D, POT, PBC, CUT, OBS, false>(a, t, lane, false, true, jbase, xi, fi, vsum, virsum, sx, sf);
        } else {
#pragma unroll 1
            for (int t = tb; t < tb + nrot; ++t)
                pair_rotation<D, POT, PBC, CUT, OBS, true>(a, t, lane, diag, vi, jbase, xi, fi, vsum, virsum, sx, sf);
        }
    }

    if (!OBS) {
        if (sp > 1) {   // parts 1.. hand their sums to part 0 (the j tile is no longer needed: it carries f_i)
            if (part != 0) {
#pragma unroll
                for (int c = 0; c < D; ++c) sx[c * 32 + lane] = fi[c];
            }
            __syncthreads();
        }
        if (live && part == 0) {
            double fj[D];
#pragma unroll
            for (int c = 0; c < D; ++c) fj[c] = sf[c * 32 + lane];
            for (int q = 1; q < sp; ++q) {
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    fi[c] += s_x[warp + q][c * 32 + lane];
                    fj[c] += s_f[warp + q][c * 32 + lane];
                }
            }
            double* s = a.scratch + (size_t)bl * a.T * a.T * D * kTile;
            if (diag) {
#pragma unroll
                for (int c = 0; c < D; ++c) s[(((size_t)I * a.T + I) * D + c) * kTile + lane] = fi[c] + fj[c];
            } else {
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    s[(((size_t)I * a.T + J) * D + c) * kTile + lane] = fi[c];
                    s[(((size_t)J * a.T + I) * D + c) * kTile + lane] = fj[c];
                }
            }
        }
    } else {
        vsum = warp_sum(vsum);
        virsum = warp_sum(virsum);
        if (lane == 0) {
            a.obs_part[2 * gw] = live ? vsum : 0.0;
            a.obs_part[2 * gw + 1] = live ? virsum : 0.0;
        }
    }
    tl_end(a.tl);
}

// ------------------------------------------------------------------------------------------------------
// Two i-particles per lane. The single-row kernel above issues one instruction per warp every ~8 cycles: two thirds of
// its stall cycles are fixed-latency dependencies of the FP64 chain of ONE pair (ncu: "wait" 3.5 of 8.1 cycles per
// instruction), and the registers do not allow more warps. Here a warp owns the tile pairs (I, J) and (I+1, J) at once:
// lane l keeps particle l of BOTH row tiles in registers and meets the same j particle in a rotation, so two independent
// pair evaluations interleave in every lane, the j coordinates are loaded once for two pairs, and the two reaction forces
// reach the j particle in one shared-memory read-modify-write. Only full, off-diagonal tiles (no masking); everything
// else -- diagonal tiles, the tile next to the diagonal, ragged last tiles, cutoffs -- goes through the kernel above.
// The reaction of both rows is accumulated together and written to the (J, I) slot of the scratch slab; the (J, I+1) slot
// gets zeros, so the fixed-order sum of the assembly is unchanged.
template <int D, int POT, bool PBC>
__global__ void __launch_bounds__(32 * kPairWarps, 2) k_pair_tiles_dual(PairArgs a, int ndual) {
    __shared__ double s_x[kPairWarps][D * 32], s_f[kPairWarps][D * 32];
    grid_launch_dependents();
    tl_begin(a.tl);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * kPairWarps + warp;
    if (gw >= (long long)a.nb * ndual) { tl_end(a.tl); return; }     // (no block-wide barrier in this kernel)
    const int bl = (int)(gw / ndual);
    const ushort4 it = a.dual_ijj[gw % ndual];
    const int I1 = it.x, I2 = it.y, J = it.z;
    const double* xb = a.x + (size_t)bl * a.S;
    double* sx = s_x[warp];
    double* sf = s_f[warp];
    double x1[D], x2[D], f1[D], f2[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        x1[c] = xb[(size_t)c * a.N + I1 * kTile + lane];
        x2[c] = xb[(size_t)c * a.N + I2 * kTile + lane];
        sx[c * 32 + lane] = xb[(size_t)c * a.N + J * kTile + lane];
        sf[c * 32 + lane] = 0.0;
        f1[c] = 0.0; f2[c] = 0.0;
    }
    __syncwarp();
#pragma unroll 1
    for (int t = 0; t < 32; ++t) {
        const int src = (lane + t) & 31;
        double d1[D], d2[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const double xo = sx[c * 32 + src];
            d1[c] = x1[c] - xo;
            d2[c] = x2[c] - xo;
        }
        if (PBC) {   // (I < J in both pairs: the separation is x_lower - x_higher as in the reference, no sign to fix at a tie)
            min_image_vec<D>(d1, a.L, a.invL, a.tie_hi, false);
            min_image_vec<D>(d2, a.L, a.invL, a.tie_hi, false);
        }
        double r1 = d1[0] * d1[0], r2 = d2[0] * d2[0];
#pragma unroll
        for (int c = 1; c < D; ++c) { r1 = fma(d1[c], d1[c], r1); r2 = fma(d2[c], d2[c], r2); }
        double g1, g2, v;
        if (POT == PIMDB_POT_AZIZ && __all_sync(kFullMask, r1 > a.az_far2 && r2 > a.az_far2)) {   // see pair_rotation
            double y1, y2;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y1) : "d"(r1));
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y2) : "d"(r2));
            y1 = fma(y1, fma(-r1, y1, 1.0), y1); y2 = fma(y2, fma(-r2, y2, 1.0), y2);
            const double u1 = fma(y1, fma(-r1, y1, 1.0), y1), u2 = fma(y2, fma(-r2, y2, 1.0), y2);
            const double q1 = u1 * u1, q2 = u2 * u2;
            g1 = (q1 * q1) * fma(fma(a.az_h2, u1, a.az_h1), u1, a.az_h0);
            g2 = (q2 * q2) * fma(fma(a.az_h2, u2, a.az_h1), u2, a.az_h0);
        } else {
            g1 = pair_eval<POT, false>(r1, a, v);
            g2 = pair_eval<POT, false>(r2, a, v);
        }
#pragma unroll
        for (int c = 0; c < D; ++c) {
            f1[c] = fma(-g1, d1[c], f1[c]);
            f2[c] = fma(-g2, d2[c], f2[c]);
            sf[c * 32 + src] = fma(g2, d2[c], fma(g1, d1[c], sf[c * 32 + src]));
        }
        __syncwarp();
    }
    double* sc = a.scratch + (size_t)bl * a.T * a.T * D * kTile;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        sc[(((size_t)I1 * a.T + J) * D + c) * kTile + lane] = f1[c];
        sc[(((size_t)I2 * a.T + J) * D + c) * kTile + lane] = f2[c];
        sc[(((size_t)J * a.T + I1) * D + c) * kTile + lane] = sf[c * 32 + lane];
        sc[(((size_t)J * a.T + I2) * D + c) * kTile + lane] = 0.0;
    }
    tl_end(a.tl);
}

// Deterministic final reduction of the per-warp (V, virial) partials: one block, fixed order.
__global__ void __launch_bounds__(1024) k_pair_obs_reduce(const double* part, long long n, double* out_v, double* out_vir) {
    __shared__ double sm[64];
    double acc[2] = {0.0, 0.0};
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        acc[0] += part[2 * i];
        acc[1] += part[2 * i + 1];
    }
    block_sum<2>(acc, sm);
    if (threadIdx.x == 0) {
        *out_v += acc[0];
        *out_vir += acc[1];
    }
}

// `early`: launch with programmatic stream serialisation. The kernel never waits for the grid in front of it (it reads
// nothing that grid writes), so it is scheduled as soon as every block of that grid has started -- used in a captured step
// to put the pair tiles right behind the resident blocks of the exchange recurrence (api.cu enqueue_forces).
template <typename K>
static void launch_tiles(K kernel, Sim* s, const PairArgs& a, int grid, bool early) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(32 * kPairWarps); lc.dynamicSmemBytes = 0; lc.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = early ? 1 : 0;
    cudaLaunchKernelEx(&lc, kernel, a);
}

template <int D, int POT, bool OBS>
static void dispatch2(Sim* s, const PairArgs& a, int grid, bool early) {
    const bool pbc = s->cfg.pbc != 0, cut = s->rc > 0.0;
    if (pbc && cut) launch_tiles(k_pair_tiles<D, POT, true, true, OBS>, s, a, grid, early);
    else if (pbc) launch_tiles(k_pair_tiles<D, POT, true, false, OBS>, s, a, grid, early);
    else if (cut) launch_tiles(k_pair_tiles<D, POT, false, true, OBS>, s, a, grid, early);
    else launch_tiles(k_pair_tiles<D, POT, false, false, OBS>, s, a, grid, early);
}

template <int D, bool OBS>
static void dispatch1(Sim* s, const PairArgs& a, int grid, bool early) {
    switch (s->cfg.int_potential) {
        case PIMDB_POT_AZIZ: dispatch2<D, PIMDB_POT_AZIZ, OBS>(s, a, grid, early); break;
        case PIMDB_POT_HARMONIC: dispatch2<D, PIMDB_POT_HARMONIC, OBS>(s, a, grid, early); break;
        case PIMDB_POT_DIPOLE: dispatch2<D, PIMDB_POT_DIPOLE, OBS>(s, a, grid, early); break;
        default: break;
    }
}

// Enqueue the tile kernel for owned beads [bead_lo, bead_lo+nb) (nb <= bead_chunk).
template <int D>
static void dispatch_dual_d(Sim* s, const PairArgs& a, int grid, int ndual, bool early) {
    auto go = [&](auto kernel) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(grid); lc.blockDim = dim3(32 * kPairWarps); lc.dynamicSmemBytes = 0; lc.stream = s->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at;
        lc.numAttrs = early ? 1 : 0;
        cudaLaunchKernelEx(&lc, kernel, a, ndual);
    };
    const bool pbc = s->cfg.pbc != 0;
    switch (s->cfg.int_potential) {
        case PIMDB_POT_AZIZ: pbc ? go(k_pair_tiles_dual<D, PIMDB_POT_AZIZ, true>) : go(k_pair_tiles_dual<D, PIMDB_POT_AZIZ, false>); break;
        case PIMDB_POT_HARMONIC: pbc ? go(k_pair_tiles_dual<D, PIMDB_POT_HARMONIC, true>) : go(k_pair_tiles_dual<D, PIMDB_POT_HARMONIC, false>); break;
        case PIMDB_POT_DIPOLE: pbc ? go(k_pair_tiles_dual<D, PIMDB_POT_DIPOLE, true>) : go(k_pair_tiles_dual<D, PIMDB_POT_DIPOLE, false>); break;
        default: break;
    }
}
static void dispatch_dual(Sim* s, const PairArgs& a, int grid, int ndual, bool early) {
    if (s->D == 1) dispatch_dual_d<1>(s, a, grid, ndual, early);
    else if (s->D == 2) dispatch_dual_d<2>(s, a, grid, ndual, early);
    else dispatch_dual_d<3>(s, a, grid, ndual, early);
}

// `scratch_lo`: the slot of the scratch slab that bead `bead_lo` writes to (several launches can fill one slab)
static int launch_chunk(Sim* s, int bead_lo, int nb, bool with_obs, bool early, int scratch_lo) {
    PairArgs a;
    a.x = s->x + (size_t)(bead_lo + 1) * s->S;
    a.scratch = s->pair_scratch + (size_t)scratch_lo * s->T * s->T * s->D * kTile;
    a.obs_part = s->pair_scratch;  // the scratch slab doubles as the (V, virial) partial buffer
    a.tile_ij = s->tile_ij;
    a.dual_ijj = s->dual_ijj;
    a.N = s->N; a.T = s->T; a.TP = s->TP; a.nb = nb; a.S = s->S;
    a.L = s->L; a.invL = 1.0 / s->L; a.rc = s->rc; a.par = s->pair_par;
    a.tie_hi = min_image_tie_threshold(s->L);
    a.ek = make_exp_consts();
    {
        const long double rm = kAzRm, g0 = (long double)kAzEps / rm, rm2 = rm * rm, rm7 = rm2 * rm2 * rm2 * rm;
        a.az_k2 = (double)(-(long double)kAzAlpha * 1.442695040888963407359924681001892137L / rm);
        a.az_drm = kAzD * kAzRm;
        a.az_far2 = getenv("PIMDB_PAIR_NOFAR") ? 1.0e300 : (3.6 * kAzRm) * (3.6 * kAzRm);
        a.az_ca = (double)(-g0 * kAzA * kAzAlpha);
        a.az_h0 = (double)(g0 * rm7 * 6.0L * kAzC6);
        a.az_h1 = (double)(g0 * rm7 * rm2 * 8.0L * kAzC8);
        a.az_h2 = (double)(g0 * rm7 * rm2 * rm2 * 10.0L * kAzC10);
    }
    // forces without a cutoff: the full off-diagonal tiles two rows at a time (k_pair_tiles_dual), the rest singly, as two
    // launches chained by a programmatic launch (no wait: they are independent) so that the short single items fill the tail
    // Measured on B200: the two-row kernel runs 8 independent pair chains per scheduler instead of 6 (98 registers, two
    // blocks per SM) and is ~1.3x faster per SM, but its blocks live twice as long: at C3 (1.5 waves of them) the tail eats
    // the gain (pair tiles 42 -> 46 us), at C4 (50 waves) it is worth 2 %. So: only for grids of at least four waves
    // (PIMDB_PAIR_DUAL=0/1 forces it off / on).
    static const char* dual_env = getenv("PIMDB_PAIR_DUAL");
    const double dual_waves = (double)nb * s->n_dual / kPairWarps / (2.0 * s->sm_count);
    const bool dual_wanted = dual_env ? atoi(dual_env) != 0 : dual_waves >= 4.0;
    const bool dual = !with_obs && dual_wanted && s->n_dual > 0 && !(s->rc > 0.0);
    if (dual) { a.tile_ij = s->rest_ij; a.TP = s->n_rest; }
    const long long items = (long long)nb * a.TP;
    // `split` > 1 cuts a tile pair's rotations over 2 or 4 warps. Measured on B200 at C3 (2.45 waves of tile pairs):
    // 52.0 / 52.2 / 58.3 us for split 1 / 2 / 4 -- the tail is not what limits the kernel, so the default stays 1.
    a.split = 1;
    a.tl = with_obs ? nullptr : tl_slot(s);
    if (const char* e = getenv("PIMDB_PAIR_SPLIT")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) a.split = v; }
    const long long gwarps = items * a.split;
    const int grid = (int)((gwarps + kPairWarps - 1) / kPairWarps);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (s->timing && !with_obs) {
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, s->stream);
    }
    if (with_obs) {
        if (s->D == 1) dispatch1<1, true>(s, a, grid, false);
        else if (s->D == 2) dispatch1<2, true>(s, a, grid, false);
        else dispatch1<3, true>(s, a, grid, false);
        k_pair_obs_reduce<<<1, 1024, 0, s->stream>>>(s->pair_scratch, (long long)grid * kPairWarps, &s->obs_d->pair_v, &s->obs_d->pair_vir);
        s->launches += 2;
    } else {
        if (dual) {
            PairArgs ad = a;
            ad.tl = a.tl; a.tl = tl_slot(s);
            const int gd = (int)(((long long)nb * s->n_dual + kPairWarps - 1) / kPairWarps);
            dispatch_dual(s, ad, gd, s->n_dual, early);
            s->launches += 1;
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing(s->stream, &cap);
            early = cap == cudaStreamCaptureStatusActive;     // the single items ride behind the dual ones
        }
        if (s->D == 1) dispatch1<1, false>(s, a, grid, early);
        else if (s->D == 2) dispatch1<2, false>(s, a, grid, early);
        else dispatch1<3, false>(s, a, grid, early);
        s->launches += 1;
    }
    if (e0) {
        cudaEventRecord(e1, s->stream);
        s->ev_pair.emplace_back(e0, e1);
    }
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_pair_chunk(Sim* s, int bead_lo, int nb, bool with_obs, bool early, int scratch_lo) {
    return launch_chunk(s, bead_lo, nb, with_obs, early, scratch_lo);
}

}  // namespace pimdb
