// Pair forces (K1) and the pair part of the energy estimators (K13), FP64, sm_100a.
//
// Replaces the i<j loop of Simulation::updatePhysicalForces (reference src/simulation.cpp:432-454) with
// getSeparation / applyMinimumImage (src/simulation.cpp:499-512, src/common.cpp:41-43) and
// AzizPotential / DipolePotential / HarmonicPotential ::gradV, ::V (src/potentials/*.cpp), and the second
// pair loop of EnergyObservable::calculatePotential (src/observables/energy.cpp:76-94).
//
// Decomposition: the N particles of a bead are cut into T = ceil(N/32) tiles. One WARP owns one unordered tile pair
// (I<=J) of one bead (or one half / quarter of its rotations, `split`, an experiment switch). Lane l holds particle
// 32I+l ("i") in registers; the j tile and its reaction-force accumulators live in shared memory, and the 32 rotations
// (l, (l+t)%32) enumerate the 32x32 tile: in a rotation every lane touches a different j, so the shared-memory
// reads and read-modify-writes are conflict-free and need no atomics. Every unordered pair is evaluated exactly once
// (Newton's third law). Diagonal tiles run the rotations t=1..16 only (t=16 on half the lanes). Force-only launches
// take TWO rotations per loop step (pair_rotation2: two independent pair chains per lane, two reaction accumulators).
// The warps that share a tile pair combine their partial sums through shared memory in a fixed order, then ONE of them
// writes the two 32-particle partial force vectors to a scratch slab S[bead][tile K][other tile M][axis][lane]; the
// assembly (k_assemble, or the closing k_integrate) sums the T partials of every particle in fixed order M=0..T-1
// (deterministic, bit-reproducible) and adds the external and spring forces.
// Work items are ordered tile pair by tile pair (all beads of one pair are neighbours), diagonal tiles last; blocks of
// 4 warps, 5 per SM.
#include "internal.cuh"
#include "device_utils.cuh"

namespace pimdb {

#ifndef PIMDB_PAIR_WARPS
#define PIMDB_PAIR_WARPS 4
#endif
#ifndef PIMDB_PAIR_MINBLOCKS
#define PIMDB_PAIR_MINBLOCKS 5
#endif
#ifndef PIMDB_PAIR_ROT2
#define PIMDB_PAIR_ROT2 1       // force-only launches take two rotations per loop step (pair_rotation2)
#endif
constexpr int kPairWarps = PIMDB_PAIR_WARPS;   // warps per block

struct PairArgs {
    const double* x;      // first bead of this launch, slab stride S
    double* scratch;      // [nb][T][T][D][32]
    double* obs_part;     // [items * split][2] (V, virial) when OBS
    const ushort2* tile_ij;
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
        const double e = fma(-r2, y, 1.0);                 // one third-order step: y (1 + e + e^2), remainder e^3 <= 2^-60
        const double u = fma(y, fma(e, e, e), y);
        const double u2 = u * u;
        g = (u2 * u2) * fma(fma(a.az_h2, u, a.az_h1), u, a.az_h0);
    } else {
        g = pair_eval<POT, OBS>(r2, a, v);
    }
    if (MASKED || CUT) {
        if (!active) { g = 0.0; v = 0.0; }
    }
    // force on i: -g d;  reaction on j: +g d  (every lane owns a distinct j slot in this rotation)
    double of[D];   // (loads first, stores after: see pair_rotation2)
#pragma unroll
    for (int c = 0; c < D; ++c) of[c] = sf[c * 32 + src];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        fi[c] = fma(-g, d[c], fi[c]);
        of[c] = fma(g, d[c], of[c]);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) sf[c * 32 + src] = of[c];
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

// shared-memory accesses by 32-bit window address (volatile: kept in program order among themselves)
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// Aziz, damped branch (x < D) and hard core, from quantities the common branch has already formed: see pair_eval.
__device__ __forceinline__ double aziz_damped(double r, double ir, double e1, const PairArgs& a) {
    const double xs = r * (1.0 / kAzRm);
    double w;
    if (xs > kEps && xs < 0.01) {
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
    }
    return w * (ir * (kAzEps / kAzRm));
}

// Two rotations (t, t+1) in one loop step, force only. One pair evaluation is a chain of dependent FP64 operations
// (~230 cycles for ~75 issue slots), and the registers allow ~5 warps per scheduler: with one pair in flight per lane the
// schedulers issue 58 % of the time (ncu, round 1). Here every lane carries TWO independent pairs -- the same i particle
// against the j slots (l+t) and (l+t+1) -- through straight-line code, so the two chains interleave. The reaction forces
// of the two rotations go to two separate shared-memory accumulators (sfa for even, sfb for odd rotations of the run):
// within a step every lane then owns one slot of each array, the read-modify-writes stay conflict-free and one
// __syncwarp per step orders them against the next step. The arithmetic of every single pair is that of pair_rotation.
template <int D, int POT, bool PBC, bool CUT, bool MASKED>
__device__ __forceinline__ void pair_rotation2(const PairArgs& a, int t, int lane, bool diag, bool vi, int jbase,
                                               const double (&xi)[D], double (&fi)[D], unsigned wsh) {
    // wsh: shared-window address of this warp's arrays x | f (even rotations) | g (odd rotations), D*32 doubles each.
    // Explicit ld/st.shared with immediate offsets: two address registers per step instead of re-deriving six generic
    // addresses (that was ~15 of the ~190 instructions of a step).
    const int sa = (lane + t) & 31, sb = (lane + t + 1) & 31;
    const unsigned aa = wsh + 8u * sa, ab = wsh + 8u * sb;
    constexpr int OF = D * 256, OG = 2 * D * 256;      // byte offsets of f and g
    double da[D], db[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        da[c] = xi[c] - lds_f64(aa + c * 256);
        db[c] = xi[c] - lds_f64(ab + c * 256);
    }
    // Minimum image: min_image_vec's arithmetic for both vectors. A suspected tie (a component within 2^-19 of +-L/2)
    // takes the exact, out-of-line evaluation of the reference's expression. For Aziz the decision rides on the far/near
    // vote below (a tie sends the rotation down the general path, which first redoes the separations exactly), so the
    // common path has no branch of its own for it.
    constexpr bool DEFER_TIE = PBC && POT == PIMDB_POT_AZIZ;
    bool tie = false;
    if (PBC) {
        double wa[D], wb[D];
        int worst = 0;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            wa[c] = fma(-a.L, rint(da[c] * a.invL), da[c]);
            wb[c] = fma(-a.L, rint(db[c] * a.invL), db[c]);
            worst = max(worst, max(__double2hiint(wa[c]) & 0x7fffffff, __double2hiint(wb[c]) & 0x7fffffff));
        }
        tie = worst >= a.tie_hi;
        if (!DEFER_TIE && tie) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
                da[c] = min_image_exact(da[c], a.L, MASKED && diag && (sa < lane));
                db[c] = min_image_exact(db[c], a.L, MASKED && diag && (sb < lane));
            }
        } else {
#pragma unroll
            for (int c = 0; c < D; ++c) { da[c] = wa[c]; db[c] = wb[c]; }
        }
    }
    double ra, rb;
    bool acta, actb;
    auto norms_and_masks = [&]() {
        ra = da[0] * da[0]; rb = db[0] * db[0];
#pragma unroll
        for (int c = 1; c < D; ++c) { ra = fma(da[c], da[c], ra); rb = fma(db[c], db[c], rb); }
        acta = true; actb = true;
        if (MASKED) {
            acta = vi && (jbase + sa < a.N) && !(diag && t == 16 && lane >= 16);
            actb = vi && (jbase + sb < a.N) && !(diag && t + 1 == 16 && lane >= 16);
        }
        if (CUT) {
            acta = acta && (sqrt(ra) < a.rc);
            actb = actb && (sqrt(rb) < a.rc);
        }
        if (MASKED || CUT) {
            if (!acta) ra = 1.0;
            if (!actb) rb = 1.0;
        }
    };
    norms_and_masks();
    double ga, gb;
    if (POT == PIMDB_POT_AZIZ) {
        if (__all_sync(kFullMask, ra > a.az_far2 && rb > a.az_far2 && !tie)) {       // see pair_rotation
            double ya, yb;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ya) : "d"(ra));
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(yb) : "d"(rb));
            // one third-order step each: y (1 + e + e^2), e = 1 - r y; the seed is good to 2^-20, the remainder e^3 to 2^-60
            const double ea = fma(-ra, ya, 1.0), eb = fma(-rb, yb, 1.0);
            const double ua = fma(ya, fma(ea, ea, ea), ya), ub = fma(yb, fma(eb, eb, eb), yb);
            const double qa = ua * ua, qb = ub * ub;
            ga = (qa * qa) * fma(fma(a.az_h2, ua, a.az_h1), ua, a.az_h0);
            gb = (qb * qb) * fma(fma(a.az_h2, ub, a.az_h1), ub, a.az_h0);
        } else {
            if (DEFER_TIE && __any_sync(kFullMask, tie)) {      // rare: the reference's expression, literally, for the whole rotation
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    da[c] = min_image_exact(xi[c] - lds_f64(aa + c * 256), a.L, MASKED && diag && (sa < lane));
                    db[c] = min_image_exact(xi[c] - lds_f64(ab + c * 256), a.L, MASKED && diag && (sb < lane));
                }
                norms_and_masks();
            }
            // the undamped branch of pair_eval for both pairs, unconditionally; lanes inside x < D redo theirs
            const double ira = rsqrt_fast(ra), irb = rsqrt_fast(rb);
            const double sra = ra * ira, srb = rb * irb;
            const double ea = exp2_lin_fast(sra, a.az_k2, a.ek), eb = exp2_lin_fast(srb, a.az_k2, a.ek);
            const double ua = ira * ira, ub = irb * irb, ua2 = ua * ua, ub2 = ub * ub;
            const double Pa = (ua2 * ua2) * fma(fma(a.az_h2, ua, a.az_h1), ua, a.az_h0);
            const double Pb = (ub2 * ub2) * fma(fma(a.az_h2, ub, a.az_h1), ub, a.az_h0);
            ga = fma(a.az_ca, ea * ira, Pa);
            gb = fma(a.az_ca, eb * irb, Pb);
            const bool dampa = !(sra >= a.az_drm), dampb = !(srb >= a.az_drm);
            if (__any_sync(kFullMask, dampa || dampb)) {
                if (dampa) ga = aziz_damped(sra, ira, ea, a);
                if (dampb) gb = aziz_damped(srb, irb, eb, a);
            }
        }
    } else {
        double v;
        ga = pair_eval<POT, false>(ra, a, v);
        gb = pair_eval<POT, false>(rb, a, v);
    }
    if (MASKED || CUT) {
        if (!acta) ga = 0.0;
        if (!actb) gb = 0.0;
    }
    // reaction forces: all six loads first, then the six stores -- written as load/FMA/store per component the compiler
    // keeps that order (it cannot prove the slots distinct) and a step pays the shared-memory latency six times (ncu: the
    // six FMAs behind the loads collected a quarter of the kernel's stall samples)
    double oa[D], ob[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        oa[c] = lds_f64(aa + OF + c * 256);
        ob[c] = lds_f64(ab + OG + c * 256);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) fi[c] = fma(-gb, db[c], fma(-ga, da[c], fi[c]));
#pragma unroll
    for (int c = 0; c < D; ++c) {
        sts_f64(aa + OF + c * 256, fma(ga, da[c], oa[c]));
        sts_f64(ab + OG + c * 256, fma(gb, db[c], ob[c]));
    }
    __syncwarp();
}

template <int D, int POT, bool PBC, bool CUT, bool OBS>
__global__ void __launch_bounds__(32 * kPairWarps, PIMDB_PAIR_MINBLOCKS) k_pair_tiles(PairArgs a) {
    constexpr bool ROT2 = PIMDB_PAIR_ROT2 && !OBS;
    __shared__ __align__(16) double s_w[kPairWarps][3][D * 32];    // per warp: j tile | reaction forces | (ROT2) those of the odd rotations
    grid_launch_dependents();   // a launch chained behind this one (the next slice of the pair tiles) may be scheduled as soon as
                                // every block of this grid has started
    tl_begin(a.tl);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sp = a.split;
    const long long gw = (long long)blockIdx.x * kPairWarps + warp;
    const long long item = gw / sp;
    const int part = (int)(gw - item * sp);
    // Items run tile pair by tile pair (all beads of one pair are neighbours), the half-size diagonal tiles last: blocks
    // are dispatched in index order, so the last wave consists of the shortest items and the SMs finish closer together.
    const bool live = item < (long long)a.nb * a.TP;      // no early exit: every warp reaches the block barrier
    const int bl = live ? (int)(item % a.nb) : 0;
    const ushort2 ij = a.tile_ij[live ? item / a.nb : 0];
    const int I = ij.x, J = ij.y;
    const bool diag = (I == J);
    const double* xb = a.x + (size_t)bl * a.S;
    const int pi = I * kTile + lane, pj = J * kTile + lane;
    const bool vi = pi < a.N, vj = pj < a.N;
    double* sx = s_w[warp][0];
    double* sf = s_w[warp][1];
    double* sg = s_w[warp][2];
    const unsigned wsh = (unsigned)__cvta_generic_to_shared(sx);

    double xi[D], fi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        xi[c] = vi ? xb[(size_t)c * a.N + pi] : 0.0;
        sx[c * 32 + lane] = vj ? xb[(size_t)c * a.N + pj] : 0.0;
        sf[c * 32 + lane] = 0.0;
        if (ROT2) sg[c * 32 + lane] = 0.0;
        fi[c] = 0.0;
    }
    __syncwarp();
    double vsum = 0.0, virsum = 0.0;
    const int jbase = J * kTile;

    if (live) {
        // my share of the rotations: t = 0..31 (off-diagonal) or 1..16 (diagonal), cut into `sp` equal runs (even lengths)
        const int nrot = (diag ? 16 : 32) / sp;
        const int tb = (diag ? 1 : 0) + part * nrot;
        if (!diag && (J + 1) * kTile <= a.N) {   // I < J, so tile I is full as well
            if constexpr (ROT2) {
#pragma unroll 1
                for (int t = tb; t < tb + nrot; t += 2)
                    pair_rotation2<D, POT, PBC, CUT, false>(a, t, lane, false, true, jbase, xi, fi, wsh);
            } else {
#pragma unroll 1
                for (int t = tb; t < tb + nrot; ++t)
                    pair_rotation<D, POT, PBC, CUT, OBS, false>(a, t, lane, false, true, jbase, xi, fi, vsum, virsum, sx, sf);
            }
        } else {
            if constexpr (ROT2) {
#pragma unroll 1
                for (int t = tb; t < tb + nrot; t += 2)
                    pair_rotation2<D, POT, PBC, CUT, true>(a, t, lane, diag, vi, jbase, xi, fi, wsh);
            } else {
#pragma unroll 1
                for (int t = tb; t < tb + nrot; ++t)
                    pair_rotation<D, POT, PBC, CUT, OBS, true>(a, t, lane, diag, vi, jbase, xi, fi, vsum, virsum, sx, sf);
            }
        }
    }
    if (ROT2) {     // (own slots only: the last step's __syncwarp has ordered every lane's updates)
#pragma unroll
        for (int c = 0; c < D; ++c) sf[c * 32 + lane] += sg[c * 32 + lane];
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
                    fi[c] += s_w[warp + q][0][c * 32 + lane];
                    fj[c] += s_w[warp + q][1][c * 32 + lane];
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
// `scratch_lo`: the slot of the scratch slab that bead `bead_lo` writes to (several launches can fill one slab)
static int launch_chunk(Sim* s, int bead_lo, int nb, bool with_obs, bool early, int scratch_lo) {
    PairArgs a;
    a.x = s->x + (size_t)(bead_lo + 1) * s->S;
    a.scratch = s->pair_scratch + (size_t)scratch_lo * s->T * s->T * s->D * kTile;
    a.obs_part = s->pair_scratch;  // the scratch slab doubles as the (V, virial) partial buffer
    a.tile_ij = s->tile_ij;
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
    const long long items = (long long)nb * a.TP;
    // `split` > 1 cuts a tile pair's rotations over 2 or 4 warps (PIMDB_PAIR_SPLIT, an experiment switch). Measured on B200 at
    // C3: 43.1 / 45.2 us (pair tiles + assembly) for split 1 / 2 -- finer items do not pay for the extra combination step.
    a.split = 1;
    a.tl = with_obs ? nullptr : tl_slot(s, 5);
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

// warps of pair tiles one SM holds at a time (api.cu sizes the slices of a large grid in waves of these)
int pair_resident_warps_per_sm() { return kPairWarps * PIMDB_PAIR_MINBLOCKS; }

int launch_pair_chunk(Sim* s, int bead_lo, int nb, bool with_obs, bool early, int scratch_lo) {
    return launch_chunk(s, bead_lo, nb, with_obs, early, scratch_lo);
}

}  // namespace pimdb
