// Pair forces (K1) and the pair part of the energy estimators (K13), FP64, sm_100a.
//
// Replaces the i<j loop of Simulation::updatePhysicalForces (reference src/simulation.cpp:432-454) with
// getSeparation / applyMinimumImage (src/simulation.cpp:499-512, src/common.cpp:41-43) and
// AzizPotential / DipolePotential / HarmonicPotential ::gradV, ::V (src/potentials/*.cpp), and the second
// pair loop of EnergyObservable::calculatePotential (src/observables/energy.cpp:76-94).
//
// Decomposition: the N particles of a bead are cut into T = ceil(N/32) tiles. One WARP owns one unordered
// tile pair (I<=J) of one bead: lane l holds particle 32I+l ("i") and particle 32J+l ("j") in registers and the
// 32 rotations (l, (l+t)%32) enumerate the 32x32 tile through warp shuffles -- no shared memory, no atomics.
// Every unordered pair is evaluated exactly once (Newton's third law): the force on i accumulates in the
// owning lane, the reaction on j travels back to j's home lane by a second shuffle. Diagonal tiles run the
// rotations t=1..16 only (t=16 on half the lanes). Each warp writes its two 32-particle partial force vectors
// to a scratch slab S[bead][tile K][other tile M][axis][lane]; the assemble kernel sums the T partials of every
// particle in fixed order M=0..T-1 (deterministic, bit-reproducible) and adds the external and spring forces.
#include "internal.cuh"
#include "device_utils.cuh"

namespace pimdb {

#ifndef PIMDB_PAIR_MINBLOCKS
#define PIMDB_PAIR_MINBLOCKS 3
#endif
struct PairArgs {
    const double* x;      // first bead of this launch, slab stride S
    double* scratch;      // [nb][T][T][D][32]
    double* obs_part;     // [items][2] (V, virial) when OBS
    const ushort2* tile_ij;
    int N, T, TP, nb;
    size_t S;
    double L, invL, rc, par;
};

// dV/dr / r (so that grad V = g * r_vec) and optionally V, from r^2.
template <int POT, bool WANT_V>
__device__ __forceinline__ double pair_eval(double r2, double par, double& v) {
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
        double ir = rsqrt_fast(r2);
        double r = r2 * ir;
        double xs = r * (1.0 / kAzRm);
        double e1 = exp_neg_fast(-kAzAlpha * xs);
        double t1 = -kAzA * kAzAlpha * e1;
        double dvdr;
        if (xs > kEps && xs < 0.01) {           // hard-core branch: repulsion only (aziz.cpp:39-42, 78-79)
            dvdr = t1 * (kAzEps / kAzRm);
            if (WANT_V) v = kAzEps * kAzA * e1;
        } else {
            double ix = kAzRm * ir;
            double ix2 = ix * ix, ix6 = ix2 * ix2 * ix2, ix8 = ix6 * ix2, ix10 = ix8 * ix2;
            double fdamp = 1.0, dfdamp = 0.0;
            if (xs < kAzD) {                    // include/potentials/aziz.h:25-34
                double q = kAzD * ix - 1.0;
                fdamp = exp_neg_fast(-q * q);
                dfdamp = 2.0 * kAzD * ix2 * q * fdamp;
            }
            double disp = kAzC6 * ix6 + kAzC8 * ix8 + kAzC10 * ix10;
            double ddisp = (6.0 * kAzC6 * ix6 + 8.0 * kAzC8 * ix8 + 10.0 * kAzC10 * ix10) * ix;
            dvdr = (kAzEps / kAzRm) * (t1 + ddisp * fdamp - disp * dfdamp);
            if (WANT_V) v = kAzEps * (kAzA * e1 - disp * fdamp);
        }
        return dvdr * ir;
    }
}

template <int D, int POT, bool PBC, bool CUT, bool OBS>
__global__ void __launch_bounds__(256, PIMDB_PAIR_MINBLOCKS) k_pair_tiles(PairArgs a) {
    const int lane = threadIdx.x & 31;
    const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (item >= (long long)a.nb * a.TP) return;  // whole warp exits together
    const int bl = (int)(item / a.TP);
    const ushort2 ij = a.tile_ij[item % a.TP];
    const int I = ij.x, J = ij.y;
    const bool diag = (I == J);
    const double* xb = a.x + (size_t)bl * a.S;
    const int pi = I * kTile + lane, pj = J * kTile + lane;
    const bool vi = pi < a.N, vj = pj < a.N;

    double xi[D], xj[D], fi[D], fj[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        xi[c] = vi ? xb[(size_t)c * a.N + pi] : 0.0;
        xj[c] = vj ? xb[(size_t)c * a.N + pj] : 0.0;
        fi[c] = 0.0;
        fj[c] = 0.0;
    }
    double vsum = 0.0, virsum = 0.0;
    const int t0 = diag ? 1 : 0, t1 = diag ? 16 : 31;
    const int jbase = J * kTile;

    for (int t = t0; t <= t1; ++t) {
        const int src = (lane + t) & 31;
        double d[D], xo[D];
        double r2 = 0.0;
        // The reference always forms x_lower_index - x_higher_index before the minimum image
        // (src/simulation.cpp:433-437, 503); mi(+L/2) = mi(-L/2) = -L/2, so the order matters for particles
        // exactly half a box apart (perfect lattices). Only diagonal tiles can see the higher index first.
        const bool swap = PBC && diag && (src < lane);
#pragma unroll
        for (int c = 0; c < D; ++c) {
            xo[c] = __shfl_sync(kFullMask, xj[c], src);
            d[c] = xi[c] - xo[c];
        }
        if (PBC) min_image_vec<D>(d, a.L, a.invL, swap);
#pragma unroll
        for (int c = 0; c < D; ++c) r2 = fma(d[c], d[c], r2);
        bool active = vi && (jbase + src < a.N) && !(diag && t == 16 && lane >= 16);
        if (CUT) active = active && (sqrt(r2) < a.rc);   // strict '<' (src/simulation.cpp:444)
        if (!active) r2 = 1.0;                            // keep the arithmetic finite on masked lanes
        double v = 0.0;
        double g = pair_eval<POT, OBS>(r2, a.par, v);
        if (!active) { g = 0.0; v = 0.0; }
        double fa[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            fa[c] = -g * d[c];                            // force on i
            fi[c] += fa[c];
        }
        if (OBS) {
            // energy.cpp:88-90: virial -= x_first . f_on_first, "first" = the lower particle index of the pair
            vsum += v;
            const bool i_first = !diag || (src > lane);
            double dot = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) dot += (i_first ? xi[c] : -xo[c]) * fa[c];
            virsum -= dot;
        }
        const int from = (lane - t) & 31;                 // the lane whose partner this lane's j was
#pragma unroll
        for (int c = 0; c < D; ++c) fj[c] -= __shfl_sync(kFullMask, fa[c], from);
    }

    if (!OBS) {
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
    } else {
        vsum = warp_sum(vsum);
        virsum = warp_sum(virsum);
        if (lane == 0) {
            a.obs_part[2 * item] = vsum;
            a.obs_part[2 * item + 1] = virsum;
        }
    }
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

template <int D, int POT, bool OBS>
static void dispatch2(Sim* s, const PairArgs& a, int grid) {
    const bool pbc = s->cfg.pbc != 0, cut = s->rc > 0.0;
    if (pbc && cut) k_pair_tiles<D, POT, true, true, OBS><<<grid, 256, 0, s->stream>>>(a);
    else if (pbc) k_pair_tiles<D, POT, true, false, OBS><<<grid, 256, 0, s->stream>>>(a);
    else if (cut) k_pair_tiles<D, POT, false, true, OBS><<<grid, 256, 0, s->stream>>>(a);
    else k_pair_tiles<D, POT, false, false, OBS><<<grid, 256, 0, s->stream>>>(a);
}

template <int D, bool OBS>
static void dispatch1(Sim* s, const PairArgs& a, int grid) {
    switch (s->cfg.int_potential) {
        case PIMDB_POT_AZIZ: dispatch2<D, PIMDB_POT_AZIZ, OBS>(s, a, grid); break;
        case PIMDB_POT_HARMONIC: dispatch2<D, PIMDB_POT_HARMONIC, OBS>(s, a, grid); break;
        case PIMDB_POT_DIPOLE: dispatch2<D, PIMDB_POT_DIPOLE, OBS>(s, a, grid); break;
        default: break;
    }
}

// Enqueue the tile kernel for owned beads [bead_lo, bead_lo+nb) (nb <= bead_chunk).
static int launch_chunk(Sim* s, int bead_lo, int nb, bool with_obs) {
    PairArgs a;
    a.x = s->x + (size_t)(bead_lo + 1) * s->S;
    a.scratch = s->pair_scratch;
    a.obs_part = s->pair_scratch;  // the scratch slab doubles as the (V, virial) partial buffer
    a.tile_ij = s->tile_ij;
    a.N = s->N; a.T = s->T; a.TP = s->TP; a.nb = nb; a.S = s->S;
    a.L = s->L; a.invL = 1.0 / s->L; a.rc = s->rc; a.par = s->pair_par;
    const long long items = (long long)nb * s->TP;
    const int grid = (int)((items * 32 + 255) / 256);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (s->timing && !with_obs) {
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, s->stream);
    }
    if (with_obs) {
        if (s->D == 1) dispatch1<1, true>(s, a, grid);
        else if (s->D == 2) dispatch1<2, true>(s, a, grid);
        else dispatch1<3, true>(s, a, grid);
        k_pair_obs_reduce<<<1, 1024, 0, s->stream>>>(s->pair_scratch, items, &s->obs_d->pair_v, &s->obs_d->pair_vir);
        s->launches += 2;
    } else {
        if (s->D == 1) dispatch1<1, false>(s, a, grid);
        else if (s->D == 2) dispatch1<2, false>(s, a, grid);
        else dispatch1<3, false>(s, a, grid);
        s->launches += 1;
    }
    if (e0) {
        cudaEventRecord(e1, s->stream);
        s->ev_pair.emplace_back(e0, e1);
    }
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_pair_chunk(Sim* s, int bead_lo, int nb, bool with_obs) { return launch_chunk(s, bead_lo, nb, with_obs); }

}  // namespace pimdb
