// Normal-mode transforms (K12): the NormalModesPropagator step and the normal-mode-coupled Langevin thermostat,
// each fused into ONE kernel. FP64.
//
// Reference: src/normal_modes.cpp:48-129 (matrix rows, length-P dot products over the bead axis, window layout
// [axis][atom][bead]), src/propagators/normal_modes_propagator.cpp:19-103 (half kick with the physical forces,
// exact free-ring-polymer rotation per mode, back transform), src/thermostats/thermostat_coupling.cpp:29-47 and
// src/thermostats/langevin.cpp:15-27 (O step on the normal-mode momenta).
//
// The reference gives every MPI rank one row of the matrix and synchronises four shared-memory windows with
// barriers. Here a thread block stages a [P beads] x [TC columns] tile of x and p in shared memory (a "column"
// is one (axis, particle) pair; the bead stride is D*N so tile rows are coalesced), applies the P x P matrix,
// the per-mode update and the inverse matrix without leaving the SM, and writes the tile back: 1 read + 1 write
// of x and p per transform pair instead of 4 window round trips. Sums run over j = 0..P-1 in the reference's order.
#include "internal.cuh"
#include "device_utils.cuh"

namespace pimdb {

struct NmArgs {
    double *x, *p;             // x: first owned slab (halo already skipped); p: first slab
    const double* fphys;       // physical forces for the half kick (propagator) or nullptr
    const double* C;           // [P][P] forward rows: C[k][j]
    const double* Cinv;        // [P][P] inverse rows as the reference stores them: Cinv[j][k]
    const double* tab;         // [3][P]: cos(w_k dt), sin(w_k dt), m w_k
    unsigned long long* draw; unsigned int* ticket;
    int P, M, N, D, TC;
    double hdt, dt_over_m, c1, c2;
    unsigned long long seed;
    const double* noise;       // reference-compatible stream: [mode][N][D] gaussians of this half-step, or nullptr (Philox)
    // bead shard: the tile is staged from the gathered slabs of ALL beads (gx / gp, [P][M]; the propagator's half kick is
    // already in gp) and only the owned beads [b0, b1) are written back, to x / p of this handle. nullptr on a full ring.
    const double *gx, *gp;
    int b0, b1;
    const unsigned int* gather_flag; const unsigned int* gather_seq; int world; unsigned long long timeout_ns; int* err; unsigned* dead;
};

// mode 0: propagator (x and p), mode 1: Langevin thermostat on the mode momenta (p only),
// mode 2 / 3: forward / inverse transform of p alone (in place) -- the two halves around a Nose-Hoover update of the
// mode momenta, which needs a reduction over all particles of a mode and therefore its own kernels (nose_hoover.cu)
template <int MODE>
__global__ void __launch_bounds__(256) k_nm_fused(NmArgs a) {
    extern __shared__ double sm[];
    const int P = a.P, TC = a.TC;
    double* sx = sm;                    // [P][TC]
    double* sp = sm + (size_t)P * TC;   // [P][TC]
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const int tc = tid % TC, kg = tid / TC, KG = blockDim.x / TC;
    const unsigned long long draw = (MODE == 1) ? *a.draw : 0ull;
    const bool shard = a.gp != nullptr;
    const double *gx = a.gx, *gp = a.gp;
    if (shard) {   // every rank has delivered its beads of this gather (bounded wait, acquire)
        const unsigned g = *a.gather_seq;
        if (tid < a.world) wait_sys_u32_ge(&a.gather_flag[tid], g, a.timeout_ns, a.err, kErrPeerTimeout, a.dead);
        __syncthreads();
        const size_t slot = (size_t)(g & 1u) * 2 * (size_t)P * a.M;     // gathers alternate between two buffers
        gx += slot; gp += slot;
    }

    for (int col0 = blockIdx.x * TC; col0 < a.M; col0 += gridDim.x * TC) {
        const int col = col0 + tc;
        const bool ok = col < a.M;
        // 1. stage (propagator: with the half kick p += dt/2 f_phys, normal_modes_propagator.cpp:73-103)
        for (int j = kg; j < P; j += KG) {
            double pv = 0.0, xv = 0.0;
            if (ok && shard) {
                pv = gp[(size_t)j * a.M + col];
                if (MODE == 0) xv = gx[(size_t)j * a.M + col];
            } else if (ok) {
                pv = a.p[(size_t)j * a.M + col];
                if (MODE == 0) {
                    pv += a.hdt * a.fphys[(size_t)j * a.M + col];
                    xv = a.x[(size_t)j * a.M + col];
                }
            }
            sp[j * TC + tc] = pv;
            if (MODE == 0) sx[j * TC + tc] = xv;
        }
        __syncthreads();
        // 2. forward transform + per-mode update, results kept in registers until all reads are done
        constexpr int KMAX = 32;        // P*TC/256 <= 32 by construction of TC
        double rn_x[KMAX], rn_p[KMAX];
        int cnt = 0;
        for (int k = kg; k < P; k += KG, ++cnt) {
            const double* crow = a.C + (size_t)k * P;
            double xn = 0.0, pn = 0.0;
            if (MODE != 3) {
                for (int j = 0; j < P; ++j) {
                    const double cj = __ldg(crow + j);
                    pn += cj * sp[j * TC + tc];
                    if (MODE == 0) xn += cj * sx[j * TC + tc];
                }
            }
            if (MODE == 0) {
                const double cs = a.tab[k], sn = a.tab[P + k], mw = a.tab[2 * P + k];
                double xo, po;
                if (mw == 0.0) {                 // freq == 0: free drift (normal_modes_propagator.cpp:33-35)
                    xo = xn + a.dt_over_m * pn;
                    po = pn;
                } else {                         // exact harmonic rotation (:36-39)
                    xo = cs * xn + sn / mw * pn;
                    po = (-1.0) * mw * sn * xn + cs * pn;
                }
                rn_x[cnt] = xo;
                rn_p[cnt] = po;
            } else if (MODE == 2) {
                rn_p[cnt] = pn;
            } else if (MODE == 3) {
                rn_p[cnt] = sp[k * TC + tc];         // already mode momenta: pass through to the inverse transform
            } else {
                // Langevin O step on mode k: noise stream row = mode k * D + axis (DESIGN.md "RNG")
                const int axis = ok ? col / a.N : 0, n = ok ? col % a.N : 0;
                double z;
                if (a.noise) {   // the generator of rank k acts on mode k (thermostat_coupling.cpp:29-47), particle-major order
                    z = ok ? a.noise[((size_t)k * a.N + n) * a.D + axis] : 0.0;
                } else {
                    double z0, z1;
                    gaussian_pair((uint32_t)(n >> 1), (uint32_t)(k * a.D + axis), draw, a.seed, z0, z1);
                    z = (n & 1) ? z1 : z0;
                }
                rn_p[cnt] = a.c1 * pn + a.c2 * z;
            }
        }
        __syncthreads();
        cnt = 0;
        for (int k = kg; k < P; k += KG, ++cnt) {
            sp[k * TC + tc] = rn_p[cnt];
            if (MODE == 0) sx[k * TC + tc] = rn_x[cnt];
        }
        __syncthreads();
        // 3. inverse transform and store
        for (int j = kg; j < P; j += KG) {
            const double* irow = a.Cinv + (size_t)j * P;
            double xc = 0.0, pc = 0.0;
            if (MODE == 2) {
                pc = sp[j * TC + tc];                // forward only: row j now holds mode j
            } else {
                for (int k = 0; k < P; ++k) {
                    const double ck = __ldg(irow + k);
                    pc += ck * sp[k * TC + tc];
                    if (MODE == 0) xc += ck * sx[k * TC + tc];
                }
            }
            if (ok && j >= a.b0 && j < a.b1) {
                a.p[(size_t)(j - a.b0) * a.M + col] = pc;
                if (MODE == 0) a.x[(size_t)(j - a.b0) * a.M + col] = xc;
            }
        }
        __syncthreads();
    }
    if (MODE == 1) {   // advance the noise draw counter once all blocks have read it
        if (tid == 0) {
            __threadfence();
            unsigned int t = atomicAdd(a.ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last && tid == 0) {
            *a.ticket = 0u;
            *a.draw = draw + 1ull;
        }
    }
}

static int launch_nm(Sim* s, int mode) {
    NmArgs a;
    a.x = s->x + s->S; a.p = s->p; a.fphys = s->fp;
    a.gx = a.gp = nullptr; a.b0 = 0; a.b1 = s->P;
    a.gather_flag = nullptr; a.gather_seq = nullptr; a.world = 1; a.timeout_ns = 0; a.err = s->err_d; a.dead = nullptr;
    if (!s->all_local) {
        if (!s->peer_on || !s->peer.gather_mine) {
            s->err = "normal-mode transforms on a bead shard need the handle attached to its peers (pimdb_peer_attach)";
            return PIMDB_ERR_INVALID_ARGUMENT;
        }
        if (mode >= 2) { s->err = "Nose-Hoover chains on the normal modes need all beads on one handle"; return PIMDB_ERR_INVALID_ARGUMENT; }
        // every rank sends its beads to every rank (momenta with the propagator's half kick applied), then transforms all modes
        int rc = launch_peer_allgather(s, mode == 0, mode == 0);
        if (rc != PIMDB_OK) return rc;
        const size_t PS = (size_t)s->P * s->S;
        // (the slot of the gather just launched: the kernel reads seq[4] after that launch has advanced it)
        a.gx = s->peer.gather_mine; a.gp = s->peer.gather_mine + PS;
        a.b0 = s->b0; a.b1 = s->b1;
        a.gather_flag = s->peer.mine->gather_flag; a.gather_seq = s->peer.seq + 4; a.world = s->peer.world;
        a.timeout_ns = s->peer.timeout_ns; a.dead = s->peer.seq + 3;
    }
    a.C = s->nmC; a.Cinv = s->nmC + (size_t)s->P * s->P; a.tab = s->nmFreq;
    a.draw = s->draw; a.ticket = s->tickets;
    a.P = s->P; a.M = (int)s->S; a.N = s->N; a.D = s->D;
    // tile width: P*TC/256 outputs per thread must stay <= 32 and the two tiles must fit in shared memory
    int TC = 32;
    while (TC > 1 && ((size_t)s->P * TC / 256 > 32 || (size_t)2 * s->P * TC * sizeof(double) > 200 * 1024)) TC >>= 1;
    // small systems: narrower tiles give more blocks (the chain per output is a P-long dot product either way)
    while (TC > 2 && (a.M + TC - 1) / TC < 2 * kNumSM && (size_t)s->P * TC >= 256) TC >>= 1;
    a.TC = TC;
    a.hdt = 0.5 * s->cfg.dt; a.dt_over_m = s->cfg.dt / s->cfg.mass; a.c1 = s->c1; a.c2 = s->c2;
    a.seed = s->cfg.seed;
    a.noise = nullptr;
    if (mode == 1 && s->rm_state) {
        int rc = launch_ranmars_fill(s);
        if (rc != PIMDB_OK) return rc;
        a.noise = s->rm_noise;
    }
    const size_t smem = (size_t)2 * s->P * TC * sizeof(double);
    const int grid = grid_for((a.M + TC - 1) / TC, 1, 4 * kNumSM);
    if (mode == 0) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_nm_fused<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_nm_fused<0><<<grid, 256, smem, s->stream>>>(a);
    } else if (mode == 2) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_nm_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_nm_fused<2><<<grid, 256, smem, s->stream>>>(a);
    } else if (mode == 3) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_nm_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_nm_fused<3><<<grid, 256, smem, s->stream>>>(a);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_nm_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_nm_fused<1><<<grid, 256, smem, s->stream>>>(a);
    }
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_nm_propagate(Sim* s) { return launch_nm(s, 0); }
int launch_nm_thermostat(Sim* s) { return launch_nm(s, 1); }
int launch_nm_momenta(Sim* s, bool forward) { return launch_nm(s, forward ? 2 : 3); }

// ------------------------------------------------------------------------------------------------------
// Element-wise estimator partials (K13): classical spring energy of the owned links, external potential and its
// virial, sum p^2. Reference: Simulation::classicalSpringEnergy (src/simulation.cpp:464-486),
// EnergyObservable::calculatePotential external part (src/observables/energy.cpp:62-74),
// ClassicalObservable::calculateKineticEnergy (src/observables/classical.cpp:32-45).
struct ObsArgs {
    const double *x, *p;
    double* part; unsigned int* ticket; DevObs* obs;
    int N, D, Ploc, skip_link;   // skip_link: owned-bead index whose incoming link is the exterior (bosonic) one, or -1
    size_t S;
    double k, kext, L, invL, mass;
    int pbc, ext_pot;
    double ext_a, ext_b;
    int bead_begin;
};

__global__ void __launch_bounds__(256) k_obs_elementwise(ObsArgs a) {
    __shared__ double sm[8 * 32];
    __shared__ bool is_last;
    const long long total = (long long)a.Ploc * a.N;
    // spring d^2, ext V, ext virial, p^2 | GSF (src/observables/gsf_action.cpp:21-73): V_ext on odd / even beads,
    // |grad V_ext|^2 on odd / even beads (bead parity is that of the GLOBAL bead index, `this_bead`)
    double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / a.N), n = (int)(idx % a.N);
        const double* xc = a.x + (size_t)(b + 1) * a.S;
        const double* xp = xc - a.S;
        const double* pc = a.p + (size_t)b * a.S;
        double r2 = 0.0, d2 = 0.0, pp = 0.0;
        double xv[3] = {0.0, 0.0, 0.0};
        for (int c = 0; c < a.D; ++c) {
            const double xi = xc[(size_t)c * a.N + n];
            xv[c] = xi;
            double d = xp[(size_t)c * a.N + n] - xi;
            if (a.pbc) d = min_image(d, a.L, a.invL);
            d2 = fma(d, d, d2);
            r2 = fma(xi, xi, r2);
            const double pi = pc[(size_t)c * a.N + n];
            pp = fma(pi, pi, pp);
        }
        if (b != a.skip_link) acc[0] += d2;
        double vext = 0.0, g2 = 0.0;      // this particle's V_ext and |grad V_ext|^2 (GSF)
        if (a.ext_pot == PIMDB_POT_HARMONIC) {
            acc[1] += r2;                 // V = k/2 sum x^2 ; virial -x.F = k sum x^2 (scaled at the end)
            vext = 0.5 * a.kext * r2;
            g2 = a.kext * a.kext * r2;
        } else if (a.ext_pot == PIMDB_POT_DOUBLE_WELL) {
            // reference src/potentials/double_well.cpp:6-40: V = m lambda sum_c (x_c^2 - a^2)^2 ; grad = 4 m lambda (|x|^2 - a^2) x
            double v = 0.0;
            for (int c = 0; c < a.D; ++c) { double t = xv[c] * xv[c] - a.ext_b * a.ext_b; v += t * t; }
            vext = a.mass * a.ext_a * v;
            acc[1] += vext;
            const double pref = 4.0 * a.mass * a.ext_a * (r2 - a.ext_b * a.ext_b);
            acc[2] += pref * r2;
            g2 = pref * pref * r2;
        } else if (a.ext_pot == PIMDB_POT_COSINE) {
            // reference src/potentials/cosine.cpp:9-35
            const double kk = 2.0 * M_PI / a.L;
            for (int c = 0; c < a.D; ++c) {
                const double vc = a.ext_a * cos(kk * xv[c] + a.ext_b), gc = -a.ext_a * kk * sin(kk * xv[c] + a.ext_b);
                acc[1] += vc;
                acc[2] += xv[c] * gc;
                vext += vc;
                g2 += gc * gc;
            }
        }
        acc[3] += pp;
        const int odd = (a.bead_begin + b) & 1;
        acc[odd ? 4 : 5] += vext;
        acc[odd ? 6 : 7] += g2;
    }
    block_sum<8>(acc, sm);
    if (threadIdx.x == 0) {
        for (int c = 0; c < 8; ++c) a.part[blockIdx.x * 8 + c] = acc[c];
        __threadfence();
        unsigned int t = atomicAdd(a.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double tot[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int blk = threadIdx.x; blk < (int)gridDim.x; blk += blockDim.x)
            for (int c = 0; c < 8; ++c) tot[c] += __ldcg(&a.part[blk * 8 + c]);
        block_sum<8>(tot, sm);
        if (threadIdx.x == 0) {
            a.obs->spring_e[0] = 0.5 * a.k * tot[0];
            if (a.ext_pot == PIMDB_POT_HARMONIC) {
                a.obs->ext_v = 0.5 * a.kext * tot[1];
                a.obs->ext_vir = a.kext * tot[1];
            } else {
                a.obs->ext_v = tot[1];
                a.obs->ext_vir = tot[2];
            }
            a.obs->p2 = tot[3];
            for (int c = 0; c < 4; ++c) a.obs->gsf[c] = tot[4 + c];
            *a.ticket = 0u;
        }
    }
}

int launch_obs_elementwise(Sim* s) {
    ObsArgs a;
    a.x = s->x; a.p = s->p; a.part = s->obs_part; a.ticket = s->tickets; a.obs = s->obs_d;
    a.N = s->N; a.D = s->D; a.Ploc = s->Ploc;
    a.skip_link = (s->bosonic && s->has_first) ? 0 : -1;
    a.S = s->S; a.k = s->kspring; a.kext = s->kext; a.L = s->L; a.invL = 1.0 / s->L; a.mass = s->cfg.mass;
    a.pbc = s->cfg.pbc; a.ext_pot = s->cfg.ext_potential;
    a.bead_begin = s->b0;
    if (s->cfg.ext_potential == PIMDB_POT_DOUBLE_WELL) { a.ext_a = s->cfg.ext_strength; a.ext_b = s->cfg.ext_location; }
    else { a.ext_a = s->cfg.ext_amplitude; a.ext_b = s->cfg.ext_phase; }
    const int grid = grid_for((size_t)s->Ploc * s->N, 256, kMaxPartials);
    k_obs_elementwise<<<grid, 256, 0, s->stream>>>(a);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

}  // namespace pimdb
