// Force assembly (K2, K3), the fused velocity-Verlet / Langevin / centre-of-mass kernels (K9-K11), halo fill
// and the AoS<->SoA boundary transposes. FP64, HBM-streaming, grid-stride with a grid capped at 8 x 148 blocks.
//
// Reference: Simulation::updateForces / updateSpringForces / updatePhysicalForces (src/simulation.cpp:353-455),
// Propagator::momentStep / coordsStep (src/propagators/velocity_verlet.cpp:24-38), LangevinThermostat
// (src/thermostats/langevin.cpp:10-27), Simulation::zeroMomentum (src/simulation.cpp:581-603).
#include "internal.cuh"
#include "device_utils.cuh"

namespace pimdb {

// ------------------------------------------------------------------------------------------------------
// Assemble: f = f_spring + f_phys for owned beads [bead_lo, bead_lo+nb).
//   f_phys   = -grad V_ext(x) + sum over the T pair-force partials of the particle's tile   (K2 + end of K1)
//   f_spring = k (mi(x_prev - x) + mi(x_next - x))  on classical beads                       (K3)
//            = exterior exchange force (exF) on bead 1 / bead P of a bosonic system          (K8 output)
struct AssembleArgs {
    const double* x;        // with halo: slab 0 = halo before first owned bead
    const double* scratch;  // pair partials of this chunk or nullptr
    const double* exF;      // [2][D][N] exterior forces (first, last)
    double *f, *fs, *fp;
    int N, D, T, bead_lo, nb;
    int first_local, last_local;   // owned-bead index of global bead 0 / P-1 when bosonic, else -1
    size_t S;
    double k, kext, L, invL;
    int pbc, ext_pot, write_split;
    unsigned long long* tl;
    double ext_a, ext_b, mass;     // double_well: strength, location ; cosine: amplitude, phase
};

template <int D>
__global__ void __launch_bounds__(256) k_assemble(AssembleArgs a) {
    tl_begin(a.tl);
    const long long total = (long long)a.nb * a.N;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int bl = (int)(idx / a.N), n = (int)(idx % a.N);
        const int b = a.bead_lo + bl;                       // owned-bead index
        const double* xc = a.x + (size_t)(b + 1) * a.S;
        double xv[D], phys[D], spring[D];
#pragma unroll
        for (int c = 0; c < D; ++c) xv[c] = xc[(size_t)c * a.N + n];
        // external force
        if (a.ext_pot == PIMDB_POT_HARMONIC) {
#pragma unroll
            for (int c = 0; c < D; ++c) phys[c] = -(a.kext * xv[c]);
        } else if (a.ext_pot == PIMDB_POT_DOUBLE_WELL) {
            // reference src/potentials/double_well.cpp:22-40: grad = 4 m lambda (|x|^2 - a^2) x
            double r2 = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) r2 += xv[c] * xv[c];
            double pref = 4.0 * a.mass * a.ext_a * (r2 - a.ext_b * a.ext_b);
#pragma unroll
            for (int c = 0; c < D; ++c) phys[c] = -(pref * xv[c]);
        } else if (a.ext_pot == PIMDB_POT_COSINE) {
            // reference src/potentials/cosine.cpp:9-35: V = A sum_c cos(k x_c + phase), k = 2 pi / L; F = A k sin(k x_c + phase)
            const double kk = 2.0 * M_PI / a.L;
#pragma unroll
            for (int c = 0; c < D; ++c) phys[c] = a.ext_a * kk * sin(kk * xv[c] + a.ext_b);
        } else {
#pragma unroll
            for (int c = 0; c < D; ++c) phys[c] = 0.0;
        }
        // pair partials, fixed order
        if (a.scratch) {
            const int K = n / kTile, lane = n % kTile;
            const double* s = a.scratch + ((size_t)bl * a.T + K) * a.T * D * kTile + lane;
            // (loads of 8 partials in flight at a time; the additions keep their order m = 0..T-1: bit-reproducible)
            int m = 0;
            for (; m + 8 <= a.T; m += 8) {
                double v[8][D];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
#pragma unroll
                    for (int c = 0; c < D; ++c) v[k][c] = s[((size_t)(m + k) * D + c) * kTile];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
#pragma unroll
                    for (int c = 0; c < D; ++c) phys[c] += v[k][c];
                }
            }
            for (; m < a.T; ++m) {
#pragma unroll
                for (int c = 0; c < D; ++c) phys[c] += s[((size_t)m * D + c) * kTile];
            }
        }
        // springs
        if (b == a.first_local || b == a.last_local) {
            const double* e = a.exF + (size_t)(b == a.first_local ? 0 : 1) * a.S;
#pragma unroll
            for (int c = 0; c < D; ++c) spring[c] = e[(size_t)c * a.N + n];
            // a system with P == 1 is never bosonic (src/simulation.cpp:690), so first != last here
        } else {
            const double* xp = xc - a.S;
            const double* xn = xc + a.S;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dp = xp[(size_t)c * a.N + n] - xv[c];
                double dn = xn[(size_t)c * a.N + n] - xv[c];
                if (a.pbc) {
                    dp = min_image(dp, a.L, a.invL);
                    dn = min_image(dn, a.L, a.invL);
                }
                spring[c] = a.k * (dp + dn);
            }
        }
        const size_t o = (size_t)b * a.S + n;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            a.f[o + (size_t)c * a.N] = spring[c] + phys[c];
            if (a.write_split) {
                a.fs[o + (size_t)c * a.N] = spring[c];
                a.fp[o + (size_t)c * a.N] = phys[c];
            }
        }
    }
    tl_end(a.tl);
}

int launch_assemble_chunk(Sim* s, int bead_lo, int nb, bool with_pair) {
    AssembleArgs a;
    a.x = s->x;
    a.scratch = with_pair ? s->pair_scratch : nullptr;
    a.exF = s->exF;
    a.f = s->f; a.fs = s->fs; a.fp = s->fp;
    a.N = s->N; a.D = s->D; a.T = s->T; a.bead_lo = bead_lo; a.nb = nb;
    a.first_local = (s->bosonic && s->has_first) ? 0 : -1;
    a.last_local = (s->bosonic && s->has_last) ? s->Ploc - 1 : -1;
    a.S = s->S;
    a.k = s->kspring; a.kext = s->kext; a.L = s->L; a.invL = 1.0 / s->L;
    a.pbc = s->cfg.pbc; a.ext_pot = s->cfg.ext_potential; a.write_split = 1;
    a.mass = s->cfg.mass;
    a.tl = tl_slot(s);
    if (s->cfg.ext_potential == PIMDB_POT_DOUBLE_WELL) { a.ext_a = s->cfg.ext_strength; a.ext_b = s->cfg.ext_location; }
    else { a.ext_a = s->cfg.ext_amplitude; a.ext_b = s->cfg.ext_phase; }
    const int grid = grid_for((size_t)nb * s->N, 256);
    if (s->D == 1) k_assemble<1><<<grid, 256, 0, s->stream>>>(a);
    else if (s->D == 2) k_assemble<2><<<grid, 256, 0, s->stream>>>(a);
    else k_assemble<3><<<grid, 256, 0, s->stream>>>(a);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

// ------------------------------------------------------------------------------------------------------
// Fused integrator. One thread owns the particle pair (2q, 2q+1) of a row (owned bead b, axis c); the
// stages run in the fixed order SUBCM -> O_PRE -> B -> O_POST -> A and are selected by `ops`:
//   SUBCM  p -= com/(N P)                       zeroMomentum, second half (subtract)
//   O_*    p  = c1 p + c2 xi                    LangevinThermostat::momentaUpdate (Cartesian coupling)
//   B      p += dt/2 f   (B_PHYS: f_phys only)  Propagator::momentStep
//   A      x += dt p / m                        Propagator::coordsStep
//   SUM    block partials of sum p, finalised in fixed order by the last block -> com[c]   (zeroMomentum, first half)
//   HALO   with A on a handle that owns all beads: also write the ring-wrap halo slabs
// The last block to finish also advances the noise draw counter when an O stage ran.
struct IntArgs {
    double *x, *p;
    const double *f;
    double* com_part; double* com; unsigned int* ticket; unsigned long long* draw;
    int N, D, Ploc, bead_begin;
    size_t S;
    double c1, c2, hdt, dt_over_m, inv_np;
    unsigned long long seed;
    unsigned ops;
    const double* noise;   // reference-compatible stream: [Ploc][N][D] gaussians of this half-step, or nullptr (Philox)
    unsigned long long* tl;
};

__global__ void __launch_bounds__(256) k_integrate(IntArgs a) {
    tl_begin(a.tl);
    __shared__ double sm[3 * 32];
    __shared__ bool is_last;
    const int Q = (a.N + 1) >> 1;
    const long long rows = (long long)a.Ploc * a.D;
    const long long total = rows * Q;
    const bool do_o = (a.ops & (OP_O_PRE | OP_O_POST)) != 0;
    const unsigned long long draw = do_o ? *a.draw : 0ull;
    double cm[3] = {0.0, 0.0, 0.0};
    if (a.ops & OP_SUBCM) {
        for (int c = 0; c < a.D; ++c) cm[c] = a.com[c] * a.inv_np;
    }
    double acc[3] = {0.0, 0.0, 0.0};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / Q;
        const int q = (int)(idx % Q);
        const int b = (int)(row / a.D), c = (int)(row % a.D);
        const int n0 = 2 * q;
        const bool two = (n0 + 1) < a.N;
        const size_t o = (size_t)row * a.N + n0;
        double p0 = a.p[o], p1 = two ? a.p[o + 1] : 0.0;
        if (a.ops & OP_SUBCM) {
            const double cmc = c == 0 ? cm[0] : (c == 1 ? cm[1] : cm[2]);
            p0 -= cmc; p1 -= cmc;
        }
        double z0 = 0.0, z1 = 0.0;
        if (do_o) {
            if (a.noise) {   // particle-major, axis-minor within a bead: the order langevin.cpp:18-26 consumes them
                z0 = a.noise[((size_t)b * a.N + n0) * a.D + c];
                z1 = two ? a.noise[((size_t)b * a.N + n0 + 1) * a.D + c] : 0.0;
            } else {
                gaussian_pair((uint32_t)q, (uint32_t)((a.bead_begin + b) * a.D + c), draw, a.seed, z0, z1);
            }
        }
        if (a.ops & OP_O_PRE) { p0 = a.c1 * p0 + a.c2 * z0; p1 = a.c1 * p1 + a.c2 * z1; }
        if (a.ops & (OP_B | OP_B_PHYS)) {
            p0 += a.hdt * a.f[o];
            if (two) p1 += a.hdt * a.f[o + 1];
        }
        if (a.ops & OP_O_POST) { p0 = a.c1 * p0 + a.c2 * z0; p1 = a.c1 * p1 + a.c2 * z1; }
        a.p[o] = p0;
        if (two) a.p[o + 1] = p1;
        if (a.ops & OP_A) {
            const size_t ox = o + a.S;  // skip the leading halo slab
            double x0 = a.x[ox] + a.dt_over_m * p0;
            a.x[ox] = x0;
            double x1 = 0.0;
            if (two) { x1 = a.x[ox + 1] + a.dt_over_m * p1; a.x[ox + 1] = x1; }
            if (a.ops & OP_HALO) {
                if (b == 0) {  // first owned bead -> halo after the last
                    size_t oh = (size_t)(a.Ploc + 1) * a.S + (size_t)c * a.N + n0;
                    a.x[oh] = x0;
                    if (two) a.x[oh + 1] = x1;
                }
                if (b == a.Ploc - 1) {  // last owned bead -> halo before the first
                    size_t oh = (size_t)c * a.N + n0;
                    a.x[oh] = x0;
                    if (two) a.x[oh + 1] = x1;
                }
            }
        }
        if (a.ops & OP_SUM) {
            // (odd N: the second slot of the last pair is no particle, but SUBCM and the Langevin step have acted on it)
            const double ps = p0 + (two ? p1 : 0.0);
            acc[0] += c == 0 ? ps : 0.0;
            acc[1] += c == 1 ? ps : 0.0;
            acc[2] += c == 2 ? ps : 0.0;
        }
    }
    if (a.ops & (OP_SUM | OP_O_PRE | OP_O_POST)) {
        if (a.ops & OP_SUM) {
            block_sum<3>(acc, sm);
            if (threadIdx.x == 0) {
                for (int c = 0; c < 3; ++c) a.com_part[blockIdx.x * 4 + c] = acc[c];
            }
        }
        if (threadIdx.x == 0) {
            __threadfence();
            unsigned int t = atomicAdd(a.ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last) {
            if (a.ops & OP_SUM) {
                __threadfence();
                double tot[3] = {0.0, 0.0, 0.0};
                // fixed order: each thread takes a strided set of blocks, then a block reduction
                for (int blk = threadIdx.x; blk < (int)gridDim.x; blk += blockDim.x) {
                    for (int c = 0; c < 3; ++c) tot[c] += __ldcg(&a.com_part[blk * 4 + c]);
                }
                block_sum<3>(tot, sm);
                if (threadIdx.x == 0) {
                    for (int c = 0; c < 3; ++c) a.com[c] = tot[c];
                }
            }
            if (threadIdx.x == 0) {
                *a.ticket = 0u;
                if (do_o) *a.draw = draw + 1ull;
            }
        }
    }
    tl_end(a.tl);
}

int launch_integrate(Sim* s, unsigned ops) {
    IntArgs a;
    a.x = s->x; a.p = s->p;
    a.f = (ops & OP_B_PHYS) ? s->fp : s->f;
    a.com_part = s->com_part; a.com = s->com; a.ticket = s->tickets; a.draw = s->draw;
    a.N = s->N; a.D = s->D; a.Ploc = s->Ploc; a.bead_begin = s->b0; a.S = s->S;
    a.c1 = s->c1; a.c2 = s->c2;
    a.hdt = 0.5 * s->cfg.dt; a.dt_over_m = s->cfg.dt / s->cfg.mass;
    a.inv_np = 1.0 / ((double)s->N * (double)s->P);
    a.seed = s->cfg.seed;
    a.ops = ops;
    a.noise = nullptr;
    if (s->rm_state && (ops & (OP_O_PRE | OP_O_POST))) {
        int rc = launch_ranmars_fill(s);
        if (rc != PIMDB_OK) return rc;
        a.noise = s->rm_noise;
    }
    a.tl = tl_slot(s);
    const size_t items = (size_t)s->Ploc * s->D * ((s->N + 1) / 2);
    const int grid = grid_for(items, 256, kMaxPartials);
    k_integrate<<<grid, 256, 0, s->stream>>>(a);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

// ------------------------------------------------------------------------------------------------------
// Simulation::updateNeighboringCoordinates on a handle that owns the whole ring: halo[0] = bead P-1,
// halo[P+1] = bead 0 (src/simulation.cpp:299-347, 379-382).
__global__ void k_fill_halos(double* x, size_t S, int Ploc) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (size_t)gridDim.x * blockDim.x) {
        x[i] = x[(size_t)Ploc * S + i];
        x[(size_t)(Ploc + 1) * S + i] = x[S + i];
    }
}

int launch_fill_halos(Sim* s) {
    k_fill_halos<<<grid_for(s->S, 256), 256, 0, s->stream>>>(s->x, s->S, s->Ploc);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

// ------------------------------------------------------------------------------------------------------
// Boundary transposes: host dVec layout [bead][particle][axis] <-> device [bead][axis][particle].
__global__ void k_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, int N, int D, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long b = i / ((long long)N * D);
        int r = (int)(i % ((long long)N * D));
        int c = r / N, n = r % N;                // i indexes the SoA side (coalesced writes)
        soa[i] = aos[(b * N + n) * D + c];
    }
}
__global__ void k_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, int N, int D, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long b = i / ((long long)N * D);
        int r = (int)(i % ((long long)N * D));
        int n = r / D, c = r % D;                // i indexes the AoS side
        aos[i] = soa[(b * D + c) * N + n];
    }
}

int launch_aos_to_soa(Sim* s, double* dst_soa, bool dst_has_halo) {
    long long total = (long long)s->Ploc * s->S;
    k_aos_to_soa<<<grid_for(total, 256), 256, 0, s->stream>>>(s->stage_d, dst_soa + (dst_has_halo ? s->S : 0), s->N, s->D, total);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}
int launch_soa_to_aos(Sim* s, const double* src_soa, bool src_has_halo) {
    long long total = (long long)s->Ploc * s->S;
    k_soa_to_aos<<<grid_for(total, 256), 256, 0, s->stream>>>(src_soa + (src_has_halo ? s->S : 0), s->stage_d, s->N, s->D, total);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

}  // namespace pimdb
