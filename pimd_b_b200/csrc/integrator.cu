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
    // bead shard over peer memory: the halo slabs are written by the ring neighbours; wait (bounded) for both slices
    const unsigned int* halo_flag; const unsigned int* halo_seq; unsigned long long timeout_ns; int* err;
};

template <int D>
__global__ void __launch_bounds__(256) k_assemble(AssembleArgs a) {
    tl_begin(a.tl);
    peer_wait_halos(a.halo_flag, a.halo_seq, a.timeout_ns, a.err);
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
    a.tl = tl_slot(s, 1);
    a.halo_flag = s->peer_on ? s->peer.mine->halo_flag : nullptr;   // (address arithmetic only: the mailbox is device memory)
    a.halo_seq = s->peer_on ? s->peer.seq + 1 : nullptr;
    a.timeout_ns = s->peer.timeout_ns; a.err = s->err_d;
    if (bead_lo + nb >= s->Ploc) s->split_stale = false;
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
// Fused integrator. One thread owns the particle pair (2q, 2q+1) of a row (owned bead b, axis c) -- one 16-byte
// access per array when N is even (VEC) -- and the stages run in the fixed order
// ASSEMBLE -> SUBCM -> O_PRE -> B -> O_POST -> A, selected by `ops`:
//   ASSEMBLE  f = springs (or exterior exchange force) + external force + pair partials, the arithmetic of k_assemble
//             in the same order (bit-identical); only f is written (f_spring / f_phys are rebuilt on request)
//   SUBCM     p -= com/(N P)                       zeroMomentum, second half (subtract)
//   O_*       p  = c1 p + c2 xi                    LangevinThermostat::momentaUpdate (Cartesian coupling)
//   B         p += dt/2 f   (B_PHYS: f_phys only)  Propagator::momentStep
//   A         x += dt p / m                        Propagator::coordsStep
//   SUM       block partials of sum p, finalised in fixed order by the last block -> com[c]   (zeroMomentum, first half)
//   HALO      with A: also write the halo slabs -- by ring wrap on a handle that owns all beads, into the ring
//             neighbours' slabs over peer memory on a bead shard
// The last block to finish also advances the noise draw counter when an O stage ran.
//
// Bead shards over peer memory (a.peer_on; internal.cuh PeerMailbox, DESIGN.md "Multi-GPU"):
//   SUM   the last block publishes this rank's sums to EVERY rank's mailbox as self-validating 8-byte words
//   SUBCM every block waits (bounded) for the words of all ranks and adds them in rank order: all ranks subtract the
//         bit-identical shift, and no host call or collective sits between the two kernels
//   HALO  every block first tells the neighbours "I am done reading the slices you sent" (reaching this kernel proves
//         it: everything that read them precedes it in stream order) and waits for the same word from them; then the
//         boundary beads are stored straight into the neighbours' halo slabs; the last block fences at system scope
//         and raises the neighbours' halo flags.
struct IntArgs {
    double *x, *p;
    const double* f;
    double* fw;            // ASSEMBLE: where the assembled forces go (= f)
    double* com_part; double* com; unsigned int* ticket; unsigned long long* draw;
    int N, D, Ploc, bead_begin;
    size_t S;
    double c1, c2, hdt, dt_over_m, inv_np;
    unsigned long long seed;
    unsigned ops;
    const double* noise;   // reference-compatible stream: [Ploc][N][D] gaussians of this half-step, or nullptr (Philox)
    // com_mode 1 (the handle owns every bead, no peers): SUM leaves block partials in com_out, SUBCM adds up the com_nblk
    // partials of com_in itself. use_ticket: the launch needs its last block (peer publication, totals for the host-driven
    // shards, advancing the draw counter). draw_off / draw_bump: see Sim::li_draw_off.
    int com_mode, com_nblk, use_ticket, draw_off, draw_bump, draw_inc;
    const double* com_in; double* com_out;
    const double* nz; const unsigned long long* nz_tag; size_t nz_slot;   // prefetched draws of the counter-based stream (k_noise_prefetch) or nullptr
    unsigned long long* tl;
    // ASSEMBLE
    const double* scratch; const double* exF;
    int T, first_local, last_local, pbc, ext_pot;
    double k, kext, L, invL, ext_a, ext_b, mass;
    // peers
    int peer_on;
    PeerDev peer;
    int* err;
    unsigned long long* stamps;   // profiling aid (PIMDB_TIMELINE): %globaltimer at the phases of this launch, or nullptr
};
#define PIMDB_STAMP(i) do { if (a.stamps && tid == 0 && (blockIdx.x == 0 || (i) >= 6)) a.stamps[i] = gtimer_ns(); } while (0)

template <bool VEC>
__device__ __forceinline__ double2 ld_pair(const double* ptr, size_t o, bool two) {
    if (VEC) return *reinterpret_cast<const double2*>(ptr + o);
    double2 r;
    r.x = ptr[o];
    r.y = two ? ptr[o + 1] : 0.0;
    return r;
}
template <bool VEC>
__device__ __forceinline__ void st_pair(double* ptr, size_t o, double2 v, bool two) {
    if (VEC) { *reinterpret_cast<double2*>(ptr + o) = v; return; }
    ptr[o] = v.x;
    if (two) ptr[o + 1] = v.y;
}

template <bool VEC>
__global__ void __launch_bounds__(256) k_integrate(IntArgs a) {
    grid_dependency_wait();      // (launched early behind the previous kernel of a captured step: wait for it, then let the next in)
    grid_launch_dependents();
    tl_begin(a.tl);
    __shared__ double sm[3 * 32];
    __shared__ double sh_cm[4];
    __shared__ unsigned long long sh_draw;
    __shared__ int sh_nz;
    __shared__ unsigned sh_seq[3];
    __shared__ unsigned sh_words[kMaxPeers * kComWords];
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const int Q = (a.N + 1) >> 1;
    const long long rows = (long long)a.Ploc * a.D;
    const long long total = rows * Q;
    const bool do_o = (a.ops & (OP_O_PRE | OP_O_POST)) != 0;
    const bool peer = a.peer_on != 0;
    // ---- prologue: counters (read by ONE thread, before the barrier: nothing below can overtake the last block's
    // update of them), centre-of-mass shift, hand-shake with the ring neighbours
    PIMDB_STAMP(0);
    if (tid == 0) {
        sh_draw = do_o ? *a.draw + (unsigned long long)(long long)a.draw_off : 0ull;
        if (a.draw_bump && blockIdx.x == 0) *a.draw += (unsigned long long)a.draw_bump;   // (a launch without an O stage: nobody reads it here)
        sh_nz = (do_o && a.nz && a.nz_tag[sh_draw & 1ull] == sh_draw) ? 1 : 0;   // this draw was made ahead: load it instead
        if (peer) { sh_seq[0] = a.peer.seq[0]; sh_seq[1] = a.peer.seq[1]; sh_seq[2] = a.peer.seq[2]; }
    }
    if (tid < 4) sh_cm[tid] = 0.0;
    __syncthreads();
    PIMDB_STAMP(1);
    if (a.ops & OP_SUBCM) {
        if (peer) {
            const unsigned seq = sh_seq[0];
            if (tid < a.peer.world * kComWords)
                sh_words[tid] = wait_sys_word(&a.peer.mine->com_in[seq & 1u][tid / kComWords][tid % kComWords], seq,
                                              a.peer.timeout_ns, a.err, kErrPeerTimeout, a.peer.seq + 3);
            __syncthreads();
            if (tid < 3) {
                double t = 0.0;
                for (int r = 0; r < a.peer.world; ++r)   // rank order: every rank forms the bit-identical total
                    t += __hiloint2double((int)sh_words[r * kComWords + 2 * tid + 1], (int)sh_words[r * kComWords + 2 * tid]);
                sh_cm[tid] = t * a.inv_np;
            }
        } else if (a.com_mode == 1) {
            // the producer's blocks left one partial each: the order of the additions is the last-block pass's (a strided set per
            // thread, then the block tree), so every block forms the same total, bit for bit
            double tot[3] = {0.0, 0.0, 0.0};
            for (int blk = tid; blk < a.com_nblk; blk += blockDim.x) {
                for (int c = 0; c < 3; ++c) tot[c] += __ldcg(&a.com_in[blk * 4 + c]);
            }
            block_sum<3>(tot, sm);
            if (tid == 0) {
                for (int c = 0; c < 3; ++c) {
                    sh_cm[c] = c < a.D ? tot[c] * a.inv_np : 0.0;
                    if (blockIdx.x == 0) a.com[c] = tot[c];
                }
            }
            __syncthreads();
        } else if (tid < 3) {
            sh_cm[tid] = tid < a.D ? a.com[tid] * a.inv_np : 0.0;
        }
    }
    PIMDB_STAMP(2);
    if (peer && (a.ops & OP_ASSEMBLE) && tid < 2)     // the springs of the boundary beads read the neighbours' slices
        wait_sys_u32_ge(&a.peer.mine->halo_flag[tid], sh_seq[1], a.peer.timeout_ns, a.err, kErrPeerTimeout, a.peer.seq + 3);
    const bool push_halo = peer && (a.ops & OP_HALO);
    if (push_halo && tid < 2) {
        const unsigned k = sh_seq[1] + 1u;                         // index of the halo push this kernel makes
        // I am the previous rank of `next` (its credit[0]) and the next rank of `prev` (its credit[1])
        st_sys_u32(tid == 0 ? &a.peer.box[a.peer.next]->credit[0] : &a.peer.box[a.peer.prev]->credit[1], k);
        wait_sys_u32_ge(&a.peer.mine->credit[tid], k, a.peer.timeout_ns, a.err, kErrPeerTimeout, a.peer.seq + 3);
    }
    __syncthreads();
    PIMDB_STAMP(3);
    const unsigned long long draw = sh_draw;
    const double* nzp = sh_nz ? a.nz + (size_t)(draw & 1ull) * a.nz_slot : nullptr;
    const double cm[3] = {sh_cm[0], sh_cm[1], sh_cm[2]};
    double acc[3] = {0.0, 0.0, 0.0};
    bool stored_remote = false;
    for (long long idx = (long long)blockIdx.x * blockDim.x + tid; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / Q;
        const int q = (int)(idx % Q);
        const int b = (int)(row / a.D), c = (int)(row % a.D);
        const int n0 = 2 * q;
        const bool two = VEC || (n0 + 1) < a.N;
        const size_t o = (size_t)row * a.N + n0;
        const size_t ox = o + a.S;  // skip the leading halo slab
        double2 pv = ld_pair<VEC>(a.p, o, two);
        double2 fv = make_double2(0.0, 0.0);
        if (a.ops & OP_ASSEMBLE) {
            const double2 xv = ld_pair<VEC>(a.x, ox, two);
            double2 phys = make_double2(0.0, 0.0), spring;
            if (a.ext_pot == PIMDB_POT_HARMONIC) {
                phys.x = -(a.kext * xv.x); phys.y = -(a.kext * xv.y);
            } else if (a.ext_pot == PIMDB_POT_DOUBLE_WELL) {   // grad = 4 m lambda (|x|^2 - a^2) x  (double_well.cpp:22-40)
                double r2x = 0.0, r2y = 0.0;
                for (int cc = 0; cc < a.D; ++cc) {
                    const double2 xo = ld_pair<VEC>(a.x, (size_t)(b + 1) * a.S + (size_t)cc * a.N + n0, two);
                    r2x += xo.x * xo.x; r2y += xo.y * xo.y;
                }
                phys.x = -((4.0 * a.mass * a.ext_a * (r2x - a.ext_b * a.ext_b)) * xv.x);
                phys.y = -((4.0 * a.mass * a.ext_a * (r2y - a.ext_b * a.ext_b)) * xv.y);
            } else if (a.ext_pot == PIMDB_POT_COSINE) {        // F = A k sin(k x + phase), k = 2 pi / L  (cosine.cpp:9-35)
                const double kk = 2.0 * M_PI / a.L;
                phys.x = a.ext_a * kk * sin(kk * xv.x + a.ext_b);
                phys.y = a.ext_a * kk * sin(kk * xv.y + a.ext_b);
            }
            if (a.scratch) {   // pair partials in the fixed order m = 0..T-1, eight loads in flight
                const int K = n0 / kTile, lane = n0 % kTile;
                const double* sp = a.scratch + (((size_t)b * a.T + K) * a.T * a.D + c) * kTile + lane;
                const size_t stride = (size_t)a.D * kTile;
                int m = 0;
                for (; m + 8 <= a.T; m += 8) {
                    double2 v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = ld_pair<VEC>(sp, (size_t)(m + k) * stride, two);
#pragma unroll
                    for (int k = 0; k < 8; ++k) { phys.x += v[k].x; phys.y += v[k].y; }
                }
                for (; m < a.T; ++m) {
                    const double2 v = ld_pair<VEC>(sp, (size_t)m * stride, two);
                    phys.x += v.x; phys.y += v.y;
                }
            }
            if (b == a.first_local || b == a.last_local) {
                spring = ld_pair<VEC>(a.exF, (size_t)((b == a.first_local ? 0 : 1) * a.D + c) * a.N + n0, two);
            } else {
                const double2 xp = ld_pair<VEC>(a.x, ox - a.S, two), xn = ld_pair<VEC>(a.x, ox + a.S, two);
                double dp0 = xp.x - xv.x, dn0 = xn.x - xv.x, dp1 = xp.y - xv.y, dn1 = xn.y - xv.y;
                if (a.pbc) {
                    dp0 = min_image(dp0, a.L, a.invL); dn0 = min_image(dn0, a.L, a.invL);
                    dp1 = min_image(dp1, a.L, a.invL); dn1 = min_image(dn1, a.L, a.invL);
                }
                spring.x = a.k * (dp0 + dn0);
                spring.y = a.k * (dp1 + dn1);
            }
            fv.x = spring.x + phys.x;
            fv.y = spring.y + phys.y;
            st_pair<VEC>(a.fw, o, fv, two);
        } else if (a.ops & (OP_B | OP_B_PHYS)) {
            fv = ld_pair<VEC>(a.f, o, two);
        }
        if (a.ops & OP_SUBCM) {
            const double cmc = c == 0 ? cm[0] : (c == 1 ? cm[1] : cm[2]);
            pv.x -= cmc; pv.y -= cmc;
        }
        double z0 = 0.0, z1 = 0.0;
        if (do_o) {
            if (a.noise) {   // particle-major, axis-minor within a bead: the order langevin.cpp:18-26 consumes them
                z0 = a.noise[((size_t)b * a.N + n0) * a.D + c];
                z1 = two ? a.noise[((size_t)b * a.N + n0 + 1) * a.D + c] : 0.0;
            } else if (nzp) {
                const double2 zz = ld_pair<VEC>(nzp, o, two);
                z0 = zz.x; z1 = zz.y;
            } else {
                gaussian_pair((uint32_t)q, (uint32_t)((a.bead_begin + b) * a.D + c), draw, a.seed, z0, z1);
            }
        }
        if (a.ops & OP_O_PRE) { pv.x = a.c1 * pv.x + a.c2 * z0; pv.y = a.c1 * pv.y + a.c2 * z1; }
        if (a.ops & (OP_B | OP_B_PHYS)) {
            pv.x += a.hdt * fv.x;
            if (two) pv.y += a.hdt * fv.y;
        }
        if (a.ops & OP_O_POST) { pv.x = a.c1 * pv.x + a.c2 * z0; pv.y = a.c1 * pv.y + a.c2 * z1; }
        if (a.ops & (OP_SUBCM | OP_O_PRE | OP_O_POST | OP_B | OP_B_PHYS)) st_pair<VEC>(a.p, o, pv, two);
        if ((a.ops & OP_HALO_EARLY) && (b == 0 || b == a.Ploc - 1)) {
            // x~ = x + dt/m (p + dt/2 f): where this bead will be after the coming kick and drift, up to the uniform shift
            const double2 fb = ld_pair<VEC>(a.f, o, two);
            double2 xt = ld_pair<VEC>(a.x, ox, two);
            xt.x += a.dt_over_m * (pv.x + a.hdt * fb.x);
            xt.y += a.dt_over_m * (pv.y + a.hdt * fb.y);
            const size_t within = (size_t)c * a.N + n0;
            const unsigned k = sh_seq[2] + 1u;                      // number of this self-validating push
            const unsigned long long tag = (unsigned long long)k << 32;
            const size_t slot = (size_t)(k & 1u) * 2;
            if (b == 0) {                  // my first bead is the previous rank's trailing halo: its inbox, side 1
                unsigned long long* d = a.peer.ll_to_prev + ((slot + 1) * a.S + within) * 2;
                st_sys_u64(d, tag | (unsigned)__double2loint(xt.x)); st_sys_u64(d + 1, tag | (unsigned)__double2hiint(xt.x));
                if (two) { st_sys_u64(d + 2, tag | (unsigned)__double2loint(xt.y)); st_sys_u64(d + 3, tag | (unsigned)__double2hiint(xt.y)); }
            }
            if (b == a.Ploc - 1) {         // my last bead is the next rank's leading halo: its inbox, side 0
                unsigned long long* d = a.peer.ll_to_next + (slot * a.S + within) * 2;
                st_sys_u64(d, tag | (unsigned)__double2loint(xt.x)); st_sys_u64(d + 1, tag | (unsigned)__double2hiint(xt.x));
                if (two) { st_sys_u64(d + 2, tag | (unsigned)__double2loint(xt.y)); st_sys_u64(d + 3, tag | (unsigned)__double2hiint(xt.y)); }
            }
        }
        if ((a.ops & OP_HALO_FIX) && (b == 0 || b == a.Ploc - 1)) {
            // unpack what the neighbours sent one kernel ago (each word carries the push number: poll the words
            // themselves), take the now-known uniform shift off, and store the slice where every reader expects it
            const double shift = a.dt_over_m * (c == 0 ? cm[0] : (c == 1 ? cm[1] : cm[2]));
            const size_t within = (size_t)c * a.N + n0;
            const unsigned k = sh_seq[2];
            const size_t slot = (size_t)(k & 1u) * 2;
            auto recv = [&](const unsigned long long* w) {
                const unsigned lo = wait_sys_word(w, k, a.peer.timeout_ns, a.err, kErrPeerTimeout, a.peer.seq + 3);
                const unsigned hi = wait_sys_word(w + 1, k, a.peer.timeout_ns, a.err, kErrPeerTimeout, a.peer.seq + 3);
                return __hiloint2double((int)hi, (int)lo) - shift;
            };
            if (b == a.Ploc - 1) {   // trailing halo slab: the next rank's first bead (inbox side 1)
                const unsigned long long* w = a.peer.ll_mine + ((slot + 1) * a.S + within) * 2;
                double2 h;
                h.x = recv(w); h.y = two ? recv(w + 2) : 0.0;
                st_pair<VEC>(a.x, (size_t)(a.Ploc + 1) * a.S + within, h, two);
            }
            if (b == 0) {            // leading halo slab: the previous rank's last bead (inbox side 0)
                const unsigned long long* w = a.peer.ll_mine + (slot * a.S + within) * 2;
                double2 h;
                h.x = recv(w); h.y = two ? recv(w + 2) : 0.0;
                st_pair<VEC>(a.x, within, h, two);
            }
        }
        if (a.ops & OP_A) {
            double2 xv = ld_pair<VEC>(a.x, ox, two);
            xv.x += a.dt_over_m * pv.x;
            if (two) xv.y += a.dt_over_m * pv.y;
            st_pair<VEC>(a.x, ox, xv, two);
            if (a.ops & OP_HALO) {
                const size_t within = (size_t)c * a.N + n0;
                if (b == 0) {  // first owned bead -> the trailing halo slab of the previous rank (own slab on a full ring)
                    if (peer) { st_pair<VEC>(a.peer.halo_to_prev, within, xv, two); stored_remote = true; }
                    else st_pair<VEC>(a.x, (size_t)(a.Ploc + 1) * a.S + within, xv, two);
                }
                if (b == a.Ploc - 1) {  // last owned bead -> the leading halo slab of the next rank
                    if (peer) { st_pair<VEC>(a.peer.halo_to_next, within, xv, two); stored_remote = true; }
                    else st_pair<VEC>(a.x, within, xv, two);
                }
            }
        }
        if (a.ops & OP_SUM) {
            // (odd N: the second slot of the last pair is no particle, but SUBCM and the Langevin step have acted on it)
            const double ps = pv.x + (two ? pv.y : 0.0);
            acc[0] += c == 0 ? ps : 0.0;
            acc[1] += c == 1 ? ps : 0.0;
            acc[2] += c == 2 ? ps : 0.0;
        }
    }
    PIMDB_STAMP(4);
    if (!a.use_ticket) {
        if (a.ops & OP_SUM) {
            block_sum<3>(acc, sm);
            if (tid == 0) {
                for (int c = 0; c < 3; ++c) a.com_out[blockIdx.x * 4 + c] = acc[c];
            }
        }
    } else {
        if (a.ops & OP_SUM) {
            block_sum<3>(acc, sm);
            if (tid == 0) {
                for (int c = 0; c < 3; ++c) a.com_out[blockIdx.x * 4 + c] = acc[c];
            }
        }
        // (remote stores of this block's threads -> barrier -> ONE system-scope fence by thread 0 -> ticket: fences are
        // cumulative, so the slices are ordered before whatever the last block publishes after it has seen every ticket)
        const bool any_remote = __syncthreads_or(stored_remote);
        if (tid == 0) {
            if (any_remote) __threadfence_system();
            else __threadfence();
            unsigned int t = atomicAdd(a.ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        PIMDB_STAMP(5);
        if (is_last) {
            PIMDB_STAMP(6);
            double tot[3] = {0.0, 0.0, 0.0};
            if ((a.ops & OP_SUM) && a.com_mode == 0) {
                __threadfence();
                // fixed order: each thread takes a strided set of blocks, then a block reduction
                for (int blk = tid; blk < (int)gridDim.x; blk += blockDim.x) {
                    for (int c = 0; c < 3; ++c) tot[c] += __ldcg(&a.com_out[blk * 4 + c]);
                }
                block_sum<3>(tot, sm);
                if (tid == 0) {
                    for (int c = 0; c < 3; ++c) { a.com[c] = tot[c]; sh_cm[c] = tot[c]; }
                }
                __syncthreads();
            }
            if (peer && (a.ops & (OP_SUM | OP_ZERO_SUM))) {
                // publish to every rank (myself included): word w carries half w&1 of sum w/2, tagged with the sequence number
                const unsigned seq = sh_seq[0] + 1u;
                if (tid < a.peer.world * kComWords) {
                    const int r = tid / kComWords, w = tid % kComWords;
                    const double v = (w < 6 && (a.ops & OP_SUM)) ? sh_cm[w >> 1] : 0.0;
                    const unsigned half = (w & 1) ? (unsigned)__double2hiint(v) : (unsigned)__double2loint(v);
                    st_sys_u64(&a.peer.box[r]->com_in[seq & 1u][a.peer.rank][w], ((unsigned long long)seq << 32) | half);
                }
                if (tid == 0) a.peer.seq[0] = seq;
            }
            if (peer && (a.ops & OP_HALO_EARLY) && tid == 0) a.peer.seq[2] = sh_seq[2] + 1u;
            if (push_halo && tid == 0) {
                __threadfence_system();      // every block fenced its slices before its ticket; order the flags behind them
                const unsigned k = sh_seq[1] + 1u;
                st_sys_u32(&a.peer.box[a.peer.prev]->halo_flag[1], k);   // my first bead is prev's trailing halo
                st_sys_u32(&a.peer.box[a.peer.next]->halo_flag[0], k);   // my last bead is next's leading halo
                a.peer.seq[1] = k;
            }
            if (peer && (a.ops & OP_ASSEMBLE) && tid < 2) {
                // every block has read the halo slabs for the last time before the next push: tell the neighbours now, so
                // that the hand-shake of their next push finds the word already there
                st_sys_u32(tid == 0 ? &a.peer.box[a.peer.next]->credit[0] : &a.peer.box[a.peer.prev]->credit[1], sh_seq[1] + 1u);
            }
            if (tid == 0) {
                *a.ticket = 0u;
                if (a.draw_inc) *a.draw = draw + 1ull;       // (draw_off is 0 whenever an O stage advances the counter itself)
            }
            PIMDB_STAMP(7);
        }
    }
    tl_end(a.tl);
}

// Simulation::updateNeighboringCoordinates on a bead shard (peer memory): the first / last owned bead slices go to the
// ring neighbours' halo slabs, same hand-shake as the HALO stage of k_integrate. One launch per state upload.
__global__ void __launch_bounds__(256) k_peer_push_halos(const double* x, size_t S, int Ploc, PeerDev peer, unsigned int* ticket, int* err) {
    __shared__ bool is_last;
    __shared__ unsigned sh_k;
    const int tid = threadIdx.x;
    if (tid == 0) sh_k = peer.seq[1] + 1u;
    __syncthreads();
    const unsigned k = sh_k;
    if (tid < 2) {
        st_sys_u32(tid == 0 ? &peer.box[peer.next]->credit[0] : &peer.box[peer.prev]->credit[1], k);
        wait_sys_u32_ge(&peer.mine->credit[tid], k, peer.timeout_ns, err, kErrPeerTimeout, peer.seq + 3);
    }
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < S; i += (size_t)gridDim.x * blockDim.x) {
        peer.halo_to_prev[i] = x[S + i];
        peer.halo_to_next[i] = x[(size_t)Ploc * S + i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && tid == 0) {
        __threadfence_system();
        st_sys_u32(&peer.box[peer.prev]->halo_flag[1], k);
        st_sys_u32(&peer.box[peer.next]->halo_flag[0], k);
        peer.seq[1] = k;
        *ticket = 0u;
    }
}

// All-gather of the owned bead slabs for the normal-mode transforms (NormalModes::shareData, src/normal_modes.cpp:13-46:
// the reference copies every rank's row into a shared window). `kick`: the momenta leave with the propagator's first half
// kick applied, p + dt/2 f_phys (normal_modes_propagator.cpp:73-103), so that the transform kernel needs no forces of beads
// it does not own.
__global__ void __launch_bounds__(256) k_peer_allgather(const double* x, const double* p, const double* fphys, size_t S, int P,
                                                        int Ploc, int b0, double hdt, int with_x, PeerDev peer,
                                                        unsigned int* ticket, int* err) {
    __shared__ bool is_last;
    __shared__ unsigned sh_g;
    const int tid = threadIdx.x;
    if (tid == 0) sh_g = peer.seq[4] + 1u;
    __syncthreads();
    const unsigned g = sh_g;
    const size_t slot = (size_t)(g & 1u) * 2 * P * S;
    const size_t own = (size_t)Ploc * S;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < own; i += (size_t)gridDim.x * blockDim.x) {
        const double pv = p[i] + (fphys ? hdt * fphys[i] : 0.0);
        const double xv = with_x ? x[S + i] : 0.0;          // (x carries a leading halo slab)
        const size_t at = (size_t)b0 * S + i;
        for (int r = 0; r < peer.world; ++r) {
            double* dst = peer.gather_to[r] + slot;
            if (with_x) dst[at] = xv;
            dst[(size_t)P * S + at] = pv;
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && tid < peer.world) {
        if (tid == 0) __threadfence_system();
        __syncwarp();
        st_sys_u32(&peer.box[tid]->gather_flag[peer.rank], g);
        if (tid == 0) { peer.seq[4] = g; *ticket = 0u; }
    }
}

int launch_peer_allgather(Sim* s, bool with_x, bool kick) {
    k_peer_allgather<<<grid_for((size_t)s->Ploc * s->S, 256, 2 * kNumSM), 256, 0, s->stream>>>(
        s->x, s->p, kick ? s->fp : nullptr, s->S, s->P, s->Ploc, s->b0, 0.5 * s->cfg.dt, with_x ? 1 : 0, s->peer, s->tickets + 1, s->err_d);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_peer_push_halos(Sim* s) {
    if (!s->peer_on) return PIMDB_OK;
    k_peer_push_halos<<<grid_for(s->S, 256, kNumSM), 256, 0, s->stream>>>(s->x, s->S, s->Ploc, s->peer, s->tickets + 1, s->err_d);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

// The Gaussians of the next `ndraws` Langevin half steps of the counter-based stream, written in the layout k_integrate
// reads the momenta in. They depend on (seed, draw index, global bead, axis, particle) only, so they can be drawn while the
// forces are computed instead of on the step's serial path (Philox + Box-Muller is ~600 dependent instructions per thread:
// the thermostat half step of C3 took 6 us for 0.8 MB of momenta). k_integrate uses a slot only when its tag names the draw
// it is about to make, and otherwise draws in place: the numbers are the same either way.
struct NoiseArgs {
    double* nz; unsigned long long* tag; const unsigned long long* draw;
    size_t slot;
    int N, D, Ploc, bead_begin, ndraws, first_off;
    unsigned long long seed;
};
template <bool VEC>
__global__ void __launch_bounds__(256) k_noise_prefetch(NoiseArgs a) {
    const unsigned long long d0 = *a.draw + (unsigned long long)(long long)a.first_off;
    const int Q = (a.N + 1) >> 1;
    const long long total = (long long)a.Ploc * a.D * Q;
    for (int k = 0; k < a.ndraws; ++k) {
        const unsigned long long draw = d0 + (unsigned long long)k;
        double* dst = a.nz + (size_t)(draw & 1ull) * a.slot;
        for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
            const long long row = idx / Q;
            const int q = (int)(idx % Q);
            const int b = (int)(row / a.D), c = (int)(row % a.D);
            const int n0 = 2 * q;
            const bool two = VEC || (n0 + 1) < a.N;
            double2 z;
            gaussian_pair((uint32_t)q, (uint32_t)((a.bead_begin + b) * a.D + c), draw, a.seed, z.x, z.y);
            st_pair<VEC>(dst, (size_t)row * a.N + n0, z, two);
        }
        // (stream order puts every consumer behind this whole grid, so the tag may be written at any point of it)
        if (blockIdx.x == 0 && threadIdx.x == 0) a.tag[draw & 1ull] = draw;
    }
}

int launch_noise_prefetch(Sim* s, cudaStream_t st, int ndraws, int first_off) {
    NoiseArgs a;
    a.nz = s->nz; a.tag = s->nz_tag; a.draw = s->draw;
    a.slot = (size_t)s->Ploc * s->D * s->N;
    a.N = s->N; a.D = s->D; a.Ploc = s->Ploc; a.bead_begin = s->b0; a.ndraws = ndraws; a.first_off = first_off;
    a.seed = s->cfg.seed;
    const long long total = (long long)s->Ploc * s->D * ((s->N + 1) / 2);
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, 4LL * s->sm_count));
    if ((s->N & 1) == 0) k_noise_prefetch<true><<<grid, 256, 0, st>>>(a);
    else k_noise_prefetch<false><<<grid, 256, 0, st>>>(a);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_integrate(Sim* s, unsigned ops) {
    IntArgs a;
    a.x = s->x; a.p = s->p;
    a.f = (ops & OP_B_PHYS) ? s->fp : s->f;
    a.fw = s->f;
    a.com_part = s->com_part; a.com = s->com; a.ticket = s->tickets; a.draw = s->draw;
    {
        const bool do_o = (ops & (OP_O_PRE | OP_O_POST)) != 0;
        const bool consumer = s->all_local && !s->peer_on;
        a.com_mode = consumer ? 1 : 0;
        a.com_in = s->com_part + (size_t)(consumer ? s->com_buf : 0) * 4 * kMaxPartials;
        a.com_out = s->com_part + (size_t)(consumer ? (s->com_buf ^ 1) : 0) * 4 * kMaxPartials;
        if (consumer && (ops & OP_SUM)) s->com_buf ^= 1;
        a.draw_off = do_o ? s->li_draw_off : 0;
        a.draw_bump = do_o ? 0 : s->li_draw_bump;
        a.draw_inc = (do_o && !s->li_no_ticket) ? 1 : 0;
        // who needs the last block: an O stage that advances the counter itself (on bead shards it also publishes the early
        // credit of the flag protocol then); on bead shards the momentum sums (published to the peers / left as totals for the
        // host-driven shards) and the flag of a halo push
        bool ticket = a.draw_inc != 0;
        if (!consumer) ticket = ticket || (ops & (OP_SUM | OP_ZERO_SUM)) != 0 || (s->peer_on && (ops & OP_HALO) != 0);
        a.use_ticket = ticket ? 1 : 0;
        s->li_draw_off = 0; s->li_draw_bump = 0; s->li_no_ticket = false;
    }
    a.N = s->N; a.D = s->D; a.Ploc = s->Ploc; a.bead_begin = s->b0; a.S = s->S;
    a.c1 = s->c1; a.c2 = s->c2;
    a.hdt = 0.5 * s->cfg.dt; a.dt_over_m = s->cfg.dt / s->cfg.mass;
    a.inv_np = 1.0 / ((double)s->N * (double)s->P);
    a.seed = s->cfg.seed;
    a.ops = ops;
    a.noise = nullptr;
    a.nz = s->nz_on ? s->nz : nullptr; a.nz_tag = s->nz_tag; a.nz_slot = (size_t)s->Ploc * s->D * s->N;
    if (s->rm_state && (ops & (OP_O_PRE | OP_O_POST))) {
        int rc = launch_ranmars_fill(s);
        if (rc != PIMDB_OK) return rc;
        a.noise = s->rm_noise;
    }
    a.tl = tl_slot(s, 1);
    a.scratch = s->pair_on ? s->pair_scratch : nullptr; a.exF = s->exF;
    a.T = s->T;
    a.first_local = (s->bosonic && s->has_first) ? 0 : -1;
    a.last_local = (s->bosonic && s->has_last) ? s->Ploc - 1 : -1;
    a.pbc = s->cfg.pbc; a.ext_pot = s->cfg.ext_potential;
    a.k = s->kspring; a.kext = s->kext; a.L = s->L; a.invL = 1.0 / s->L; a.mass = s->cfg.mass;
    if (s->cfg.ext_potential == PIMDB_POT_DOUBLE_WELL) { a.ext_a = s->cfg.ext_strength; a.ext_b = s->cfg.ext_location; }
    else { a.ext_a = s->cfg.ext_amplitude; a.ext_b = s->cfg.ext_phase; }
    a.peer_on = s->peer_on ? 1 : 0;
    a.peer = s->peer;
    a.err = s->err_d;
    a.stamps = s->stamps ? s->stamps + 8 * (s->stamp_next++ % 8) : nullptr;
    if (ops & OP_ASSEMBLE) s->split_stale = true;
    const size_t items = (size_t)s->Ploc * s->D * ((s->N + 1) / 2);
    const int grid = grid_for(items, 256, kMaxPartials);
    a.com_nblk = grid;      // (every k_integrate launch of a handle has this grid)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (s->timing) {
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, s->stream);
    }
    {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(grid); lc.blockDim = dim3(256); lc.dynamicSmemBytes = 0; lc.stream = s->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at;
        lc.numAttrs = s->pdl_next ? 1 : 0;
        s->pdl_next = false;
        if (s->N % 2 == 0) cudaLaunchKernelEx(&lc, k_integrate<true>, a);
        else cudaLaunchKernelEx(&lc, k_integrate<false>, a);
    }
    s->launches += 1;
    if (e0) {
        cudaEventRecord(e1, s->stream);
        s->ev_integ.emplace_back(e0, e1);
        // algorithmic bytes of this launch per degree of freedom: p read + written when a stage changes it, f read by a kick
        // (written, with x and its two neighbours and the T pair partials read, when the forces are assembled here),
        // x read + written by the drift
        double per_dof = 0.0;
        if (ops & (OP_SUBCM | OP_O_PRE | OP_O_POST | OP_B | OP_B_PHYS)) per_dof += 16.0;
        else if (ops & (OP_SUM | OP_A)) per_dof += 8.0;
        if (ops & OP_ASSEMBLE) per_dof += 8.0 + 24.0 + (s->pair_on ? 8.0 * s->T : 0.0);
        else if (ops & (OP_B | OP_B_PHYS)) per_dof += 8.0;
        if (ops & OP_A) per_dof += 16.0;
        s->integ_bytes += per_dof * (double)s->Ploc * (double)s->S;
    }
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
// Boundary transposes: host dVec layout [bead][particle][axis] <-> device [bead][axis][particle], up to three arrays
// per launch (staging buffer: the arrays one after the other).
struct TransposeArgs {
    double* soa[3];        // device arrays (already offset past a leading halo slab)
    double* aos;           // staging buffer
    int n, N, D;
    long long total;       // elements per array
    int ring_arr, Ploc;    // k_aos_to_soa: array `ring_arr` (-1: none) owns the whole ring -- its first / last bead also go to
                           // the trailing / leading halo slab (what k_fill_halos does, without the extra launch)
    double* zero; int nzero;   // k_aos_to_soa: doubles to clear on the way (the pending momentum sums when p is replaced)
};
__global__ void k_aos_to_soa(TransposeArgs a) {
    const long long ND = (long long)a.N * a.D;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.nzero; i += gridDim.x * blockDim.x) a.zero[i] = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.total * a.n; i += (long long)gridDim.x * blockDim.x) {
        const int arr = (int)(i / a.total);
        const long long j = i - arr * a.total;
        const long long b = j / ND;
        const int r = (int)(j % ND);
        const int c = r / a.N, n = r % a.N;       // j indexes the SoA side (coalesced writes)
        double* dst = arr == 0 ? a.soa[0] : (arr == 1 ? a.soa[1] : a.soa[2]);
        const double v = a.aos[arr * a.total + (b * a.N + n) * a.D + c];
        dst[j] = v;
        if (arr == a.ring_arr) {
            if (b == 0) dst[(long long)a.Ploc * ND + r] = v;             // trailing halo = bead 0
            if (b == a.Ploc - 1) dst[r - ND] = v;                         // leading halo = last bead
        }
    }
}
__global__ void k_soa_to_aos(TransposeArgs a) {
    const long long ND = (long long)a.N * a.D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.total * a.n; i += (long long)gridDim.x * blockDim.x) {
        const int arr = (int)(i / a.total);
        const long long j = i - arr * a.total;
        const long long b = j / ND;
        const int r = (int)(j % ND);
        const int n = r / a.D, c = r % a.D;       // j indexes the AoS side
        const double* src = arr == 0 ? a.soa[0] : (arr == 1 ? a.soa[1] : a.soa[2]);
        a.aos[i] = src[(b * a.D + c) * a.N + n];
    }
}

int launch_aos_to_soa(Sim* s, int n, double* const* dst_soa, const bool* has_halo, bool fill_ring, bool zero_com) {
    TransposeArgs a{};
    a.ring_arr = -1; a.Ploc = s->Ploc;
    if (zero_com) { a.zero = s->com_part; a.nzero = 4 * kMaxPartials; }
    for (int i = 0; i < n; ++i) {
        a.soa[i] = dst_soa[i] + (has_halo[i] ? s->S : 0);
        if (fill_ring && has_halo[i]) a.ring_arr = i;        // (x is the only array with halo slabs)
    }
    a.aos = s->stage_d; a.n = n; a.N = s->N; a.D = s->D; a.total = (long long)s->Ploc * s->S;
    k_aos_to_soa<<<grid_for(a.total * n, 256), 256, 0, s->stream>>>(a);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}
int launch_soa_to_aos(Sim* s, int n, const double* const* src_soa, const bool* has_halo, cudaStream_t st) {
    TransposeArgs a{};
    for (int i = 0; i < n; ++i) a.soa[i] = const_cast<double*>(src_soa[i]) + (has_halo[i] ? s->S : 0);
    a.aos = s->stage_d; a.n = n; a.N = s->N; a.D = s->D; a.total = (long long)s->Ploc * s->S;
    k_soa_to_aos<<<grid_for(a.total * n, 256), 256, 0, st ? st : s->stream>>>(a);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

// Bead shard: a one-block kernel that waits (bounded) for both halo slices; used ahead of kernels that read the halo
// slabs but are not part of the force evaluation (estimators right after an upload).
__global__ void k_peer_wait_halos(const unsigned int* halo_flag, const unsigned int* halo_seq, unsigned long long timeout_ns, int* err) {
    peer_wait_halos(halo_flag, halo_seq, timeout_ns, err);
}
int launch_peer_wait_halos(Sim* s) {
    if (!s->peer_on) return PIMDB_OK;
    k_peer_wait_halos<<<1, 32, 0, s->stream>>>(s->peer.mine->halo_flag, s->peer.seq + 1, s->peer.timeout_ns, s->err_d);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

}  // namespace pimdb
