// Internal declarations shared by the CUDA translation units of libpimdb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/pimdb200.h"

namespace pimdb {

constexpr int kTile = 32;            // particles per pair-force tile (= warp width)
constexpr int kMaxPartials = 1184;   // upper bound on blocks of reducing kernels (8 x 148 SMs)
constexpr int kNumSM = 148;
constexpr double kEps = 1.0e-7;      // reference include/common.h:45

// Aziz HFDHE2 constants (reference src/potentials/aziz.cpp:6-13), atomic units.
constexpr double kAzRm = 5.60738, kAzA = 0.5448504e6, kAzEps = 3.42016E-5, kAzAlpha = 13.353384,
                 kAzD = 1.241314, kAzC6 = 1.3732412, kAzC8 = 0.4253785, kAzC10 = 0.1781;

// device-side error flags (host-mapped)
enum : int { kErrOverflowFwd = 1, kErrOverflowBwd = 2, kErrSyncTimeout = 4, kErrPeerTimeout = 8 };

// ---- bead sharding over peer memory (DESIGN.md "Multi-GPU") --------------------------------------------
// Every handle owns a small MAILBOX in its own device memory that its peers write over NVLink (cudaIpc mapping, or
// plain pointers inside one process) and that only the owner reads:
//   com_in[slot][rank][word]  centre-of-mass momentum sums of every rank, as self-validating 8-byte words
//                             {32 data bits, 32-bit sequence number} (an 8-byte store is single-copy atomic, so no
//                             fence / flag round trip sits between the sum and its consumers); slot = sequence & 1
//   halo_flag[2]              number of halo slices received so far from the previous / next rank (the slices land
//                             directly in the halo slabs of x; flag written after a system-scope fence)
//   credit[2]                 number of force evaluations the previous / next rank has completed, i.e. how many of
//                             the slices this rank sent have been consumed (a slice is only overwritten after that)
constexpr int kMaxPeers = 8;          // GPUs of one box
constexpr int kComWords = 8;          // 4 doubles as 8 half-words
struct PeerMailbox {
    unsigned long long com_in[2][kMaxPeers][kComWords];
    unsigned int halo_flag[2];
    unsigned int credit[2];
    unsigned int gather_flag[kMaxPeers];   // all-gather of bead slabs (normal-mode paths): number of gathers rank r has delivered
    unsigned int pad[52];
};
struct PeerDev {                      // passed by value to the kernels that talk to peers
    int world, rank;
    PeerMailbox* mine;                // local mailbox
    PeerMailbox* box[kMaxPeers];      // every rank's mailbox (box[rank] == mine)
    double* halo_to_prev;             // previous rank's trailing halo slab (receives this rank's first bead)
    double* halo_to_next;             // next rank's leading halo slab (receives this rank's last bead)
    // Self-validating halo slices (the step's own halo traffic, OP_HALO_EARLY / OP_HALO_FIX): every double travels as two
    // 8-byte words {32 data bits, 32-bit push number}, so the receiver polls the words themselves and neither side needs a
    // system-scope fence (3.5 us each on this platform) or a flag round trip. Inbox layout, behind the PeerMailbox in the
    // same allocation: [slot = push number & 1][side: 0 from the previous rank, 1 from the next][S doubles][2 words].
    // All-gather of whole bead slabs for the normal-mode transforms (every rank needs every bead of a column): each rank
    // stores its owned beads of x and p into EVERY rank's gather buffer [slot = gather number & 1][x | p][P][S], fences at
    // system scope and raises gather_flag[its rank] there. Behind the self-validating inbox in the mailbox allocation.
    double* gather_mine;              // local gather buffer (nullptr unless a normal-mode path is configured)
    double* gather_to[kMaxPeers];     // every rank's gather buffer (gather_to[rank] == gather_mine)
    unsigned long long* ll_mine;      // local inbox
    unsigned long long* ll_to_prev;   // previous rank's inbox (my first bead goes to its side 1)
    unsigned long long* ll_to_next;   // next rank's inbox (my last bead goes to its side 0)
    unsigned int* seq;                // local counters: [0] momentum-sum pushes made, [1] flag-protocol halo pushes made (= slices
                                      // expected from each neighbour), [2] self-validating halo pushes made, [3] "a peer timed
                                      // out" (sticky; later waits return at once), [4] all-gathers made
    int prev, next;
    unsigned long long timeout_ns;    // bound of every device-side wait
};

struct DevObs {  // partial sums produced on device; assembled into pimdb_observables on the host
    double spring_e[1];      // sum over owned classical links of 0.5 k |x_b - x_{b-1}|^2 (exterior link excluded for bosons)
    double ext_v;            // sum_b V_ext
    double ext_vir;          // sum_b -x . F_ext
    double pair_v;           // sum_b sum_{i<j} v
    double pair_vir;         // sum_b sum_{i<j} -x_i . f_ij
    double p2;               // sum p^2
    double prim_est;         // e[N] of the primitive-estimator recursion (bead 0 owner only)
    double v_n;              // V[N]
    double e_diag_sum;       // sum_m E^{[m..m]}
    double e_full;           // E^{[0..N-1]}
    double nh_energy;        // Nose-Hoover addition to the conserved quantity, owned beads
    double gsf[4];           // GSF observable: sum of V_ext over odd / even beads, sum of |grad V_ext|^2 over odd / even beads
    double pad[1];
};

// RANMAR state of one bead (reference libs/random_mars.h): u[1..97], c, the lag indices and the cached second gaussian
struct RanMarsState {
    double u[98];
    double c, spare;
    int i97, j97, have_spare, pad;
};

struct Sim {
    pimdb_config cfg{};
    int N = 0, P = 0, D = 0, Ploc = 0, b0 = 0, b1 = 0;
    int T = 0, TP = 0;                 // tiles per bead, upper-triangular tile pairs
    bool all_local = false, has_first = false, has_last = false, bosonic = false;
    bool pair_on = false;
    bool factorial = false;            // the reference's N!-permutation exchange class instead of Feldman-Hirshberg (factorial.cu)
    double* fact_work = nullptr;
    size_t S = 0;                      // slab stride = D*N
    double beta = 0, thermo_beta = 0, exch_beta = 0, omega_p = 0, kspring = 0, rc = 0, L = 0, kext = 0;
    double c1 = 1, c2 = 0;             // Langevin friction / noise coefficients
    double pair_par = 0;               // harmonic pair k or dipole strength

    int device = 0;
    int sm_count = 148;
    int prio_hi = 0, prio_lo = 0;      // stream / launch priorities (numerically lower = more urgent)
    cudaStream_t stream = nullptr, stream_x = nullptr;  // main stream, exchange side stream
    bool own_stream = true;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_copy = nullptr;
    std::vector<cudaEvent_t> ev_slice;  // completion of the slices of a pair-tile grid cut into chained launches

    // state (device). x has Ploc+2 slabs (halo, owned..., halo); the others Ploc slabs.
    double *x = nullptr, *p = nullptr, *f = nullptr, *fs = nullptr, *fp = nullptr;
    double *stage_d = nullptr;         // AoS staging [3][Ploc][N][D]
    double *stage_h = nullptr;         // pinned host staging, same size
    // pair forces
    ushort2* tile_ij = nullptr;        // TP entries (I,J), I<=J
    double* pair_scratch = nullptr;    // [bead_chunk][T][T][D][32]
    int bead_chunk = 0;
    // exchange
    double *exA = nullptr, *exV = nullptr, *exVb = nullptr, *exF = nullptr;  // A[N], V[N+1], Vb[N+1], F[2][D][N]
    int4 *exC = nullptr;                                                   // Boltzmann factors [2][N][N], packed extended-range numbers
    double *exK = nullptr; int *exB = nullptr;                             // block-scaled factors [2][N*N + 512], exponents [2][N/32][N] (N <= 512)
    int* exSync = nullptr;                                                 // tiles -> recurrence hand-shake counters (exchange.cu)
    cudaStream_t stream_r = nullptr; cudaEvent_t ev_join2 = nullptr;       // recurrence stream (resident early) + its join
    double *exG = nullptr; int *exGok = nullptr;                           // diagonal-block inverses [2][N/32][32][32] + validity flags (N <= 512)
    double *exWm = nullptr; int *exWe = nullptr;                           // W/Wb [2][N+1]: mantissas, binary exponents
    long long* dbg_buf = nullptr;                                          // profiling aid (PIMDB_EXCH_DEBUG)
    double *exTab = nullptr; size_t exTabCap = 0;                          // on-demand E / prob tables
    // reductions
    double* com_part = nullptr;        // [2][kMaxPartials][4]
    double* com = nullptr;             // [4] finalized sum of momenta over owned beads (allreduce target)
    unsigned int* tickets = nullptr;   // last-block-done counters
    unsigned long long* draw = nullptr;  // device counter of thermostat half-steps (noise draw index)
    DevObs* obs_d = nullptr; DevObs* obs_h = nullptr;
    double* obs_part = nullptr;        // [kMaxPartials][8]
    int* err_h = nullptr; int* err_d = nullptr;  // mapped error flags
    // normal modes
    double *nmC = nullptr;             // [P][P] Cartesian->NM matrix rows (row k = mode k)
    double *nmFreq = nullptr;          // [P] cos/sin tables: [3][P] = cos(w dt), sin(w dt), m*w
    // Nose-Hoover chains: eta | eta_dot | eta_dot_dot, each [bead][group][nchains]
    double *nh_state = nullptr; size_t nh_len = 0;
    RanMarsState* rm_state = nullptr; double* rm_noise = nullptr;   // reference-compatible noise mode (ranmars.cu): [Ploc], [Ploc][N][D]
    unsigned long long* stamps = nullptr; int stamp_next = 0;   // phase stamps of the k_integrate launches [8][8] (PIMDB_TIMELINE=1)
    unsigned char tl_kind[32] = {};
    unsigned long long* tl = nullptr; int tl_next = 0;   // in-kernel timeline slots [32][2] (PIMDB_TIMELINE=1), next slot
    // graph
    cudaGraph_t graph = nullptr; cudaGraphExec_t graph_exec = nullptr;
    unsigned long long graph_kernels = 0;
    // pimdb_step_download: the same iteration with the coordinates leaving for the host as soon as they are final
    cudaGraph_t graph_dl = nullptr; cudaGraphExec_t graph_dl_exec = nullptr;
    unsigned long long graph_dl_kernels = 0;
    double* dl_x_host = nullptr;       // page-locked destination baked into graph_dl
    double* dl_hook = nullptr;         // set while an iteration is being enqueued with that hook
    cudaStream_t stream_c = nullptr; cudaEvent_t ev_dl_fork = nullptr, ev_dl_join = nullptr; bool dl_forked = false;
    // Langevin noise of the counter-based stream, drawn ahead of its half step while the forces are computed (integrator.cu
    // k_noise_prefetch): two slots [own beads][D][N], nz_tag[slot] = index of the draw the slot holds
    double* nz = nullptr; unsigned long long* nz_tag = nullptr;
    cudaStream_t stream_n = nullptr; cudaEvent_t ev_nz_fork = nullptr, ev_nz_join = nullptr;
    bool nz_on = false, nz_want = false;
    bool peer_shares_gpu = false;      // another shard of the ring lives on this handle's GPU (tests, PIMDB_SHARD_SAME_DEVICE)
    bool no_ticketless = false;        // PIMDB_NO_TICKETLESS=1 when the handle was created (api.cu ticketless_langevin_step)
    bool p_shift_pending = false;      // fixcom: COM shift computed but not yet subtracted from p
    // peer-memory bead sharding (pimdb_peer_attach)
    bool peer_on = false;
    PeerDev peer{};
    PeerMailbox* mailbox = nullptr;    // own mailbox (exported)
    unsigned int* peer_seq = nullptr;  // own counters
    std::vector<void*> ipc_opened;     // mappings to close
    bool z_owed = false;               // peer mode, Langevin / no thermostat: the closing zeroMomentum of the last iteration
                                       // has not been carried out (it is subsumed by the first one of the next iteration)
    bool pdl_recur = false;            // the next recurrence launch may use programmatic stream serialisation
    // momentum sums on a handle that owns every bead (no peers): a SUM stage leaves one partial per block, the SUBCM stage of a
    // later launch adds them up in the fixed order itself -- no last-block pass on the producer's critical path. Two partial
    // arrays, because one launch may consume the old sums and produce new ones: com_buf = the array holding the latest.
    int com_buf = 0;
    // one-shot modifiers of the next launch_integrate (a captured step places them statically): offset added to the draw
    // counter an O stage reads; amount a launch WITHOUT an O stage adds to the counter; O stage leaves the counter alone
    int li_draw_off = 0, li_draw_bump = 0; bool li_no_ticket = false;
    int nz_first_off = 0;              // k_noise_prefetch draws counter + nz_first_off, + 1 (see enqueue_step)
    bool pdl_next = false;             // the next k_integrate / factor-tile launch follows a kernel of ours on the same stream in
                                       // a captured step: launch it with programmatic stream serialisation (it waits for that
                                       // grid first thing; what is hidden is its launch latency)
    bool split_stale = false;          // f is current but f_spring / f_phys are not (fused closing kernel)
    unsigned long long launches = 0;
    // timing
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pair, ev_step, ev_integ;
    double integ_bytes = 0.0;          // algorithmic bytes moved by the k_integrate launches in ev_integ
    std::string err;
};

// kind of kernel a timeline slot belongs to: 1 k_integrate / k_assemble, 2 exchange factor tiles (+ prefix), 3 recurrences,
// 4 exterior forces, 5 pair tiles
inline unsigned long long* tl_slot(Sim* s, int kind) {
    if (!s->tl) return nullptr;
    const int i = s->tl_next++ % 32;
    s->tl_kind[i] = (unsigned char)kind;
    return s->tl + 2 * i;
}

// launch helpers --------------------------------------------------------------------------------------
inline int grid_for(size_t items, int block, int max_blocks = 8 * kNumSM) {
    size_t g = (items + block - 1) / block;
    if (g < 1) g = 1;
    if (g > (size_t)max_blocks) g = max_blocks;
    return (int)g;
}

#define PIMDB_CUDA_TRY(sim, expr)                                                              \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            (sim)->err = std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr;   \
            return PIMDB_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)

// kernels.cu entry points (host wrappers that enqueue on s->stream unless noted)
int launch_pair_forces(Sim* s, bool with_obs);
int launch_assemble(Sim* s);
int launch_exchange(Sim* s, cudaStream_t st);          // prep + forward/backward + exterior forces
int launch_exchange_tables(Sim* s, int table);
int launch_factorial_exchange(Sim* s, cudaStream_t st);
int launch_exchange_estimators(Sim* s);
int launch_fill_halos(Sim* s);
enum : unsigned { OP_SUBCM = 1, OP_O_PRE = 2, OP_B = 4, OP_O_POST = 8, OP_A = 16, OP_SUM = 32, OP_HALO = 64, OP_B_PHYS = 128,
                  OP_ASSEMBLE = 256,   // form f = springs + external + pair partials in the same pass (before B)
                  OP_CREDIT = 512,     // peer mode: tell the ring neighbours that their halo slices have been consumed
                  OP_ZERO_SUM = 1024,  // peer mode: publish zero momentum sums (uniform entry state of a captured step)
                  // peer mode, fixcom: the boundary beads' NEW coordinates are sent to the ring neighbours one kernel early, by the
                  // kernel that takes the momentum sums, without the (not yet known) centre-of-mass shift
                  //     x~ = x + dt/m (p + dt/2 f),    x_new = x~ - dt/m com/(N P);
                  // the next kernel (SUBCM|B|A), which waits for the sums anyway, subtracts the uniform shift from the two halo
                  // slabs it received (OP_HALO_FIX). Halo slices then travel while the sums do, and neither the hand-shake nor
                  // the remote stores sit on the kernel that the force evaluation waits for.
                  OP_HALO_EARLY = 2048, OP_HALO_FIX = 4096
};
int launch_integrate(Sim* s, unsigned ops);
int launch_noise_prefetch(Sim* s, cudaStream_t st, int ndraws, int first_off);
int launch_peer_push_halos(Sim* s);
int launch_peer_allgather(Sim* s, bool with_x, bool kick);   // owned beads of (x and) p -> every rank's gather buffer
int launch_nm_propagate(Sim* s);
int launch_nm_thermostat(Sim* s);
int launch_nm_momenta(Sim* s, bool forward);   // p <-> normal-mode momenta, in place
int launch_obs_elementwise(Sim* s);
int ranmars_create(Sim* s);
int launch_ranmars_fill(Sim* s);   // the next thermostat half-step's gaussians, all owned beads
int launch_nose_hoover(Sim* s);
int launch_nose_hoover_energy(Sim* s, double* out_dev);
int launch_aos_to_soa(Sim* s, int n, double* const* dst_soa, const bool* has_halo, bool fill_ring = false, bool zero_com = false);          // up to 3 arrays per launch
int launch_soa_to_aos(Sim* s, int n, const double* const* src_soa, const bool* has_halo, cudaStream_t st = nullptr);
int launch_peer_wait_halos(Sim* s);

}  // namespace pimdb
