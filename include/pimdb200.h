/* pimdb200.h — C ABI of the B200-native PIMD force-and-propagate path (libpimdb200.so).
 *
 * This is the drop-in boundary for the hot path of higj/pimd-b (PIMD-B++). The reference has no FFI: its
 * "plugins" are C++ subclasses (Potential / BosonicExchangeBase / Propagator / Thermostat / Observable)
 * built by factories in Simulation and called from Simulation::run. The host-side C++ wrapper classes in
 * pimd_b_b200/host/ keep those class interfaces and forward to the entry points below; INTEGRATION.md shows
 * the binding a maintainer adds to the reference. Each entry point cites the reference interface it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain C, opaque handle, int status codes, no exceptions across the boundary, no torch types;
 *   - atomic units, hbar = kB = 1, i-PI convention (include/common.h:30-42), FP64 everywhere;
 *   - host arrays are what the reference holds per MPI rank (dVec, AoS [N][NDIM], include/common.h:79-239),
 *     concatenated over the beads owned by the handle: [nbeads_local][natoms][ndim], bead-major;
 *   - device state is SoA [bead][axis][particle] with one halo slice on either side of the owned bead range;
 *   - all work is enqueued on the handle's stream; calls that return host data synchronise it;
 *   - there is NO CPU fallback: every entry point fails with PIMDB_ERR_CUDA if no sm_100-class device is usable.
 *
 * Error mapping back to the reference's exception types (src/pimdb.cpp:57-63):
 *   PIMDB_ERR_INVALID_ARGUMENT -> std::invalid_argument   (src/params.cpp validation messages)
 *   PIMDB_ERR_OVERFLOW         -> std::overflow_error     (non-finite exchange potential,
 *                                                          src/bosonic_exchange/quadratic_bosonic_exchange.cpp:92-97,119-124)
 *   PIMDB_ERR_RUNTIME          -> std::runtime_error
 */
#ifndef PIMDB200_H
#define PIMDB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIMDB_ABI_VERSION 2

enum pimdb_status {
    PIMDB_OK = 0,
    PIMDB_ERR_INVALID_ARGUMENT = 1,
    PIMDB_ERR_OVERFLOW = 2,
    PIMDB_ERR_RUNTIME = 3,
    PIMDB_ERR_CUDA = 4
};

/* [interaction_potential] name / [external_potential] name (src/params.cpp:194-247, src/simulation.cpp:612-645) */
enum pimdb_potential {
    PIMDB_POT_FREE = 0,
    PIMDB_POT_AZIZ = 1,
    PIMDB_POT_HARMONIC = 2,
    PIMDB_POT_DIPOLE = 3,
    PIMDB_POT_DOUBLE_WELL = 4, /* external only */
    PIMDB_POT_COSINE = 5       /* external only */
};

/* [simulation] propagator (src/simulation.cpp:652-660) */
enum pimdb_propagator { PIMDB_PROP_CARTESIAN = 0, PIMDB_PROP_NORMAL_MODES = 1 };

/* [simulation] thermostat (src/simulation.cpp:667-684) */
enum pimdb_thermostat {
    PIMDB_THERMO_NONE = 0,
    PIMDB_THERMO_LANGEVIN = 1,
    PIMDB_THERMO_NOSE_HOOVER = 2,
    PIMDB_THERMO_NOSE_HOOVER_NP = 3,
    PIMDB_THERMO_NOSE_HOOVER_NP_DIM = 4
};

/* noise stream of the Langevin thermostat. The reference draws from one RANMAR generator per rank, seed + rank
 * (src/simulation.cpp:58-59, libs/random_mars.cpp, src/thermostats/langevin.cpp:15-27). */
enum pimdb_rng {
    PIMDB_RNG_PHILOX = 0,   /* counter-based Philox4x32-10, parallel (DESIGN.md "Noise stream"): statistical agreement */
    PIMDB_RNG_RANMARS = 1   /* the reference's own stream, sequential per bead: trajectory-level agreement, slow     */
};

/* bosonic exchange class (src/simulation.cpp:689-700) */
enum pimdb_exchange_alg {
    PIMDB_EXCH_QUADRATIC = 0,  /* Feldman-Hirshberg, O(N^2 + PN): src/bosonic_exchange/quadratic_bosonic_exchange.cpp           */
    PIMDB_EXCH_FACTORIAL = 1   /* sum over all N! permutations, natoms <= 10: src/bosonic_exchange/factorial_bosonic_exchange.cpp */
};

/* which state array (include/simulation.h:59-60; the split forces are the two locals of
 * Simulation::updateForces, src/simulation.cpp:357-361) */
enum pimdb_array {
    PIMDB_X = 0,        /* coord   */
    PIMDB_P = 1,        /* momenta */
    PIMDB_F = 2,        /* forces = spring + physical */
    PIMDB_F_SPRING = 3, /* ring springs / exterior exchange forces */
    PIMDB_F_PHYS = 4    /* external + pair forces */
};

/* exchange tables (src/bosonic_exchange/quadratic_bosonic_exchange.cpp) */
enum pimdb_exchange_table {
    PIMDB_EXCH_V = 0,        /* V[0..N]            evaluateVBn            :73-99   */
    PIMDB_EXCH_VB = 1,       /* V_backwards[0..N]  evaluateVBackwards     :101-128 */
    PIMDB_EXCH_E = 2,        /* E_kn, N(N+1)/2 entries in the reference's serial order :34-71 */
    PIMDB_EXCH_PROB = 3      /* connection_probabilities, N*N row-major [l][u]        :142-157 */
};

/* What Params + the Simulation constructor hold (src/params.cpp:8-260, src/simulation.cpp:16-101),
 * already converted to atomic units. */
typedef struct pimdb_config {
    int natoms;              /* [system] natoms                                         */
    int nbeads;              /* [simulation] nbeads (P)                                 */
    int ndim;                /* compile-time NDIM of the reference (CMakeLists.txt:44-48): 1, 2 or 3 */
    int bosonic;             /* [simulation] bosonic (only effective when nbeads > 1, src/simulation.cpp:690) */
    int fixcom;              /* [simulation] fixcom                                     */
    int pbc;                 /* [simulation] pbc                                        */
    int propagator;          /* enum pimdb_propagator                                   */
    int thermostat;          /* enum pimdb_thermostat                                   */
    int nmthermostat;        /* [simulation] nmthermostat                               */
    int nchains;             /* [simulation] nchains (Nose-Hoover only)                 */
    int int_potential;       /* enum pimdb_potential                                    */
    int ext_potential;       /* enum pimdb_potential                                    */
    double int_omega;        /* harmonic pair spring: omega                             */
    double int_strength;     /* dipole: strength                                        */
    double ext_omega;        /* harmonic trap: omega                                    */
    double ext_strength;     /* double_well: strength                                   */
    double ext_location;     /* double_well: location                                   */
    double ext_amplitude;    /* cosine: amplitude                                       */
    double ext_phase;        /* cosine: phase                                           */
    double cutoff;           /* [interaction_potential] cutoff as parsed; the library applies free -> 0 and
                                the PBC clamp min(cutoff, L/2) itself (src/simulation.cpp:84-90)            */
    double mass, temperature, dt, gamma, size;
    unsigned long long seed; /* [simulation] seed; keys the counter-based noise stream */
    /* bead sharding (one handle per GPU): this handle owns beads [bead_begin, bead_end).
       Single GPU: 0 and nbeads. */
    int bead_begin, bead_end;
    int device;              /* CUDA device ordinal */
    int rng;                 /* enum pimdb_rng: noise stream of the Langevin thermostat (0 = default)                  */
    int exchange_alg;        /* enum pimdb_exchange_alg: the reference picks its exchange class at compile time
                                (CMakeLists.txt:50-54, -DFACTORIAL_BOSONIC_ALGORITHM)                                  */
    int reserved[2];
} pimdb_config;

/* Columns of output/simulation.out (src/observables/energy.cpp, classical.cpp, bosonic.cpp), summed over the
 * beads owned by this handle, in ATOMIC units (the host applies Units::convertToUser; `temperature` is in
 * atomic units too). With bead sharding the host sums the structs of all handles, which is what
 * ObservablesLogger::log does with MPI_Allreduce (src/observables/observable.cpp:92-116). */
typedef struct pimdb_observables {
    double kinetic, potential, ext_pot, int_pot, virial;
    double temperature, cl_kinetic, cl_spring;
    double prob_dist, prob_all;
    double nh_energy;        /* Nose-Hoover: sum over beads of getAdditionToH() (src/thermostats/nose_hoover.cpp:37-67) */
    double w_gsf, pot_gsf;   /* GSF action observable (src/observables/gsf_action.cpp:21-73), free interaction only:
                                with an interaction potential the reference reads past a one-row gradient (:36), which
                                cannot be reproduced -- both fields are NaN then. w_gsf is dimensionless. */
    double reserved[3];
} pimdb_observables;

typedef struct pimdb_sim pimdb_sim;

/* ---- lifecycle ------------------------------------------------------------------------------------- */
int pimdb_abi_version(void);

/* Replaces the Simulation constructor's allocation of coord/momenta/forces/prev_coord/next_coord and of the
 * Potential / BosonicExchange / Propagator / Thermostat objects (src/simulation.cpp:16-101, 612-700).
 * Rejects what Params rejects (bosonic + normal_modes, nmthermostat with thermostat none, ...). Forces start
 * at zero like the reference's (App. A-1 of SURVEY.md). On failure *out is NULL and pimdb_last_error(NULL)
 * holds the message. */
int pimdb_create(const pimdb_config* cfg, pimdb_sim** out);
void pimdb_destroy(pimdb_sim* sim);

/* Message of the last failing call on this handle (NULL handle: last pimdb_create failure, thread-local). */
const char* pimdb_last_error(const pimdb_sim* sim);

/* ---- state ----------------------------------------------------------------------------------------- */
/* Simulation::coord / momenta / forces (include/simulation.h:59). host: [bead_end-bead_begin][natoms][ndim]. */
int pimdb_set_state(pimdb_sim* sim, int which, const double* host);
int pimdb_get_state(pimdb_sim* sim, int which, double* host);

/* Several arrays per call: one PCIe copy per array back to back and ONE transpose kernel (upload), one transpose kernel,
 * the copies and ONE synchronisation (download). NULL = skip that array. Same layout as above; page-locked arrays that sit
 * back to back in host memory (x | p | f slices of one allocation) travel as a single copy.
 * BUFFER LIFETIME: pimdb_set_state returns when the caller's buffer may be reused, whatever its kind. pimdb_upload_state
 * does the same for pageable buffers, but with PAGE-LOCKED buffers it only enqueues the copy (the contract of
 * cudaMemcpyAsync): the buffers must stay untouched until the next synchronising call on the handle returns
 * (pimdb_download_state, pimdb_step_download, pimdb_get_state, pimdb_synchronize). That lets the host enqueue the step
 * right behind the upload instead of waiting for the PCIe copy first. What a host loop that keeps its own copy of the state
 * pays per step (bench.py's e2e figure). */
int pimdb_upload_state(pimdb_sim* sim, const double* x, const double* p);
int pimdb_download_state(pimdb_sim* sim, double* x, double* p, double* f);

/* ---- the reference's per-step calls, one by one (so its loop order can be driven call by call) ------- */
/* Simulation::updateNeighboringCoordinates (src/simulation.cpp:379-382, getPrev/NextCoords :299-347).
 * With all beads on one handle it fills the two halo slices by ring wrap; with bead sharding the host
 * exchanges the halo slices (pimdb_halo_* below) instead. */
int pimdb_update_neighbors(pimdb_sim* sim);
/* Simulation::updateForces (src/simulation.cpp:353-374): springs or exchange forces + external + pair. */
int pimdb_update_forces(pimdb_sim* sim);
/* Propagator::momentStep / coordsStep (src/propagators/velocity_verlet.cpp:24-38). */
int pimdb_moment_step(pimdb_sim* sim);
int pimdb_coords_step(pimdb_sim* sim);
/* Propagator::step — VelocityVerletPropagator (velocity_verlet.cpp:7-22) or NormalModesPropagator
 * (normal_modes_propagator.cpp:19-71) according to cfg.propagator. */
int pimdb_propagator_step(pimdb_sim* sim);
/* Thermostat::step (src/thermostats/thermostat.cpp:15-19) incl. the Cartesian / normal-mode coupling. */
int pimdb_thermostat_step(pimdb_sim* sim);
/* Simulation::zeroMomentum (src/simulation.cpp:581-603). */
int pimdb_zero_momentum(pimdb_sim* sim);

/* nsteps iterations of the body of Simulation::run (src/simulation.cpp:246-259): thermostat, [COM],
 * propagator, thermostat, [COM]; fused kernels, captured once into a CUDA graph and replayed. Asynchronous.
 * Requires all beads on the handle (bead sharding drives the phases below). */
int pimdb_step(pimdb_sim* sim, int nsteps);

/* pimdb_step(nsteps) followed by pimdb_download_state(x, p, f), with the copy of the coordinates overlapping the force
 * evaluation of the last iteration (they are final once the drift has run). What a host loop that wants the state back
 * after every step calls; page-locked `x` (otherwise the plain sequence is used). NULL p / f: not downloaded. */
int pimdb_step_download(pimdb_sim* sim, int nsteps, double* x, double* p, double* f);

/* Wait for the stream; reports deferred device-side errors (PIMDB_ERR_OVERFLOW). */
int pimdb_synchronize(pimdb_sim* sim);

/* ---- bosonic exchange (BosonicExchangeBase, include/bosonic_exchange/bosonic_exchange_base.h:16-54) -- */
/* BosonicExchange::prepare (quadratic_bosonic_exchange.cpp:30-32). */
int pimdb_exchange_prepare(pimdb_sim* sim);
/* Tables; E and PROB are materialised on demand (the step itself never stores them). n = capacity of out. */
int pimdb_exchange_get(pimdb_sim* sim, int table, double* out, size_t n);

/* ---- observables (Observable::calculate + ObservablesLogger::log) ------------------------------------ */
int pimdb_observables_calc(pimdb_sim* sim, pimdb_observables* out);

/* ---- plumbing for bead sharding (one process per GPU; the host moves halos with NCCL) ---------------- */
/* Raw CUDA stream (cudaStream_t) all work of this handle is enqueued on; pimdb_set_stream adopts a
 * caller-owned stream (e.g. torch's current stream) so collectives and events order with the kernels. */
void* pimdb_get_stream(pimdb_sim* sim);
int pimdb_set_stream(pimdb_sim* sim, void* cuda_stream);
/* Device pointers + element counts of the halo traffic (contiguous ndim*natoms doubles each):
 *   which = 0: first owned bead (send to previous rank)   1: last owned bead (send to next rank)
 *   which = 2: halo before the first owned bead (recv)    3: halo after the last owned bead (recv) */
void* pimdb_halo_ptr(pimdb_sim* sim, int which, size_t* count);
/* Device pointer of the ndim partial centre-of-mass momentum sums of this handle's beads (allreduce target,
 * padded to 4 doubles), and of the pimdb_observables partial. */
void* pimdb_com_ptr(pimdb_sim* sim);
/* Phases of one iteration when beads are sharded (the host runs the collectives between them):
 *   0: thermostat half step + local COM partial        (then allreduce pimdb_com_ptr if fixcom)
 *   1: COM removal, B, A                                (then halo exchange)
 *   2: forces, B, thermostat half step + COM partial    (then allreduce if fixcom)
 *   3: COM removal */
int pimdb_step_phase(pimdb_sim* sim, int phase);

/* ---- bead sharding over peer memory (one handle per GPU; NVLink peer stores instead of host-driven collectives) ----
 * Replaces the MPI_Sendrecv of getPrev/NextCoords (src/simulation.cpp:299-347) and the MPI_Allreduce of zeroMomentum
 * (:595) by stores into the peers' memory issued by the step's own kernels. Start-up, on every rank:
 *     pimdb_create(cfg with this rank's [bead_begin, bead_end))  ->  pimdb_peer_export(blob)
 *     gather the blobs of all ranks in rank order (any transport: torch.distributed, MPI, a pipe; a single process that
 *     drives several GPUs just concatenates them)  ->  pimdb_peer_attach(world, rank, blobs)
 * After that pimdb_step, the call-by-call entry points, pimdb_update_forces and pimdb_observables_calc work on the shard
 * (observables stay per-handle partial sums). Every rank must make the same sequence of calls; device-side waits are
 * bounded (PIMDB_PEER_TIMEOUT_MS, default 20 s) and a peer that never shows up surfaces as PIMDB_ERR_RUNTIME at the next
 * synchronising call. With fixcom and a Langevin thermostat (or none) the closing zeroMomentum of an iteration is
 * subsumed by the first one of the next (Z O Z = Z O); it is carried out when the momenta are read, so reading them is a
 * collective operation too. Cartesian propagator / thermostat coupling only. */
#define PIMDB_PEER_BLOB_BYTES 256
int pimdb_peer_export(pimdb_sim* sim, void* blob_out /* PIMDB_PEER_BLOB_BYTES */);
int pimdb_peer_attach(pimdb_sim* sim, int world, int rank, const void* blobs /* world x PIMDB_PEER_BLOB_BYTES */);
int pimdb_peer_attached(const pimdb_sim* sim);
/* Enqueue, without waiting, whatever deferred work the next read of the momenta would trigger (the lazily closed
 * zeroMomentum above). One host thread driving several shards calls it on EVERY handle before it reads any of them
 * (pimdb_get_state(P), pimdb_observables_calc): those reads block, and the deferred work is collective. Harmless no-op
 * otherwise. */
int pimdb_settle(pimdb_sim* sim);

/* Number of kernels this handle has launched (graph replays count their kernel nodes). */
unsigned long long pimdb_launch_count(const pimdb_sim* sim);
/* Average duration in ms of the pair-force kernel (what = 0) / of whole steps (1) over the launches recorded since
 * pimdb_timing_reset, measured with CUDA events on the handle's stream (bench.py's roofline). */
int pimdb_timing_enable(pimdb_sim* sim, int on);
int pimdb_timing_get(pimdb_sim* sim, int what, double* ms_avg, unsigned long long* count);
/* what = 2 of pimdb_timing_get averages over the fused integrator launches; this returns their algorithmic bytes per
 * launch (p, f, x read / written once per stage group, plus the pair partials when the closing kick
 * assembles the forces), so that achieved GB/s = bytes / time. */
int pimdb_timing_integrator_bytes(pimdb_sim* sim, double* bytes_per_launch);
/* Measured FP64 FMA throughput of the device (TFLOP/s, dependent-FMA micro-benchmark): the denominator of the
 * pair-force roofline, which MEASURED_PEAKS.json does not provide. Not part of the reference surface. */
int pimdb_bench_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* PIMDB200_H */
