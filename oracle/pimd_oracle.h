/* TEST INFRASTRUCTURE ONLY — never linked into, imported by, or called from the product path.
 *
 * pimd_oracle: a plain-C, single-threaded CPU restatement of the reference's (higj/pimd-b) per-step
 * force-and-propagate hot path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * use it, and only as the checker. Every function names the reference file:line it restates
 * (paths relative to /root/reference).
 *
 * Parity status: PINNED. tests/test_oracle_golden.py checks this restatement against
 *   (1) frames of the reference's own golden cases (tests/cases/<case>/position_b.xyz + force_b.dat +
 *       velocity_b.dat, copied as small fixtures into tests/golden/ by tests/golden/make_fixtures.py), and
 *   (2) raw-double outputs of the unmodified reference compiled in this container (oracle/_ref/ref_probe_ndim*,
 *       recipe: oracle/Makefile) for the paths no reference test covers: Aziz / dipole / harmonic-pair
 *       forces, minimum image, cutoff, NDIM=2, fixcom, normal modes, the observables columns.
 *
 * Conventions (same as the reference): atomic units, hbar = kB = 1, i-PI convention
 * (include/common.h:30-42): thermo_beta = beta/P, omega_P = P/(beta*hbar), k = m*omega_P^2.
 * The reference holds one bead ("time slice") per MPI rank as an AoS dVec [N][NDIM]
 * (include/common.h:79-239, include/simulation.h:59-60); here all P slabs sit in one array
 * [P][N][NDIM], bead-major, which is also the host layout of the C ABI in include/pimdb200.h.
 */
#ifndef PIMD_ORACLE_H
#define PIMD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_POT_FREE = 0, ORC_POT_AZIZ = 1, ORC_POT_HARMONIC = 2, ORC_POT_DIPOLE = 3, ORC_POT_DOUBLE_WELL = 4,
       ORC_POT_COSINE = 5 };
enum { ORC_PROP_CARTESIAN = 0, ORC_PROP_NORMAL_MODES = 1 };
enum { ORC_THERMO_NONE = 0, ORC_THERMO_LANGEVIN = 1, ORC_THERMO_NOSE_HOOVER = 2, ORC_THERMO_NOSE_HOOVER_NP = 3,
       ORC_THERMO_NOSE_HOOVER_NP_DIM = 4 };

typedef struct {
    int natoms, nbeads, ndim;
    int bosonic, fixcom, pbc;
    int propagator, thermostat, nmthermostat;
    int int_pot, ext_pot;        /* ORC_POT_* (ext: FREE, HARMONIC, DOUBLE_WELL or COSINE) */
    double int_omega;            /* harmonic pair: omega (energy units == a.u. frequency) */
    double int_strength;         /* dipole strength */
    double ext_omega;            /* harmonic trap omega */
    double cutoff;               /* [interaction_potential] cutoff as parsed (a.u.); <0 all pairs, 0 off */
    double mass, temperature, dt, gamma, size;
    unsigned int seed;
    int nchains;                 /* Nose-Hoover chain length ([simulation] nchains, default 4) */
    double ext_strength, ext_location;   /* double_well: strength (energy), location (length), src/params.cpp:238-242 */
    double ext_amplitude, ext_phase;     /* cosine: amplitude (energy), phase; wavelength = box size (src/simulation.cpp:634-638) */
    int factorial;               /* the reference built with -DFACTORIAL_BOSONIC_ALGORITHM: the sum over all N! permutations
                                    (src/bosonic_exchange/factorial_bosonic_exchange.cpp) instead of Feldman-Hirshberg */
} orc_config;

typedef struct orc_sim orc_sim;

/* scalar columns of output/simulation.out, already summed over beads (src/observables/observable.cpp:92-116)
 * and still in atomic units (the reference converts to the user's unit when storing). */
typedef struct {
    double kinetic, potential, ext_pot, int_pot, virial;   /* energy observable  */
    double temperature, cl_kinetic, cl_spring;             /* classical          */
    double prob_dist, prob_all;                            /* bosonic            */
    double nh_energy;                                      /* classical, Nose-Hoover runs only */
    double w_gsf, pot_gsf;                                 /* gsf (free interaction only; NaN otherwise) */
} orc_observables;

orc_sim* orc_create(const orc_config* cfg);
void orc_destroy(orc_sim* s);

/* derived constants, for tests */
double orc_beta(const orc_sim* s);
double orc_spring_constant(const orc_sim* s);
double orc_cutoff_effective(const orc_sim* s);

/* state access: which = 'x','p','f','s' (spring part of f), 'e' (physical part of f); host AoS [P][N][NDIM] */
void orc_set(orc_sim* s, char which, const double* src);
void orc_get(const orc_sim* s, char which, double* dst);

/* Simulation::updateForces for every bead (src/simulation.cpp:353-374) */
void orc_update_forces(orc_sim* s);

/* one iteration of the body of Simulation::run (src/simulation.cpp:246-259) */
void orc_run_iteration(orc_sim* s);
/* the pieces, callable one by one */
void orc_thermostat_step(orc_sim* s);
void orc_zero_momentum(orc_sim* s);
void orc_propagator_step(orc_sim* s);

/* exchange tables after the last orc_update_forces (bosonic only): V[N+1], Vb[N+1],
 * E_kn[N(N+1)/2] in the reference's serial order, prob[N*N] */
int orc_exchange_get(const orc_sim* s, char which, double* dst);  /* 'V','B','E','P'; returns count */

void orc_observables_calc(orc_sim* s, orc_observables* out);

/* RANMAR (libs/random_mars.cpp) exposed for its own known-answer test */
typedef struct orc_ranmars orc_ranmars;
orc_ranmars* orc_ranmars_new(int seed);
double orc_ranmars_uniform(orc_ranmars* r);
double orc_ranmars_gaussian(orc_ranmars* r);
void orc_ranmars_free(orc_ranmars* r);

/* stand-alone pair potential / gradient-magnitude helpers (for potential unit tests):
 * returns V(r) and writes dV/dr divided by r (so that grad = out * r_vec) */
double orc_pair_potential(int pot, double r, double omega_or_strength, double mass, double* dVdr_over_r);

/* The exchange algorithm (cycle energies, V, V_backwards, connection probabilities, exterior forces, primitive
 * estimator e[N]) with every intermediate in long double: the yardstick for the double-precision algorithm's own rounding
 * noise at large N (see pimd_oracle.c). Slices are host AoS [N][D]; any output pointer may be NULL. Returns 0. */
int orc_exchange_ld(int N, int D, int pbc, double size, double kspring, double beta_exch, const double* x1, const double* xP,
                    const double* x2, const double* xPm1, double* V, double* Vb, double* f_first, double* f_last,
                    double* prim, int l_stride /* forces for particles 0, l_stride, 2 l_stride, ... only */);

#ifdef __cplusplus
}
#endif
#endif
