/* TEST INFRASTRUCTURE ONLY — see pimd_oracle.h. Plain C11, single thread, scalar loops.
 * CPU restatement of the reference hot path; citations are /root/reference/<file>:<lines>. */
#include "pimd_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EPS 1.0e-7 /* include/common.h:45 */
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ RANMAR (libs/random_mars.cpp) */
struct orc_ranmars {
    double u[98];
    int i97, j97;
    double c, cd, cm;
    int have_spare;
    double spare;
};

/* libs/random_mars.cpp:62-77 — lagged-Fibonacci subtract-with-borrow step plus the arithmetic sequence c */
double orc_ranmars_uniform(orc_ranmars* r) {
    double v = r->u[r->i97] - r->u[r->j97];
    if (v < 0.0) v += 1.0;
    r->u[r->i97] = v;
    if (--r->i97 == 0) r->i97 = 97;
    if (--r->j97 == 0) r->j97 = 97;
    r->c -= r->cd;
    if (r->c < 0.0) r->c += r->cm;
    v -= r->c;
    if (v < 0.0) v += 1.0;
    return v;
}

/* libs/random_mars.cpp:10-56 — seed expansion into 97 24-bit fractions; one uniform is burnt at the end */
orc_ranmars* orc_ranmars_new(int seed) {
    if (seed <= 0 || seed > 900000000) return NULL;
    orc_ranmars* r = (orc_ranmars*)calloc(1, sizeof *r);
    int ij = (seed - 1) / 30082;
    int kl = (seed - 1) - 30082 * ij;
    int i = (ij / 177) % 177 + 2;
    int j = ij % 177 + 2;
    int k = (kl / 169) % 178 + 1;
    int l = kl % 169;
    for (int ii = 1; ii <= 97; ++ii) {
        double s = 0.0, t = 0.5;
        for (int jj = 1; jj <= 24; ++jj) {
            int m = ((i * j) % 179) * k % 179;
            i = j; j = k; k = m;
            l = (53 * l + 1) % 169;
            if ((l * m) % 64 >= 32) s += t;
            t *= 0.5;
        }
        r->u[ii] = s;
    }
    r->c = 362436.0 / 16777216.0;
    r->cd = 7654321.0 / 16777216.0;
    r->cm = 16777213.0 / 16777216.0;
    r->i97 = 97;
    r->j97 = 33;
    r->have_spare = 0;
    (void)orc_ranmars_uniform(r);
    return r;
}

/* libs/random_mars.cpp:83-103 — polar Box–Muller; the FIRST value returned is v2*fac, v1*fac is cached */
double orc_ranmars_gaussian(orc_ranmars* r) {
    if (r->have_spare) {
        r->have_spare = 0;
        return r->spare;
    }
    double v1, v2, rsq;
    do {
        v1 = 2.0 * orc_ranmars_uniform(r) - 1.0;
        v2 = 2.0 * orc_ranmars_uniform(r) - 1.0;
        rsq = v1 * v1 + v2 * v2;
    } while (rsq >= 1.0 || rsq == 0.0);
    double fac = sqrt(-2.0 * log(rsq) / rsq);
    r->spare = v1 * fac;
    r->have_spare = 1;
    return v2 * fac;
}

void orc_ranmars_free(orc_ranmars* r) { free(r); }

/* ------------------------------------------------------------------ state */
struct orc_sim {
    orc_config c;
    int N, P, D;
    double beta, thermo_beta, omega_p, k_spring; /* src/simulation.cpp:37-52 */
    double rc;                                   /* effective cutoff, src/simulation.cpp:84-90 */
    double *x, *p, *f, *f_spring, *f_phys;       /* [P][N][D] */
    /* exchange tables (src/bosonic_exchange/quadratic_bosonic_exchange.cpp:8-16) */
    double *E_kn, *V, *Vb, *prob, *tmp, *prim;
    orc_ranmars** rng;                           /* one RANMAR per bead, seed+bead (src/simulation.cpp:58-59) */
    double *nm_fwd, *nm_inv;                     /* [P][P] rows: fwd[k][j] = C_kj ; inv[j][k] as the reference stores them */
    double *nm_ext;                              /* NormalModesPropagator::ext_forces (zero before the first step) */
    double *scratch_x, *scratch_p;               /* [P][N][D] NM-space copies */
    double *scratch_slab;                        /* [N][D] external gradient of one bead (observables) */
    /* Nose-Hoover chains, one thermostat object per bead (rank): eta, eta_dot, eta_dot_dot [P][groups][nchains] */
    double *nh_eta, *nh_ed, *nh_edd;
    int nh_groups;
};

static size_t slab(const orc_sim* s) { return (size_t)s->N * s->D; }
static size_t idx(const orc_sim* s, int b, int i, int a) { return ((size_t)b * s->N + i) * s->D + a; }

/* src/common.cpp:41-43 */
static double min_image(double dx, double L) { return dx - L * floor(dx / L + 0.5); }

/* ------------------------------------------------------------------ pair / external potentials */
/* Aziz HFDHE2 constants: src/potentials/aziz.cpp:6-13 */
static const double AZ_RM = 5.60738, AZ_A = 0.5448504e6, AZ_EPS = 3.42016E-5, AZ_ALPHA = 13.353384,
                    AZ_D = 1.241314, AZ_C6 = 1.3732412, AZ_C8 = 0.4253785, AZ_C10 = 0.1781;

/* include/potentials/aziz.h:25-34 */
static double aziz_F(double x) {
    if (x >= AZ_D) return 1.0;
    double q = AZ_D / x - 1.0;
    return exp(-q * q);
}
static double aziz_dF(double x) {
    if (x >= AZ_D) return 0.0;
    double ix = 1.0 / x, q = AZ_D * ix - 1.0;
    return 2.0 * AZ_D * ix * ix * q * exp(-q * q);
}

/* V(r): src/potentials/aziz.cpp:18-54 ; dV/dr: :56-106 (returned divided by r so grad = g * r_vec) */
static double aziz_eval(double r, double* g_over_r) {
    double xs = r / AZ_RM;
    double rep = AZ_A * exp(-AZ_ALPHA * xs);
    double t1 = -AZ_A * AZ_ALPHA * exp(-AZ_ALPHA * xs);
    double v, dvdr;
    if (xs > ORC_EPS && xs < 0.01) { /* hard-core branch: repulsion only */
        v = AZ_EPS * rep;
        dvdr = t1 * (AZ_EPS / AZ_RM);
    } else {
        double ix = 1.0 / xs, ix2 = ix * ix, ix6 = ix2 * ix2 * ix2, ix7 = ix6 * ix, ix8 = ix6 * ix2,
               ix9 = ix8 * ix, ix10 = ix8 * ix2, ix11 = ix10 * ix;
        /* V uses 1/(x*x) (aziz.cpp:44) while gradV uses (1/x)^2 (aziz.cpp:81-82); same to rounding */
        double jx2 = 1.0 / (xs * xs), jx6 = jx2 * jx2 * jx2, jx8 = jx6 * jx2, jx10 = jx8 * jx2;
        v = AZ_EPS * (rep - (AZ_C6 * jx6 + AZ_C8 * jx8 + AZ_C10 * jx10) * aziz_F(xs));
        double t2 = (6.0 * AZ_C6 * ix7 + 8.0 * AZ_C8 * ix9 + 10.0 * AZ_C10 * ix11) * aziz_F(xs);
        double t3 = -(AZ_C6 * ix6 + AZ_C8 * ix8 + AZ_C10 * ix10) * aziz_dF(xs);
        dvdr = (AZ_EPS / AZ_RM) * (t1 + t2 + t3);
    }
    *g_over_r = dvdr / r;
    return v;
}

double orc_pair_potential(int pot, double r, double par, double mass, double* g_over_r) {
    double g = 0.0, v = 0.0;
    switch (pot) {
        case ORC_POT_AZIZ:
            v = aziz_eval(r, &g);
            break;
        case ORC_POT_DIPOLE: /* src/potentials/dipole.cpp:5-45: V = s/r^3, grad = -3 s r_vec / r^5 */
            v = par / (r * r * r);
            g = -3.0 * par / (r * r * r * r * r);
            break;
        case ORC_POT_HARMONIC: { /* src/potentials/harmonic.cpp:3-23: k = m w^2, V = k r^2/2, grad = k r_vec */
            double k = mass * par * par;
            v = 0.5 * k * r * r;
            g = k;
            break;
        }
        default: /* include/potentials/potential.h:12-19 */
            break;
    }
    if (g_over_r) *g_over_r = g;
    return v;
}

static double int_param(const orc_sim* s) {
    return s->c.int_pot == ORC_POT_DIPOLE ? s->c.int_strength : s->c.int_omega;
}

/* src/simulation.cpp:499-512 */
static double separation(const orc_sim* s, const double* xb, int i, int j, double* d) {
    double r2 = 0.0;
    for (int a = 0; a < s->D; ++a) {
        double dx = xb[(size_t)i * s->D + a] - xb[(size_t)j * s->D + a];
        if (s->c.pbc) dx = min_image(dx, s->c.size);
        d[a] = dx;
        r2 += dx * dx;
    }
    return sqrt(r2);
}

/* gradient of the external potential on one bead slice, written the way the reference's gradV does:
 * free -> 0 (include/potentials/potential.h:12-24); harmonic k x (src/potentials/harmonic.cpp:21-23);
 * double_well 4 m lambda (|x|^2 - a^2) x, the prefactor summed over the axes first (src/potentials/double_well.cpp:22-40);
 * cosine -A k sin(k x + phase) per component, k = 2 pi / wavelength, wavelength = box size
 * (src/potentials/cosine.cpp:4-35, src/simulation.cpp:634-638). */
static void external_gradient(const orc_sim* s, const double* xb, double* g) {
    const int N = s->N, D = s->D;
    switch (s->c.ext_pot) {
        case ORC_POT_HARMONIC: {
            const double kext = s->c.mass * s->c.ext_omega * s->c.ext_omega;
            for (size_t q = 0; q < slab(s); ++q) g[q] = kext * xb[q];
            break;
        }
        case ORC_POT_DOUBLE_WELL: {
            const double loc2 = s->c.ext_location * s->c.ext_location;
            for (int i = 0; i < N; ++i) {
                double prefactor = 0;
                for (int a = 0; a < D; ++a) prefactor += xb[(size_t)i * D + a] * xb[(size_t)i * D + a];
                prefactor = 4 * s->c.mass * s->c.ext_strength * (prefactor - loc2);
                for (int a = 0; a < D; ++a) g[(size_t)i * D + a] = prefactor * xb[(size_t)i * D + a];
            }
            break;
        }
        case ORC_POT_COSINE: {
            const double k = 2 * 3.141592653589793238462643383279502884 / s->c.size;   /* std::numbers::pi */
            const double prefactor = -s->c.ext_amplitude * k;
            for (size_t q = 0; q < slab(s); ++q) g[q] = prefactor * sin(k * xb[q] + s->c.ext_phase);
            break;
        }
        default:
            for (size_t q = 0; q < slab(s); ++q) g[q] = 0.0;
    }
}

/* V_ext of one bead slice: harmonic (k/2) sum x^2 (src/potentials/harmonic.cpp:3-19); double_well
 * m lambda sum over particles AND axes of (x_c^2 - a^2)^2 (src/potentials/double_well.cpp:6-20 -- per component, unlike
 * its gradient); cosine A sum cos(k x_c + phase) (src/potentials/cosine.cpp:9-21). */
static double external_energy(const orc_sim* s, const double* xb) {
    switch (s->c.ext_pot) {
        case ORC_POT_HARMONIC: {
            const double kext = s->c.mass * s->c.ext_omega * s->c.ext_omega;
            double v = 0;
            for (size_t q = 0; q < slab(s); ++q) v += xb[q] * xb[q];
            return v * (0.5 * kext);
        }
        case ORC_POT_DOUBLE_WELL: {
            const double loc2 = s->c.ext_location * s->c.ext_location;
            double v = 0;
            for (size_t q = 0; q < slab(s); ++q) v += (xb[q] * xb[q] - loc2) * (xb[q] * xb[q] - loc2);
            return v * (s->c.mass * s->c.ext_strength);
        }
        case ORC_POT_COSINE: {
            const double k = 2 * 3.141592653589793238462643383279502884 / s->c.size;
            double v = 0;
            for (size_t q = 0; q < slab(s); ++q) v += cos(k * xb[q] + s->c.ext_phase);
            return v * s->c.ext_amplitude;
        }
        default:
            return 0.0;
    }
}

/* src/simulation.cpp:428-455 — external force then the i<j pair loop with strict '<' cutoff */
static void physical_forces(const orc_sim* s, const double* xb, double* out) {
    const int N = s->N, D = s->D;
    external_gradient(s, xb, out);
    for (size_t q = 0; q < slab(s); ++q) out[q] = (-1.0) * out[q];
    if (s->rc == 0.0) return;
    double d[3];
    for (int i = 0; i < N; ++i) {
        for (int j = i + 1; j < N; ++j) {
            double r = separation(s, xb, i, j, d);
            if (r < s->rc || s->rc < 0.0) {
                double g;
                orc_pair_potential(s->c.int_pot, r, int_param(s), s->c.mass, &g);
                for (int a = 0; a < D; ++a) {
                    double f1 = (-1.0) * (g * d[a]);
                    out[(size_t)i * D + a] += f1;
                    out[(size_t)j * D + a] -= f1;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ springs */
/* src/simulation.cpp:405-419 */
static void ring_spring_forces(const orc_sim* s, int b, double* out) {
    const int P = s->P;
    const double* xc = s->x + (size_t)b * slab(s);
    const double* xp = s->x + (size_t)((b - 1 + P) % P) * slab(s);
    const double* xn = s->x + (size_t)((b + 1) % P) * slab(s);
    for (size_t q = 0; q < slab(s); ++q) {
        double dp = xp[q] - xc[q], dn = xn[q] - xc[q];
        if (s->c.pbc) {
            dp = min_image(dp, s->c.size);
            dn = min_image(dn, s->c.size);
        }
        out[q] = s->k_spring * (dp + dn);
    }
}

/* src/simulation.cpp:464-486 — spring energy of the link (b-1 -> b) */
static double ring_spring_energy(const orc_sim* s, int b) {
    const int P = s->P;
    const double* xc = s->x + (size_t)b * slab(s);
    const double* xp = s->x + (size_t)((b - 1 + P) % P) * slab(s);
    double e = 0.0;
    for (size_t q = 0; q < slab(s); ++q) {
        double d = xp[q] - xc[q];
        if (s->c.pbc) d = min_image(d, s->c.size);
        e += d * d;
    }
    return e * (0.5 * s->k_spring);
}

/* ------------------------------------------------------------------ bosonic exchange (Feldman–Hirshberg) */
/* src/bosonic_exchange/bosonic_exchange_base.cpp:30-64 — separation "second minus first" */
static void bead_sep(const orc_sim* s, const double* x1, int l1, const double* x2, int l2, double* d) {
    l1 %= s->N;
    l2 %= s->N;
    for (int a = 0; a < s->D; ++a) {
        double dx = x2[(size_t)l2 * s->D + a] - x1[(size_t)l1 * s->D + a];
        if (s->c.pbc) dx = min_image(dx, s->c.size);
        d[a] = dx;
    }
}
static double bead_sep2(const orc_sim* s, const double* x1, int l1, const double* x2, int l2) {
    double d[3], r2 = 0.0;
    bead_sep(s, x1, l1, x2, l2, d);
    for (int a = 0; a < s->D; ++a) r2 += d[a] * d[a];
    return r2;
}

/* quadratic_bosonic_exchange.cpp:63-71 — E_m^k lives at E_kn[m(m+1)/2 - k] */
static double* Eat(const orc_sim* s, int m, int k) { return &s->E_kn[(size_t)m * (m + 1) / 2 - k]; }

/* beta used by the exchange class = beta/P (bosonic_exchange_base.cpp:16-18) */
static double exch_beta(const orc_sim* s) { return s->beta / s->P; }

static void exchange_prepare(orc_sim* s) {
    const int N = s->N;
    const double k = s->k_spring, beta = exch_beta(s);
    const double* first = s->x;                                 /* bead 1  (index 0)   */
    const double* last = s->x + (size_t)(s->P - 1) * slab(s);   /* bead P  (index P-1) */

    /* cycle energies, quadratic_bosonic_exchange.cpp:34-61 */
    for (int v = 0; v < N; ++v) {
        *Eat(s, v + 1, 1) = 0.5 * k * bead_sep2(s, first, v, last, v);
        for (int u = v - 1; u >= 0; --u) {
            double val = *Eat(s, v + 1, v - u) +
                         0.5 * k * (+bead_sep2(s, last, u, first, u + 1) - bead_sep2(s, first, u + 1, last, v) +
                                    bead_sep2(s, first, u, last, v));
            *Eat(s, v + 1, v - u + 1) = val;
        }
    }
    /* forward potentials, :73-99 */
    s->V[0] = 0.0;
    for (int m = 1; m <= N; ++m) {
        double shift = DBL_MAX;
        for (int kk = m; kk > 0; --kk) {
            double val = *Eat(s, m, kk) + s->V[m - kk];
            if (val < shift) shift = val;
            s->tmp[kk - 1] = val;
        }
        double denom = 0.0;
        for (int kk = m; kk > 0; --kk) denom += exp(-beta * (s->tmp[kk - 1] - shift));
        s->V[m] = shift - (1.0 / beta) * log(denom / (double)m);
    }
    /* backward potentials, :101-128 */
    s->Vb[N] = 0.0;
    for (int l = N - 1; l > 0; --l) {
        double shift = DBL_MAX;
        for (int p = l; p < N; ++p) {
            double val = *Eat(s, p + 1, p - l + 1) + s->Vb[p + 1];
            if (val < shift) shift = val;
            s->tmp[p] = val;
        }
        double denom = 0.0;
        for (int p = l; p < N; ++p) denom += 1.0 / (p + 1) * exp(-beta * (s->tmp[p] - shift));
        s->Vb[l] = shift - log(denom) / beta;
    }
    s->Vb[0] = s->V[N];
    /* connection probabilities, :142-157 (entries with u > l+1 stay 0) */
    memset(s->prob, 0, sizeof(double) * (size_t)N * N);
    for (int l = 0; l < N - 1; ++l)
        s->prob[(size_t)N * l + (l + 1)] = 1.0 - exp(-beta * (s->V[l + 1] + s->Vb[l + 1] - s->V[N]));
    for (int u = 0; u < N; ++u)
        for (int l = u; l < N; ++l)
            s->prob[(size_t)N * l + u] =
                1.0 / (l + 1) * exp(-beta * (s->V[u] + *Eat(s, l + 1, l - u + 1) + s->Vb[l + 1] - s->V[N]));
}

/* quadratic_bosonic_exchange.cpp:188-215 (bead index 0) */
static void exchange_force_first(const orc_sim* s, double* out) {
    const int N = s->N, D = s->D;
    const double* x = s->x;
    const double* xprev = s->x + (size_t)(s->P - 1) * slab(s);
    const double* xnext = s->x + (size_t)(1 % s->P) * slab(s);
    double d[3];
    for (int l = 0; l < N; ++l) {
        double acc[3] = {0, 0, 0};
        for (int u = (l - 1 > 0 ? l - 1 : 0); u < N; ++u) {
            bead_sep(s, x, l, xprev, u, d);
            double pr = s->prob[(size_t)N * u + l];
            for (int a = 0; a < D; ++a) acc[a] += pr * d[a];
        }
        bead_sep(s, x, l, xnext, l, d);
        for (int a = 0; a < D; ++a) out[(size_t)l * D + a] = (acc[a] + d[a]) * s->k_spring;
    }
}

/* quadratic_bosonic_exchange.cpp:159-186 (bead index P-1) */
static void exchange_force_last(const orc_sim* s, double* out) {
    const int N = s->N, D = s->D, P = s->P;
    const double* x = s->x + (size_t)(P - 1) * slab(s);
    const double* xnext = s->x;
    const double* xprev = s->x + (size_t)((P - 2 + P) % P) * slab(s);
    double d[3];
    for (int l = 0; l < N; ++l) {
        double acc[3] = {0, 0, 0};
        for (int u = 0; u <= l + 1 && u < N; ++u) {
            bead_sep(s, x, l, xnext, u, d);
            double pr = s->prob[(size_t)N * l + u];
            for (int a = 0; a < D; ++a) acc[a] += pr * d[a];
        }
        bead_sep(s, x, l, xprev, l, d);
        for (int a = 0; a < D; ++a) out[(size_t)l * D + a] = (acc[a] + d[a]) * s->k_spring;
    }
}

/* quadratic_bosonic_exchange.cpp:250-279 */
static double exchange_prim_estimator(orc_sim* s) {
    const int N = s->N;
    const double beta = exch_beta(s);
    s->prim[0] = 0.0;
    for (int m = 1; m <= N; ++m) {
        double shift = DBL_MAX;
        for (int k = m; k > 0; --k) {
            double val = *Eat(s, m, k) + s->V[m - k];
            if (val < shift) shift = val;
        }
        double sig = 0.0;
        for (int k = m; k > 0; --k) {
            double e = *Eat(s, m, k);
            sig += (s->prim[m - k] - e) * exp(-beta * (e + s->V[m - k] - shift));
        }
        s->prim[m] = sig / (m * exp(-beta * (s->V[m] - shift)));
    }
    return s->prim[N] / s->P; /* IPI convention */
}


/* ------------------------------------------------------------------ factorial exchange (-DFACTORIAL_BOSONIC_ALGORITHM) */
/* src/bosonic_exchange/factorial_bosonic_exchange.cpp: every quantity is a sum over all N! permutations, enumerated with
 * next_permutation from the identity; labels[l] is the particle whose first bead follows the last bead of particle l. */
static int next_permutation_int(int* a, int n) { /* std::next_permutation */
    int i = n - 1;
    while (i > 0 && a[i - 1] >= a[i]) --i;
    if (i <= 0) { for (int l = 0, r = n - 1; l < r; ++l, --r) { int t = a[l]; a[l] = a[r]; a[r] = t; } return 0; }
    int j = n - 1;
    while (a[j] <= a[i - 1]) --j;
    { int t = a[i - 1]; a[i - 1] = a[j]; a[j] = t; }
    for (int l = i, r = n - 1; l < r; ++l, --r) { int t = a[l]; a[l] = a[r]; a[r] = t; }
    return 1;
}
/* getMinExteriorSpringEnergy :57-81 */
static double factorial_e_shift(const orc_sim* s) {
    const int N = s->N;
    const double* first = s->x;
    const double* last = s->x + (size_t)(s->P - 1) * slab(s);
    int labels[16];
    for (int i = 0; i < N; ++i) labels[i] = i;
    double min_delta = DBL_MAX;
    do {
        double diff2 = 0.0;
        for (int l = 0; l < N; ++l) diff2 += bead_sep2(s, first, l, last, labels[l]);   /* (as written in the reference) */
        double e = 0.5 * s->k_spring * diff2;
        if (e < min_delta) min_delta = e;
    } while (next_permutation_int(labels, N));
    return min_delta;
}
/* effectivePotential :90-113 */
static double factorial_effective_potential(const orc_sim* s) {
    const int N = s->N;
    const double* first = s->x;
    const double* last = s->x + (size_t)(s->P - 1) * slab(s);
    const double beta = exch_beta(s), beta_half_k = beta * 0.5 * s->k_spring;
    int labels[16];
    for (int i = 0; i < N; ++i) labels[i] = i;
    long count = 0;
    double sum = 0.0;
    do {
        ++count;
        double diff2 = 0.0;
        for (int l = 0; l < N; ++l) diff2 += bead_sep2(s, first, l, last, labels[l]);
        sum += exp(-beta_half_k * diff2);
    } while (next_permutation_int(labels, N));
    return (-1.0 / beta) * log(sum / count);
}
/* springForceLastBead :121-170 (which = 1) / springForceFirstBead :177-227 (which = 0) */
static void factorial_force(const orc_sim* s, int which, double e_shift, double* out) {
    const int N = s->N, D = s->D, P = s->P;
    const double beta = exch_beta(s), k = s->k_spring;
    const double* x = which ? s->x + (size_t)(P - 1) * slab(s) : s->x;
    const double* other = which ? s->x : s->x + (size_t)(P - 1) * slab(s);           /* x_next of the last / x_prev of the first */
    const double* inner = which ? s->x + (size_t)((P - 2 + P) % P) * slab(s) : s->x + (size_t)(1 % P) * slab(s);
    int labels[16];
    for (int i = 0; i < N; ++i) labels[i] = i;
    double temp[16][3], denom = 0.0, d[3], din[3];
    for (size_t q = 0; q < slab(s); ++q) out[q] = 0.0;
    do {
        double weight = 0.0;
        for (int l = 0; l < N; ++l) {
            int nb;
            if (which) nb = labels[l];                                                /* lastBeadNeighbor  */
            else { nb = 0; while (labels[nb] != l) ++nb; }                            /* firstBeadNeighbor */
            bead_sep(s, x, l, other, nb, d);
            for (int a = 0; a < D; ++a) weight += d[a] * d[a];
            bead_sep(s, x, l, inner, l, din);
            for (int a = 0; a < D; ++a) temp[l][a] = (din[a] + d[a]) * k;
        }
        weight = exp(-beta * (0.5 * k * weight - e_shift));
        for (int l = 0; l < N; ++l)
            for (int a = 0; a < D; ++a) out[(size_t)l * D + a] += weight * temp[l][a];
        denom += weight;
    } while (next_permutation_int(labels, N));
    for (size_t q = 0; q < slab(s); ++q) out[q] /= denom;
}
/* primEstimator :258-291 */
static double factorial_prim_estimator(const orc_sim* s, double e_shift) {
    const int N = s->N;
    const double beta = exch_beta(s);
    const double* first = s->x;
    const double* last = s->x + (size_t)(s->P - 1) * slab(s);
    int labels[16];
    for (int i = 0; i < N; ++i) labels[i] = i;
    double num = 0.0, denom = 0.0;
    do {
        double w2 = 0.0;
        for (int l = 0; l < N; ++l) {
            int nb = 0;
            while (labels[nb] != l) ++nb;                                             /* firstBeadNeighbor(l) */
            w2 += bead_sep2(s, first, l, last, nb);
        }
        double de = 0.5 * s->k_spring * w2, w = exp(-beta * (de - e_shift));
        num += de * w;
        denom += w;
    } while (next_permutation_int(labels, N));
    return (-1.0) * (num / denom) / s->P;   /* IPI convention */
}

/* ------------------------------------------------------------------ force assembly */
static int bosonic_active(const orc_sim* s) { return s->c.bosonic && s->P > 1; } /* src/simulation.cpp:690 */

/* src/simulation.cpp:353-374, 394-420 for all beads */
void orc_update_forces(orc_sim* s) {
    const int P = s->P;
    const int fact = bosonic_active(s) && s->c.factorial;
    double e_shift = 0.0;
    if (fact) e_shift = factorial_e_shift(s);                 /* FactorialBosonicExchange::prepare :18-20 */
    else if (bosonic_active(s)) exchange_prepare(s);
    for (int b = 0; b < P; ++b) {
        double* fs = s->f_spring + (size_t)b * slab(s);
        double* fp = s->f_phys + (size_t)b * slab(s);
        if (fact && (b == 0 || b == P - 1))
            factorial_force(s, b == P - 1, e_shift, fs);
        else if (bosonic_active(s) && b == 0)
            exchange_force_first(s, fs);
        else if (bosonic_active(s) && b == P - 1)
            exchange_force_last(s, fs);
        else
            ring_spring_forces(s, b, fs);
        physical_forces(s, s->x + (size_t)b * slab(s), fp);
        double* f = s->f + (size_t)b * slab(s);
        for (size_t q = 0; q < slab(s); ++q) f[q] = fs[q] + fp[q];
    }
}

/* ------------------------------------------------------------------ propagators */
/* src/propagators/velocity_verlet.cpp:24-38 */
static void moment_step(orc_sim* s, const double* force) {
    size_t n = (size_t)s->P * slab(s);
    for (size_t q = 0; q < n; ++q) s->p[q] += 0.5 * s->c.dt * force[q];
}
static void coords_step(orc_sim* s) {
    size_t n = (size_t)s->P * slab(s);
    for (size_t q = 0; q < n; ++q) s->x[q] += s->c.dt * s->p[q] / s->c.mass;
}

/* src/normal_modes.cpp:48-79 — row k of the real orthogonal Cartesian->NM matrix, and the row the
 * reference stores for the inverse on rank j (column j of the same matrix) */
static void build_nm(orc_sim* s) {
    const int P = s->P;
    for (int k = 0; k < P; ++k) {
        double fund = 2.0 * M_PI / P * k;
        double* row = s->nm_fwd + (size_t)k * P;
        if (k == 0) {
            for (int j = 0; j < P; ++j) row[j] = 1.0 / sqrt((double)P);
        } else if (k < 0.5 * P) {
            for (int j = 0; j < P; ++j) row[j] = sqrt(2.0 / P) * cos(fund * j);
        } else if (k == 0.5 * P) {
            for (int j = 0; j < P; ++j) row[j] = 1.0 / sqrt((double)P) * (j % 2 == 0 ? 1.0 : -1.0);
        } else {
            for (int j = 0; j < P; ++j) row[j] = -sqrt(2.0 / P) * sin(fund * j);
        }
        /* inverse row held by rank "k" (normal_modes.cpp:70-78); here index = this_bead */
        double* inv = s->nm_inv + (size_t)k * P;
        double pref = sqrt(2.0 / P);
        memset(inv, 0, sizeof(double) * P);
        inv[0] = 1.0 / sqrt((double)P);
        for (int i = 1; i < 0.5 * P; ++i) inv[i] = pref * cos(fund * i);
        if (P % 2 == 0) inv[P / 2] = 1.0 / sqrt((double)P) * (k % 2 == 0 ? 1.0 : -1.0);
        for (int i = (int)ceil(0.5 * (P + 1)); i < P; ++i) inv[i] = -pref * sin(fund * i);
    }
}

/* length-P dot products over the bead axis (src/normal_modes.cpp:99-129) */
static void nm_apply(const orc_sim* s, const double* mat, const double* src, double* dst) {
    const int P = s->P;
    const size_t sl = slab(s);
    for (int k = 0; k < P; ++k)
        for (size_t q = 0; q < sl; ++q) {
            double acc = 0;
            for (int j = 0; j < P; ++j) acc += mat[(size_t)k * P + j] * src[(size_t)j * sl + q];
            dst[(size_t)k * sl + q] = acc;
        }
}

/* src/propagators/normal_modes_propagator.cpp:19-103 (distinguishable branch; Params rejects bosonic+NM,
 * src/params.cpp:146-148) */
static void nm_propagator_step(orc_sim* s) {
    const int P = s->P;
    const size_t sl = slab(s), n = (size_t)P * sl;
    moment_step(s, s->nm_ext); /* half kick with the physical (external + pair) forces only */
    nm_apply(s, s->nm_fwd, s->x, s->scratch_x);
    nm_apply(s, s->nm_fwd, s->p, s->scratch_p);
    for (int k = 0; k < P; ++k) {
        double freq = 2 * s->omega_p * sin(k * M_PI / P);
        double c = cos(freq * s->c.dt), sn = sin(freq * s->c.dt), mw = s->c.mass * freq;
        for (size_t q = 0; q < sl; ++q) {
            double xq = s->scratch_x[(size_t)k * sl + q], pq = s->scratch_p[(size_t)k * sl + q];
            if (freq == 0) {
                s->scratch_x[(size_t)k * sl + q] = xq + s->c.dt / s->c.mass * pq;
                s->scratch_p[(size_t)k * sl + q] = pq;
            } else {
                s->scratch_x[(size_t)k * sl + q] = c * xq + sn / mw * pq;
                s->scratch_p[(size_t)k * sl + q] = (-1) * mw * sn * xq + c * pq;
            }
        }
    }
    nm_apply(s, s->nm_inv, s->scratch_x, s->x);
    nm_apply(s, s->nm_inv, s->scratch_p, s->p);
    orc_update_forces(s);
    memcpy(s->nm_ext, s->f_phys, sizeof(double) * n);
    moment_step(s, s->nm_ext);
}

void orc_propagator_step(orc_sim* s) {
    if (s->c.propagator == ORC_PROP_NORMAL_MODES) {
        nm_propagator_step(s);
        return;
    }
    /* src/propagators/velocity_verlet.cpp:7-22 */
    moment_step(s, s->f);
    coords_step(s);
    orc_update_forces(s);
    moment_step(s, s->f);
}

/* ------------------------------------------------------------------ thermostat */
/* src/thermostats/thermostat.cpp:15-19, langevin.cpp:10-27, thermostat_coupling.cpp:29-47 */
/* src/thermostats/nose_hoover.cpp:102-146 — half-step chain integrator; returns the momentum scaling factor */
static double nh_chain_step(const orc_sim* s, double current_energy, double ndof, double* eta, double* ed, double* edd) {
    const int nc = s->c.nchains;
    const double Qi = s->beta / s->P, Q1 = ndof * Qi;            /* :16-21, hbar = 1 */
    const double dt2 = 0.5 * s->c.dt, dt4 = 0.25 * s->c.dt, dt8 = 0.125 * s->c.dt;
    const double required = ndof / s->thermo_beta;
    double exp_factor = 0.0;
    edd[0] = (current_energy - required) / Q1;
    ed[nc - 1] += edd[nc - 1] * dt4;
    for (int i = nc - 2; i >= 0; i--) {
        exp_factor = exp(-dt8 * ed[i + 1]);
        ed[i] *= exp_factor;
        ed[i] += edd[i] * dt4;
        ed[i] *= exp_factor;
    }
    double scale = exp(-dt2 * ed[0]);
    for (int i = 0; i < nc; i++) eta[i] += dt2 * ed[i];
    edd[0] = (current_energy * scale * scale - required) / Q1;
    ed[0] *= exp_factor;
    ed[0] += edd[0] * dt4;
    ed[0] *= exp_factor;
    double Q_former = Q1;
    for (int i = 1; i < nc - 1; i++) {
        exp_factor = exp(-dt8 * ed[i + 1]);
        ed[i] *= exp_factor;
        edd[i] = (Q_former * ed[i - 1] * ed[i - 1] - 1 / s->thermo_beta) / Qi;
        ed[i] += edd[i] * dt4;
        ed[i] *= exp_factor;
        Q_former = Qi;
    }
    if (nc >= 2) { /* nchains = 1 reads eta_dot[-1] in the reference (App. A-12); not restated */
        edd[nc - 1] = (Qi * ed[nc - 2] * ed[nc - 2] - 1 / s->thermo_beta) / Qi;
        ed[nc - 1] += edd[nc - 1] * dt4;
    }
    return scale;
}

/* momentaUpdate of the three variants, src/thermostats/nose_hoover.cpp:69-91, 162-182, 207-222, on the momenta in
 * `pall`: the Cartesian momenta (CartesianCoupling) or the normal-mode momenta (NMCoupling,
 * src/thermostats/thermostat_coupling.cpp:29-47: rank k thermostats mode k with its own chains) */
static void nose_hoover_apply(orc_sim* s, double* pall) {
    const int P = s->P, N = s->N, D = s->D, nc = s->c.nchains;
    for (int b = 0; b < P; ++b) {
        double* pb = pall + (size_t)b * slab(s);
        size_t base = (size_t)b * s->nh_groups * nc;
        if (s->c.thermostat == ORC_THERMO_NOSE_HOOVER) {
            double e = 0.0;
            for (size_t q = 0; q < slab(s); ++q) e += pb[q] * pb[q];
            e /= s->c.mass;
            double sc = nh_chain_step(s, e, (double)D * N, s->nh_eta + base, s->nh_ed + base, s->nh_edd + base);
            for (size_t q = 0; q < slab(s); ++q) pb[q] = pb[q] * sc;
        } else if (s->c.thermostat == ORC_THERMO_NOSE_HOOVER_NP) {
            for (int i = 0; i < N; ++i) {
                double e = 0.0;
                for (int a = 0; a < D; ++a) e += pb[(size_t)i * D + a] * pb[(size_t)i * D + a];
                e /= s->c.mass;
                size_t o = base + (size_t)i * nc;
                double sc = nh_chain_step(s, e, (double)D, s->nh_eta + o, s->nh_ed + o, s->nh_edd + o);
                for (int a = 0; a < D; ++a) pb[(size_t)i * D + a] = pb[(size_t)i * D + a] * sc;
            }
        } else {
            for (int i = 0; i < N; ++i)
                for (int a = 0; a < D; ++a) {
                    double pv = pb[(size_t)i * D + a];
                    size_t o = base + ((size_t)i * D + a) * nc;
                    double sc = nh_chain_step(s, pv * pv / s->c.mass, 1.0, s->nh_eta + o, s->nh_ed + o, s->nh_edd + o);
                    pb[(size_t)i * D + a] = pv * sc;
                }
        }
    }
}

/* getAdditionToH summed over beads, src/thermostats/nose_hoover.cpp:37-67, 184-190, 224-232 */
static double nose_hoover_energy(const orc_sim* s) {
    const int nc = s->c.nchains;
    const double ndof = s->c.thermostat == ORC_THERMO_NOSE_HOOVER ? (double)s->D * s->N
                        : (s->c.thermostat == ORC_THERMO_NOSE_HOOVER_NP ? (double)s->D : 1.0);
    const double Qi = s->beta / s->P, Q1 = ndof * Qi;
    double total = 0.0;
    for (int b = 0; b < s->P; ++b) {
        double per_bead = 0.0;
        for (int g = 0; g < s->nh_groups; ++g) {
            size_t o = ((size_t)b * s->nh_groups + g) * nc;
            double h = 0.5 * Q1 * s->nh_ed[o] * s->nh_ed[o];
            h += ndof * s->nh_eta[o] / s->thermo_beta;
            for (int i = 1; i < nc; i++) {
                h += 0.5 * Qi * s->nh_ed[o + i] * s->nh_ed[o + i];
                h += s->nh_eta[o + i] / s->thermo_beta;
            }
            per_bead += h;
        }
        total += per_bead;
    }
    return total;
}

void orc_thermostat_step(orc_sim* s) {
    if (s->c.thermostat >= ORC_THERMO_NOSE_HOOVER) {
        if (!s->c.nmthermostat) {
            nose_hoover_apply(s, s->p);
        } else {   /* Thermostat::step, src/thermostats/thermostat.cpp:15-19: share, update mode momenta, transform back */
            nm_apply(s, s->nm_fwd, s->p, s->scratch_p);
            nose_hoover_apply(s, s->scratch_p);
            nm_apply(s, s->nm_inv, s->scratch_p, s->p);
        }
        return;
    }
    if (s->c.thermostat != ORC_THERMO_LANGEVIN) return;
    const int P = s->P, N = s->N, D = s->D;
    const size_t sl = slab(s);
    double c1 = exp(-0.5 * s->c.gamma * s->c.dt);
    double c2 = sqrt((1 - c1 * c1) * s->c.mass / s->thermo_beta);
    if (!s->c.nmthermostat) {
        for (int b = 0; b < P; ++b)
            for (int i = 0; i < N; ++i)
                for (int a = 0; a < D; ++a) {
                    double noise = orc_ranmars_gaussian(s->rng[b]);
                    size_t q = idx(s, b, i, a);
                    s->p[q] = c1 * s->p[q] + c2 * noise;
                }
        return;
    }
    nm_apply(s, s->nm_fwd, s->p, s->scratch_p);
    for (int k = 0; k < P; ++k)
        for (int i = 0; i < N; ++i)
            for (int a = 0; a < D; ++a) {
                double noise = orc_ranmars_gaussian(s->rng[k]);
                size_t q = (size_t)k * sl + (size_t)i * D + a;
                s->scratch_p[q] = c1 * s->scratch_p[q] + c2 * noise;
            }
    nm_apply(s, s->nm_inv, s->scratch_p, s->p);
}

/* src/simulation.cpp:581-603 — per-bead partial divided by N*P, then summed over beads in rank order */
void orc_zero_momentum(orc_sim* s) {
    const int P = s->P, N = s->N, D = s->D;
    double cm[3] = {0, 0, 0};
    for (int b = 0; b < P; ++b) {
        double part[3] = {0, 0, 0};
        for (int i = 0; i < N; ++i)
            for (int a = 0; a < D; ++a) part[a] += s->p[idx(s, b, i, a)];
        for (int a = 0; a < D; ++a) cm[a] += part[a] / (N * P);
    }
    for (int b = 0; b < P; ++b)
        for (int i = 0; i < N; ++i)
            for (int a = 0; a < D; ++a) s->p[idx(s, b, i, a)] -= cm[a];
}

/* src/simulation.cpp:246-259 */
void orc_run_iteration(orc_sim* s) {
    orc_thermostat_step(s);
    if (s->c.fixcom) orc_zero_momentum(s);
    orc_propagator_step(s);
    orc_thermostat_step(s);
    if (s->c.fixcom) orc_zero_momentum(s);
}

/* ------------------------------------------------------------------ observables */
/* src/observables/energy.cpp:30-111, classical.cpp:32-78, bosonic.cpp:17-22; bead partials summed in
 * bead order like ObservablesLogger::log (observable.cpp:92-116). Values stay in atomic units. */
void orc_observables_calc(orc_sim* s, orc_observables* o) {
    const int P = s->P, N = s->N, D = s->D;
    const int bos = bosonic_active(s);
    memset(o, 0, sizeof *o);
    const int fact = bos && s->c.factorial;
    const double fact_shift = fact ? factorial_e_shift(s) : 0.0;
    if (bos && !fact) exchange_prepare(s);
    for (int b = 0; b < P; ++b) {
        const double* xb = s->x + (size_t)b * slab(s);
        /* kinetic (primitive estimator) */
        double kin = 0.5 * D * N / s->beta;
        if (b == 0 && bos)
            kin += fact ? factorial_prim_estimator(s, fact_shift) : exchange_prim_estimator(s);
        else
            kin -= ring_spring_energy(s, b) / P;
        o->kinetic += kin;
        /* potential + "virial" */
        double pot = 0, vir = 0, ipot = 0, epot = 0;
        if (s->c.ext_pot != ORC_POT_FREE) {
            epot = external_energy(s, xb);
            pot += epot;
            external_gradient(s, xb, s->scratch_slab);
            for (size_t q = 0; q < slab(s); ++q) vir -= xb[q] * ((-1.0) * s->scratch_slab[q]);
        }
        if (s->rc != 0.0) {
            double d[3];
            for (int i = 0; i < N; ++i)
                for (int j = i + 1; j < N; ++j) {
                    double r = separation(s, xb, i, j, d);
                    if (r < s->rc || s->rc < 0.0) {
                        double g;
                        double v = orc_pair_potential(s->c.int_pot, r, int_param(s), s->c.mass, &g);
                        pot += v;
                        ipot += v;
                        for (int a = 0; a < D; ++a) vir -= xb[(size_t)i * D + a] * ((-1.0) * (g * d[a]));
                    }
                }
        }
        if (s->c.ext_pot != ORC_POT_FREE && s->c.int_pot != ORC_POT_FREE) {
            o->ext_pot += epot / P;
            o->int_pot += ipot / P;
        }
        if (s->c.ext_pot != ORC_POT_FREE || s->c.int_pot != ORC_POT_FREE) {
            o->potential += pot / P;
            o->virial += vir * (0.5 / P);
        }
        /* gsf: src/observables/gsf_action.cpp:21-73, per bead, summed by the logger. With an interaction potential
         * the reference adds a one-row gradient to an N-row array and reads past its end (:36): not restated. */
        if (s->c.int_pot == ORC_POT_FREE) {
            const double alpha = 0.0;
            double total_potential = external_energy(s, xb);
            external_gradient(s, xb, s->scratch_slab);
            double total_force_squared = 0.0;
            for (size_t q = 0; q < slab(s); ++q) total_force_squared += s->scratch_slab[q] * s->scratch_slab[q];
            double sp_constant = s->k_spring / P;                      /* IPI_CONVENTION (include/common.h:40-42) */
            double potential_term = total_potential / (3 * P);
            double force_squared_term = total_force_squared / (9 * sp_constant * P * P);
            double w;
            if (b % 2 != 0) {
                w = (-1.0) * potential_term + alpha * force_squared_term;
                o->pot_gsf += total_potential / (0.5 * P);
            } else {
                w = potential_term + (1 - alpha) * force_squared_term;
            }
            o->w_gsf += w * ((-1.0) * s->beta);
        } else {
            o->w_gsf = o->pot_gsf = NAN;
        }
        /* classical */
        double ke = 0;
        const double* pb = s->p + (size_t)b * slab(s);
        for (size_t q = 0; q < slab(s); ++q) ke += pb[q] * pb[q];
        ke *= 0.5 / s->c.mass;
        o->cl_kinetic += ke;
        double dof = (double)D * N * P;
        o->temperature += 2.0 * ke / dof / P;
        o->cl_spring += (b == 0 && bos) ? (fact ? factorial_effective_potential(s) : s->V[N]) : ring_spring_energy(s, b);
    }
    if (s->c.thermostat >= ORC_THERMO_NOSE_HOOVER) o->nh_energy = nose_hoover_energy(s); /* classical.cpp:24-26 */
    if (bos && !fact) {   /* (the factorial class returns 0 for both, factorial_bosonic_exchange.cpp:234-251) */
        /* quadratic_bosonic_exchange.cpp:222-240 */
        double beta = exch_beta(s), sum = 0;
        for (int m = 1; m <= N; ++m) sum += *Eat(s, m, 1);
        o->prob_dist = exp(-beta * (sum - s->V[N]) - lgamma(N + 1));
        o->prob_all = exp(-beta * (*Eat(s, N, N) - s->V[N]));
    }
}

/* ------------------------------------------------------------------ lifecycle */
orc_sim* orc_create(const orc_config* cfg) {
    orc_sim* s = (orc_sim*)calloc(1, sizeof *s);
    s->c = *cfg;
    s->N = cfg->natoms;
    s->P = cfg->nbeads;
    s->D = cfg->ndim;
    s->beta = 1.0 / cfg->temperature;
    s->thermo_beta = s->beta / s->P;
    s->omega_p = s->P / s->beta;
    s->k_spring = cfg->mass * s->omega_p * s->omega_p;
    /* src/simulation.cpp:84-90 */
    s->rc = (cfg->int_pot == ORC_POT_FREE) ? 0.0 : cfg->cutoff;
    if (cfg->pbc) s->rc = fmin(s->rc, 0.5 * cfg->size);
    size_t n = (size_t)s->P * s->N * s->D;
    s->x = calloc(n, sizeof(double));
    s->p = calloc(n, sizeof(double));
    s->f = calloc(n, sizeof(double));
    s->f_spring = calloc(n, sizeof(double));
    s->f_phys = calloc(n, sizeof(double));
    s->nm_ext = calloc(n, sizeof(double));
    s->scratch_x = calloc(n, sizeof(double));
    s->scratch_p = calloc(n, sizeof(double));
    s->scratch_slab = calloc((size_t)s->N * s->D, sizeof(double));
    size_t N = s->N;
    s->E_kn = calloc(N * (N + 1) / 2, sizeof(double));
    s->V = calloc(N + 1, sizeof(double));
    s->Vb = calloc(N + 1, sizeof(double));
    s->prob = calloc(N * N, sizeof(double));
    s->tmp = calloc(N, sizeof(double));
    s->prim = calloc(N + 1, sizeof(double));
    s->nm_fwd = calloc((size_t)s->P * s->P, sizeof(double));
    s->nm_inv = calloc((size_t)s->P * s->P, sizeof(double));
    build_nm(s);
    s->nh_groups = cfg->thermostat == ORC_THERMO_NOSE_HOOVER_NP ? s->N : (cfg->thermostat == ORC_THERMO_NOSE_HOOVER_NP_DIM ? s->N * s->D : 1);
    {
        size_t len = (size_t)s->P * s->nh_groups * (cfg->nchains > 0 ? cfg->nchains : 1);
        s->nh_eta = calloc(len, sizeof(double));
        s->nh_ed = calloc(len, sizeof(double));
        s->nh_edd = calloc(len, sizeof(double));
    }
    s->rng = calloc(s->P, sizeof(orc_ranmars*));
    for (int b = 0; b < s->P; ++b) s->rng[b] = orc_ranmars_new((int)(cfg->seed + (unsigned)b));
    return s;
}

void orc_destroy(orc_sim* s) {
    if (!s) return;
    for (int b = 0; b < s->P; ++b) orc_ranmars_free(s->rng[b]);
    free(s->rng);
    free(s->x); free(s->p); free(s->f); free(s->f_spring); free(s->f_phys); free(s->nm_ext);
    free(s->scratch_x); free(s->scratch_p); free(s->scratch_slab);
    free(s->E_kn); free(s->V); free(s->Vb); free(s->prob); free(s->tmp); free(s->prim);
    free(s->nm_fwd); free(s->nm_inv);
    free(s->nh_eta); free(s->nh_ed); free(s->nh_edd);
    free(s);
}

double orc_beta(const orc_sim* s) { return s->beta; }
double orc_spring_constant(const orc_sim* s) { return s->k_spring; }
double orc_cutoff_effective(const orc_sim* s) { return s->rc; }

static double* pick(const orc_sim* s, char which) {
    switch (which) {
        case 'x': return s->x;
        case 'p': return s->p;
        case 'f': return s->f;
        case 's': return s->f_spring;
        case 'e': return s->f_phys;
        default: return NULL;
    }
}
void orc_set(orc_sim* s, char which, const double* src) {
    double* d = pick(s, which);
    if (d) memcpy(d, src, sizeof(double) * (size_t)s->P * slab(s));
}
void orc_get(const orc_sim* s, char which, double* dst) {
    const double* d = pick(s, which);
    if (d) memcpy(dst, d, sizeof(double) * (size_t)s->P * slab(s));
}

int orc_exchange_get(const orc_sim* s, char which, double* dst) {
    size_t N = s->N, n = 0;
    const double* src = NULL;
    switch (which) {
        case 'V': src = s->V; n = N + 1; break;
        case 'B': src = s->Vb; n = N + 1; break;
        case 'E': src = s->E_kn; n = N * (N + 1) / 2; break;
        case 'P': src = s->prob; n = N * N; break;
        default: return 0;
    }
    if (dst) memcpy(dst, src, n * sizeof(double));
    return (int)n;
}


/* ------------------------------------------------------------------ extended-precision exchange (rounding-noise yardstick)
 * The same Feldman-Hirshberg algorithm (quadratic_bosonic_exchange.cpp:34-215: cycle-energy recurrence, shifted
 * log-sum-exp recursions for V and V_backwards, connection probabilities, exterior spring forces) with every
 * intermediate in `long double` (x87, 64-bit mantissa). The double-precision algorithm loses digits where beta*V is
 * large: the exponent -beta (V[u] + E + Vb[l+1] - V[N]) is a difference of numbers ~beta*|V| (6e4 at N = 8192), so the
 * connection probabilities of the reference itself carry ~1e-11 .. 1e-9 of rounding noise at N >= 512 (rows of its
 * probability matrix sum to 1 only to 1e-11). This version says which of two double-precision results is closer to
 * the algorithm's exact answer. Host AoS slices [N][D]; x2 = bead after the first, xPm1 = bead before the last. */
static long double ld_sep2(int D, int pbc, long double L, const double* xa, int ia, const double* xb, int ib, long double* d) {
    long double r2 = 0.0L;
    for (int a = 0; a < D; ++a) {
        long double dx = (long double)xb[(size_t)ib * D + a] - (long double)xa[(size_t)ia * D + a];
        if (pbc) dx -= L * floorl(dx / L + 0.5L);
        if (d) d[a] = dx;
        r2 += dx * dx;
    }
    return r2;
}

int orc_exchange_ld(int N, int D, int pbc, double size, double kspring, double beta_exch, const double* x1, const double* xP,
                    const double* x2, const double* xPm1, double* V_out, double* Vb_out, double* f_first, double* f_last,
                    double* prim_out, int l_stride) {
    const long double L = size, k = kspring, beta = beta_exch;
    const size_t ntri = (size_t)N * (N + 1) / 2;
    long double* E = (long double*)malloc(sizeof(long double) * ntri);
    long double* V = (long double*)malloc(sizeof(long double) * (N + 1));
    long double* Vb = (long double*)malloc(sizeof(long double) * (N + 1));
    long double* tmp = (long double*)malloc(sizeof(long double) * (N + 1));
    if (!E || !V || !Vb || !tmp) { free(E); free(V); free(Vb); free(tmp); return -1; }
#define ELD(m, kk) E[(size_t)(m) * ((m) + 1) / 2 - (kk)]
    /* (OpenMP where the build has it -- rows, force particles and the terms of a sum are independent; in long double the
     * order of a sum does not matter at the 1e-10 level this yardstick serves) */
#pragma omp parallel for schedule(dynamic, 16)
    for (int v = 0; v < N; ++v) {
        ELD(v + 1, 1) = 0.5L * k * ld_sep2(D, pbc, L, x1, v, xP, v, NULL);
        for (int u = v - 1; u >= 0; --u)
            ELD(v + 1, v - u + 1) = ELD(v + 1, v - u) + 0.5L * k * (ld_sep2(D, pbc, L, xP, u, x1, u + 1, NULL) -
                                                                   ld_sep2(D, pbc, L, x1, u + 1, xP, v, NULL) +
                                                                   ld_sep2(D, pbc, L, x1, u, xP, v, NULL));
    }
    V[0] = 0.0L;
    for (int m = 1; m <= N; ++m) {
        long double shift = LDBL_MAX;
        for (int kk = m; kk > 0; --kk) {
            tmp[kk - 1] = ELD(m, kk) + V[m - kk];
            if (tmp[kk - 1] < shift) shift = tmp[kk - 1];
        }
        long double denom = 0.0L;
#pragma omp parallel for reduction(+ : denom) if (m > 512)
        for (int kk = m; kk > 0; --kk) denom += expl(-beta * (tmp[kk - 1] - shift));
        V[m] = shift - logl(denom / (long double)m) / beta;
    }
    Vb[N] = 0.0L;
    for (int l = N - 1; l > 0; --l) {
        long double shift = LDBL_MAX;
        for (int p = l; p < N; ++p) {
            tmp[p] = ELD(p + 1, p - l + 1) + Vb[p + 1];
            if (tmp[p] < shift) shift = tmp[p];
        }
        long double denom = 0.0L;
#pragma omp parallel for reduction(+ : denom) if (N - l > 512)
        for (int p = l; p < N; ++p) denom += expl(-beta * (tmp[p] - shift)) / (long double)(p + 1);
        Vb[l] = shift - logl(denom) / beta;
    }
    Vb[0] = V[N];
    for (int i = 0; i <= N; ++i) { if (V_out) V_out[i] = (double)V[i]; if (Vb_out) Vb_out[i] = (double)Vb[i]; }
    if (f_first) {   /* :188-215: sum over u >= l-1 of P(u -> l) (r^P_u - r^1_l) + (r^2_l - r^1_l) */
#pragma omp parallel for schedule(dynamic, 16)
        for (int l = 0; l < N; l += l_stride) {   /* (l_stride > 1: a sample of the particles, the others are left untouched) */
            long double d[3];
            long double acc[3] = {0, 0, 0};
            for (int u = (l - 1 > 0 ? l - 1 : 0); u < N; ++u) {
                long double pr;
                if (l == u + 1) pr = 1.0L - expl(-beta * (V[u + 1] + Vb[u + 1] - V[N]));
                else pr = expl(-beta * (V[l] + ELD(u + 1, u - l + 1) + Vb[u + 1] - V[N])) / (long double)(u + 1);
                ld_sep2(D, pbc, L, x1, l, xP, u, d);
                for (int a = 0; a < D; ++a) acc[a] += pr * d[a];
            }
            ld_sep2(D, pbc, L, x1, l, x2, l, d);
            for (int a = 0; a < D; ++a) f_first[(size_t)l * D + a] = (double)((acc[a] + d[a]) * k);
        }
    }
    if (f_last) {    /* :159-186: sum over u <= l+1 of P(l -> u) (r^1_u - r^P_l) + (r^{P-1}_l - r^P_l) */
#pragma omp parallel for schedule(dynamic, 16)
        for (int l = 0; l < N; l += l_stride) {
            long double d[3];
            long double acc[3] = {0, 0, 0};
            for (int u = 0; u <= l + 1 && u < N; ++u) {
                long double pr;
                if (u == l + 1) pr = 1.0L - expl(-beta * (V[l + 1] + Vb[l + 1] - V[N]));
                else pr = expl(-beta * (V[u] + ELD(l + 1, l - u + 1) + Vb[l + 1] - V[N])) / (long double)(l + 1);
                ld_sep2(D, pbc, L, xP, l, x1, u, d);
                for (int a = 0; a < D; ++a) acc[a] += pr * d[a];
            }
            ld_sep2(D, pbc, L, xP, l, xPm1, l, d);
            for (int a = 0; a < D; ++a) f_last[(size_t)l * D + a] = (double)((acc[a] + d[a]) * k);
        }
    }
    if (prim_out) {  /* :250-279, without the 1/P of the caller */
        long double* e = tmp;
        long double* prim = (long double*)malloc(sizeof(long double) * (N + 1));
        prim[0] = 0.0L;
        for (int m = 1; m <= N; ++m) {
            long double shift = LDBL_MAX;
            for (int kk = m; kk > 0; --kk) { long double val = ELD(m, kk) + V[m - kk]; if (val < shift) shift = val; }
            long double sig = 0.0L;
            for (int kk = m; kk > 0; --kk) sig += (prim[m - kk] - ELD(m, kk)) * expl(-beta * (ELD(m, kk) + V[m - kk] - shift));
            prim[m] = sig / (m * expl(-beta * (V[m] - shift)));
        }
        *prim_out = (double)prim[N];
        free(prim);
        (void)e;
    }
#undef ELD
    free(E); free(V); free(Vb); free(tmp);
    return 0;
}
