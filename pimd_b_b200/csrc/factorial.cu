// The reference's other exchange class: the sum over all N! permutations (-DFACTORIAL_BOSONIC_ALGORITHM,
// src/bosonic_exchange/factorial_bosonic_exchange.cpp), for N <= kFactMaxN. A different potential from the
// Feldman-Hirshberg one on the same positions (same partition function, other forces), kept because the reference ships
// it and two of its twelve golden regression cases belong to it.
//
//   weight of a permutation sigma (particle l's last bead is followed by the first bead of sigma(l)):
//       w_sigma = exp(-beta (k/2 sum_l |r^1_sigma(l) - r^P_l|^2 - e_shift)),   e_shift = min_sigma of the bracket   (:57-81)
//   exterior forces (:121-227) need only the marginals  M[l][j] = sum_{sigma(l) = j} w_sigma / Z,  Z = sum_sigma w_sigma:
//       last bead  l:  k [ mi(r^{P-1}_l - r^P_l) + sum_j M[l][j] mi(r^1_j - r^P_l) ]
//       first bead j:  k [ mi(r^2_j - r^1_j)     + sum_l M[l][j] mi(r^P_l - r^1_j) ]
//   effectivePotential (:90-113) = e_shift - ln(Z / N!) / beta,   primEstimator (:258-291) = -(k/2) sum_lj M[l][j] d2(l,j) / P.
// One thread block per (l, j) sums the (N-1)! permutations with sigma(l) = j (decoded from their index), in a fixed order:
// no atomics, bit-reproducible. The reference enumerates with next_permutation; only the order of the sums differs.
#include "internal.cuh"
#include "device_utils.cuh"

namespace pimdb {

constexpr int kFactMaxN = 10;

struct FactArgs {
    const double *x1, *xP, *x2, *xPm1;   // bead 1, bead P, bead 2, bead P-1 slices [D][N]
    double* work;                        // d[N][N][3] | d2[N][N] | macc[N][N] | pmin[N]
    double *F, *V;                       // F[2][D][N], V[N+1]
    DevObs* obs;
    int N, D, pbc, do_first, do_last;
    double k, beta, L, invL;
    long long nperm_rest;                // (N-1)!
};

__device__ __forceinline__ double* fact_d(const FactArgs& a) { return a.work; }
__device__ __forceinline__ double* fact_d2(const FactArgs& a) { return a.work + 3 * a.N * a.N; }
__device__ __forceinline__ double* fact_macc(const FactArgs& a) { return a.work + 4 * a.N * a.N; }
__device__ __forceinline__ double* fact_pmin(const FactArgs& a) { return a.work + 5 * a.N * a.N; }

// d(l, j) = mi(r^1_j - r^P_l) and its square, bosonic_exchange_base.cpp:30-64
__global__ void k_fact_prep(FactArgs a) {
    const int i = threadIdx.x;
    if (i >= a.N * a.N) return;
    const int l = i / a.N, j = i % a.N;
    double r2 = 0.0;
    for (int c = 0; c < a.D; ++c) {
        double dx = a.x1[(size_t)c * a.N + j] - a.xP[(size_t)c * a.N + l];
        if (a.pbc) dx = min_image(dx, a.L, a.invL);
        fact_d(a)[(size_t)i * 3 + c] = dx;
        r2 += dx * dx;
    }
    fact_d2(a)[i] = r2;
}

// sum of d2 over the permutation number t of the remaining particles, given sigma(l0) = j0
__device__ __forceinline__ double fact_perm_d2(const double* sd2, int N, int l0, int j0, long long t) {
    int avail[kFactMaxN];
    int n = 0;
    for (int v = 0; v < N; ++v) if (v != j0) avail[n++] = v;
    double diff2 = sd2[l0 * N + j0];
    for (int l = 0; l < N; ++l) {
        if (l == l0) continue;
        const int pick = (int)(t % n);
        t /= n;
        diff2 += sd2[l * N + avail[pick]];
        for (int q = pick; q + 1 < n; ++q) avail[q] = avail[q + 1];
        --n;
    }
    return diff2;
}

// MODE 0: block j -> min over the permutations with sigma(0) = j of (k/2) diff2;  MODE 1: block (l, j) -> sum of weights
template <int MODE>
__global__ void __launch_bounds__(256) k_fact_perms(FactArgs a) {
    __shared__ double sd2[kFactMaxN * kFactMaxN];
    __shared__ double red[32];
    const int N = a.N, tid = threadIdx.x;
    for (int i = tid; i < N * N; i += blockDim.x) sd2[i] = fact_d2(a)[i];
    __syncthreads();
    const int l0 = MODE == 0 ? 0 : blockIdx.x / N, j0 = MODE == 0 ? blockIdx.x : blockIdx.x % N;
    double e_shift = 0.0;
    if (MODE == 1) {
        e_shift = fact_pmin(a)[0];
        for (int j = 1; j < N; ++j) e_shift = fmin(e_shift, fact_pmin(a)[j]);
    }
    double acc = MODE == 0 ? 1.0e300 : 0.0;
    for (long long t = tid; t < a.nperm_rest; t += blockDim.x) {
        const double e = 0.5 * a.k * fact_perm_d2(sd2, N, l0, j0, t);
        if (MODE == 0) acc = fmin(acc, e);
        else acc += exp(-a.beta * (e - e_shift));
    }
    if (MODE == 0) {
        acc = -acc;                       // block max of the negated value = min
        acc = warp_max(acc);
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid < 32) {
            double v = tid < (int)(blockDim.x >> 5) ? red[tid] : -1.0e300;
            v = warp_max(v);
            if (tid == 0) fact_pmin(a)[blockIdx.x] = -v;
        }
    } else {
        double v[1] = {acc};
        block_sum<1>(v, red);
        if (tid == 0) fact_macc(a)[blockIdx.x] = v[0];
    }
}

__global__ void __launch_bounds__(128) k_fact_finish(FactArgs a) {
    __shared__ double sM[kFactMaxN * kFactMaxN];
    __shared__ double sZ, sShift;
    const int N = a.N, tid = threadIdx.x;
    if (tid == 0) {
        double z = 0.0;
        for (int j = 0; j < N; ++j) z += fact_macc(a)[j];          // fixed order: the row of particle 0
        double e_shift = fact_pmin(a)[0];
        for (int j = 1; j < N; ++j) e_shift = fmin(e_shift, fact_pmin(a)[j]);
        sZ = z; sShift = e_shift;
    }
    __syncthreads();
    for (int i = tid; i < N * N; i += blockDim.x) sM[i] = fact_macc(a)[i] / sZ;
    __syncthreads();
    const double* d = fact_d(a);
    // last bead (which = 1): particle l = tid;  first bead (which = 0): particle j = tid - N
    if (tid < N && a.do_last) {
        const int l = tid;
        for (int c = 0; c < a.D; ++c) {
            double acc = 0.0;
            for (int j = 0; j < N; ++j) acc += sM[l * N + j] * d[(size_t)(l * N + j) * 3 + c];
            double din = a.xPm1[(size_t)c * N + l] - a.xP[(size_t)c * N + l];
            if (a.pbc) din = min_image(din, a.L, a.invL);
            a.F[(size_t)(1 * a.D + c) * N + l] = (din + acc) * a.k;
        }
    } else if (tid >= N && tid < 2 * N && a.do_first) {
        const int j = tid - N;
        for (int c = 0; c < a.D; ++c) {
            double acc = 0.0;
            for (int l = 0; l < N; ++l) acc += sM[l * N + j] * (-d[(size_t)(l * N + j) * 3 + c]);
            double din = a.x2[(size_t)c * N + j] - a.x1[(size_t)c * N + j];
            if (a.pbc) din = min_image(din, a.L, a.invL);
            a.F[(size_t)(0 * a.D + c) * N + j] = (din + acc) * a.k;
        }
    }
    if (tid == 0) {
        double nfact = 1.0;
        for (int i = 2; i <= N; ++i) nfact *= i;
        const double veff = sShift - log(sZ / nfact) / a.beta;
        double prim = 0.0;
        for (int i = 0; i < N * N; ++i) prim += fact_d2(a)[i] * sM[i];
        for (int i = 0; i < N; ++i) a.V[i] = 0.0;
        a.V[N] = veff;
        a.obs->prim_est = -0.5 * a.k * prim;      // the host divides by P (i-PI convention), like e[N] of the FH estimator
        a.obs->v_n = veff;
        a.obs->e_diag_sum = 0.0;
        a.obs->e_full = 0.0;
    }
}

int launch_factorial_exchange(Sim* s, cudaStream_t st) {
    FactArgs a;
    const size_t S = s->S;
    if (s->has_first) {
        a.x1 = s->x + 1 * S;
        a.xP = s->all_local ? s->x + (size_t)s->Ploc * S : s->x;
        a.x2 = s->x + 2 * S;
    } else {
        a.x1 = s->x + (size_t)(s->Ploc + 1) * S;
        a.xP = s->x + (size_t)s->Ploc * S;
        a.x2 = a.x1;                     // (unused: this handle does not own the first bead)
    }
    a.xPm1 = s->has_last ? s->x + (size_t)(s->Ploc - 1) * S : a.xP;
    a.work = s->fact_work; a.F = s->exF; a.V = s->exV; a.obs = s->obs_d;
    a.N = s->N; a.D = s->D; a.pbc = s->cfg.pbc; a.do_first = s->has_first; a.do_last = s->has_last;
    a.k = s->kspring; a.beta = s->exch_beta; a.L = s->L; a.invL = 1.0 / s->L;
    a.nperm_rest = 1;
    for (int i = 2; i < s->N; ++i) a.nperm_rest *= i;
    k_fact_prep<<<1, 128, 0, st>>>(a);
    k_fact_perms<0><<<s->N, 256, 0, st>>>(a);
    k_fact_perms<1><<<s->N * s->N, 256, 0, st>>>(a);
    k_fact_finish<<<1, 128, 0, st>>>(a);
    s->launches += 4;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

}  // namespace pimdb
