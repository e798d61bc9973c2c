// Bosonic exchange (Feldman-Hirshberg, O(N^2 + PN)) on the GPU: K4-K8 and the exchange estimators of K13.
//
// Reference: src/bosonic_exchange/quadratic_bosonic_exchange.cpp
//   evaluateCycleEnergies :34-61, evaluateVBn :73-99, evaluateVBackwards :101-128,
//   evaluateConnectionProbabilities :142-157, springForceLastBead :159-186, springForceFirstBead :188-215,
//   getDistinctProbability :222-229, getLongestProbability :238-240, primEstimator :250-279
// and src/bosonic_exchange/bosonic_exchange_base.cpp:30-64 (minimum-image bead separations).
//
// Design (see DESIGN.md "exchange"):
//   * E_kn is never materialised. With d2(u,v) = |r^P_v - r^1_u|^2 (minimum image) and the prefix sum
//     A(w) = sum_{t<w} |r^1_{t+1} - r^P_t|^2, the reference's recurrence telescopes to
//         E^{[u..v]} = k/2 [ A(v) - A(u) + d2(u,v) ],
//     so every entry is 1 subtraction pattern + one distance, evaluated on the fly.
//   * The N-step recursions run column-wise: as soon as V[j] is known every row m>j folds the term
//     -beta (E^{[j..m-1]} + V[j]) into its own running (max, sum) pair ("online" log-sum-exp), so a step costs
//     one barrier instead of two block reductions, and row j+1 is complete the moment column j has been applied.
//     Forward and backward recursions are independent and run as two concurrent thread blocks.
//   * Connection probabilities are evaluated on the fly inside the exterior-force kernel (one warp per particle
//     and exterior bead); the N x N matrix is only built when a caller asks for it (pimdb_exchange_get).
#include "internal.cuh"
#include "device_utils.cuh"

namespace pimdb {

struct ExArgs {
    const double *x1, *xP;     // bead 1 and bead P slices, [D][N]
    const double *x2, *xPm1;   // bead 2 (next of first) and bead P-1 (previous of last)
    double *A, *V, *Vb, *F;    // A[N], V[N+1], Vb[N+1], F[2][D][N]
    double* prim;              // e[N+1] scratch of the primitive-estimator recursion
    DevObs* obs;
    int* err;
    int N, D, pbc, do_first, do_last;
    double k, beta, L, invL;
};

template <int D>
__device__ __forceinline__ double dist2(const ExArgs& a, const double* xa, int ia, const double* xb, int ib) {
    double r2 = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        double dx = xb[(size_t)c * a.N + ib] - xa[(size_t)c * a.N + ia];
        if (a.pbc) dx = min_image(dx, a.L, a.invL);
        r2 = fma(dx, dx, r2);
    }
    return r2;
}

// E^{[u..v]} for u <= v.
template <int D>
__device__ __forceinline__ double cycle_energy(const ExArgs& a, int u, int v) {
    return 0.5 * a.k * (a.A[v] - a.A[u] + dist2<D>(a, a.x1, u, a.xP, v));
}

// ---------------------------------------------------------------- prefix sums A(w), by one whole block
// Called at the top of both recursion blocks (each fills its own copy of A, so the two blocks never wait for
// each other): N distances + a block scan, a few microseconds.
template <int D>
__device__ __forceinline__ void prefix_block(const ExArgs& a, double* A) {
    __shared__ double warp_tot[32];
    __shared__ double carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 32) warp_tot[tid] = 0.0;
    if (tid == 0) carry = 0.0;
    __syncthreads();
    // A[w] = sum_{t<w} link[t], link[t] = d2(P_t, 1_{t+1}); processed in chunks of blockDim, inclusive scan per chunk
    for (int base = 0; base < a.N; base += blockDim.x) {
        const int w = base + tid;                      // produces A[w+1]
        double link = (w < a.N - 1) ? dist2<D>(a, a.xP, w, a.x1, w + 1) : 0.0;
        double v = link;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(kFullMask, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        if (warp == 0) {
            double t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                double u = __shfl_up_sync(kFullMask, t, o);
                if (lane >= o) t += u;
            }
            warp_tot[lane] = t;
        }
        __syncthreads();
        double incl = carry + (warp > 0 ? warp_tot[warp - 1] : 0.0) + v;
        if (w + 1 < a.N) A[w + 1] = incl;
        __syncthreads();
        if (tid == blockDim.x - 1) carry = incl;
        __syncthreads();
    }
    if (tid == 0) A[0] = 0.0;
    __syncthreads();
}

// ---------------------------------------------------------------- forward / backward recursions
// online log-sum-exp accumulator
struct Lse {
    double mx, s;
    __device__ __forceinline__ void init() { mx = -INFINITY; s = 0.0; }
    __device__ __forceinline__ void add(double t) {
        if (t > mx) { s = s * exp(mx - t) + 1.0; mx = t; }
        else s += exp(t - mx);
    }
};

// block 0: V[1..N]  (V[0] = 0);  block 1: Vb[N-1..1]  (Vb[N] = 0)
template <int D, int R>
__global__ void __launch_bounds__(1024) k_exch_recursion(ExArgs a) {
    extern __shared__ double sv[];   // V or Vb values, N+1 doubles
    const int tid = threadIdx.x, nt = blockDim.x, N = a.N;
    const double beta = a.beta;
    Lse acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r].init();
    a.A += (size_t)blockIdx.x * N;   // private copy of the prefix sums (copy 0 is the one later kernels read)
    prefix_block<D>(a, a.A);

    if (blockIdx.x == 0) {
        // thread owns rows v = tid + r*nt  (V index m = v+1)
        if (tid == 0) sv[0] = 0.0;
        __syncthreads();
        for (int j = 0; j < N; ++j) {
            const double vj = sv[j];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int v = tid + r * nt;
                if (v >= j && v < N) acc[r].add(-beta * (cycle_energy<D>(a, j, v) + vj));
            }
            // row v == j is now complete
            const int owner = j % nt, rr = j / nt;
            if (tid == owner) {
                double mx = 0.0, s = 1.0;
#pragma unroll
                for (int r = 0; r < R; ++r) if (r == rr) { mx = acc[r].mx; s = acc[r].s; }
                double val = -(mx + log(s / (double)(j + 1))) / beta;
                if (!isfinite(val)) atomicOr(a.err, kErrOverflowFwd);
                sv[j + 1] = val;
                a.V[j + 1] = val;
            }
            __syncthreads();
        }
        if (tid == 0) a.V[0] = 0.0;
    } else {
        // thread owns rows l = tid + r*nt, l in [1, N-1]; column q = p+1 from N down to 2
        if (tid == 0) sv[N] = 0.0;
        __syncthreads();
        for (int q = N; q >= 2; --q) {
            const int p = q - 1;
            const double vq = sv[q];
            const double lq = log((double)q);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int l = tid + r * nt;
                if (l >= 1 && l <= p) acc[r].add(-beta * (cycle_energy<D>(a, l, p) + vq) - lq);
            }
            const int owner = p % nt, rr = p / nt;   // row l == p complete
            if (tid == owner) {
                double mx = 0.0, s = 1.0;
#pragma unroll
                for (int r = 0; r < R; ++r) if (r == rr) { mx = acc[r].mx; s = acc[r].s; }
                double val = -(mx + log(s)) / beta;
                if (!isfinite(val)) atomicOr(a.err, kErrOverflowBwd);
                sv[p] = val;
                a.Vb[p] = val;
            }
            __syncthreads();
        }
        if (tid == 0) a.Vb[N] = 0.0;
    }
}

// ---------------------------------------------------------------- exterior spring forces (K7 + K8)
// one warp per (exterior bead, particle l)
template <int D>
__global__ void __launch_bounds__(256) k_exch_forces(ExArgs a) {
    const int lane = threadIdx.x & 31;
    const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int N = a.N;
    if (w >= 2 * N) return;
    const int which = w / N, l = w % N;   // 0: first bead, 1: last bead
    if ((which == 0 && !a.do_first) || (which == 1 && !a.do_last)) return;
    const double beta = a.beta, VN = a.V[N];
    double acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.0;

    if (which == 0) {
        // f_l = k [ sum_{u=max(0,l-1)}^{N-1} P(u->l) mi(r^P_u - r^1_l) + mi(r^2_l - r^1_l) ]
        const double Vl = a.V[l];
        for (int u = max(0, l - 1) + lane; u < N; u += 32) {
            double pr;
            if (u == l - 1) pr = 1.0 - exp(-beta * (Vl + a.Vb[l] - VN));
            else pr = exp(-beta * (Vl + cycle_energy<D>(a, l, u) + a.Vb[u + 1] - VN)) / (double)(u + 1);
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dx = a.xP[(size_t)c * N + u] - a.x1[(size_t)c * N + l];
                if (a.pbc) dx = min_image(dx, a.L, a.invL);
                acc[c] = fma(pr, dx, acc[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = warp_sum(acc[c]);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dx = a.x2[(size_t)c * N + l] - a.x1[(size_t)c * N + l];
                if (a.pbc) dx = min_image(dx, a.L, a.invL);
                a.F[(size_t)c * N + l] = (acc[c] + dx) * a.k;
            }
        }
    } else {
        // f_l = k [ sum_{u=0}^{min(l+1,N-1)} P(l->u) mi(r^1_u - r^P_l) + mi(r^{P-1}_l - r^P_l) ]
        const double Vbl1 = a.Vb[l + 1];
        const int uend = min(l + 1, N - 1);
        for (int u = lane; u <= uend; u += 32) {
            double pr;
            if (u == l + 1) pr = 1.0 - exp(-beta * (a.V[l + 1] + Vbl1 - VN));
            else pr = exp(-beta * (a.V[u] + cycle_energy<D>(a, u, l) + Vbl1 - VN)) / (double)(l + 1);
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dx = a.x1[(size_t)c * N + u] - a.xP[(size_t)c * N + l];
                if (a.pbc) dx = min_image(dx, a.L, a.invL);
                acc[c] = fma(pr, dx, acc[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = warp_sum(acc[c]);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dx = a.xPm1[(size_t)c * N + l] - a.xP[(size_t)c * N + l];
                if (a.pbc) dx = min_image(dx, a.L, a.invL);
                a.F[(size_t)(D + c) * N + l] = (acc[c] + dx) * a.k;
            }
        }
    }
    if (w == 0 && lane == 0) a.Vb[0] = VN;   // V_backwards[0] = V[N] (quadratic_bosonic_exchange.cpp:127)
}

// ---------------------------------------------------------------- estimators (bead-0 owner only), one block
// e[m] = sum_{j<m} w(m,j) (e[j] - E^{[j..m-1]}),  w(m,j) = exp(-beta (E^{[j..m-1]} + V[j] - V[m])) / m
// Column-wise like the recursions; the dependency chain per step is one FMA (no exp/log on it).
template <int D, int R>
__global__ void __launch_bounds__(1024) k_exch_estimators(ExArgs a) {
    extern __shared__ double se[];   // e[0..N]
    __shared__ double red[32];
    const int tid = threadIdx.x, nt = blockDim.x, N = a.N;
    const double beta = a.beta;
    double acc[R], vm[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        acc[r] = 0.0;
        const int v = tid + r * nt;
        vm[r] = v < N ? a.V[v + 1] : 0.0;
    }
    if (tid == 0) se[0] = 0.0;
    __syncthreads();
    for (int j = 0; j < N; ++j) {
        const double ej = se[j], vj = a.V[j];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int v = tid + r * nt;
            if (v >= j && v < N) {
                double e = cycle_energy<D>(a, j, v);
                double wgt = exp(-beta * (e + vj - vm[r])) / (double)(v + 1);
                acc[r] = fma(wgt, ej - e, acc[r]);
            }
        }
        const int owner = j % nt, rr = j / nt;
        if (tid == owner) {
            double val = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) if (r == rr) val = acc[r];
            se[j + 1] = val;
        }
        __syncthreads();
    }
    // sum_m E^{[m..m]} (prob_dist) in a fixed order
    double part[1] = {0.0};
    for (int m = tid; m < N; m += nt) part[0] += 0.5 * a.k * dist2<D>(a, a.x1, m, a.xP, m);
    block_sum<1>(part, red);
    if (tid == 0) {
        a.obs->prim_est = se[N];
        a.obs->v_n = a.V[N];
        a.obs->e_diag_sum = part[0];
        a.obs->e_full = cycle_energy<D>(a, 0, N - 1);
    }
}

// ---------------------------------------------------------------- on-demand tables (tests / debugging)
template <int D>
__global__ void k_exch_table_E(ExArgs a, double* out) {
    const long long tot = (long long)a.N * (a.N + 1) / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        // serial order of the reference: index = m(m+1)/2 - k, m = v+1 in [1,N], k = v-u+1 in [1,m]
        // rows m occupy [m(m-1)/2, m(m+1)/2): offset within the row = m - k = u
        long long m = (long long)((sqrt(8.0 * (double)i + 1.0) - 1.0) * 0.5) + 1;
        while (m * (m - 1) / 2 > i) --m;
        while (m * (m + 1) / 2 <= i) ++m;
        int u = (int)(i - m * (m - 1) / 2);
        out[i] = cycle_energy<D>(a, u, (int)m - 1);
    }
}

template <int D>
__global__ void k_exch_table_prob(ExArgs a, double* out) {
    const int N = a.N;
    const long long tot = (long long)N * N;
    const double beta = a.beta, VN = a.V[N];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i / N), u = (int)(i % N);
        double pr = 0.0;
        if (u == l + 1) pr = 1.0 - exp(-beta * (a.V[l + 1] + a.Vb[l + 1] - VN));
        else if (u <= l) pr = exp(-beta * (a.V[u] + cycle_energy<D>(a, u, l) + a.Vb[l + 1] - VN)) / (double)(l + 1);
        out[i] = pr;
    }
}

// ---------------------------------------------------------------- host side
static ExArgs make_args(Sim* s) {
    ExArgs a;
    const size_t S = s->S;
    // slab index (with halo offset): owned bead j -> x + (j+1) S
    if (s->has_first) {
        a.x1 = s->x + 1 * S;
        a.xP = s->all_local ? s->x + (size_t)s->Ploc * S : s->x;       // halo before first = bead P-1
        a.x2 = s->x + 2 * S;                                           // next of first (owned or trailing halo)
    } else {
        a.x1 = s->x + (size_t)(s->Ploc + 1) * S;                       // trailing halo = bead 0
        a.xP = s->x + (size_t)s->Ploc * S;
        a.x2 = nullptr;
    }
    if (s->has_last) {
        if (!s->has_first) a.xP = s->x + (size_t)s->Ploc * S;
        a.xPm1 = s->x + (size_t)(s->Ploc - 1) * S;                     // previous of last (owned or leading halo)
    } else {
        a.xPm1 = nullptr;
    }
    a.A = s->exA; a.V = s->exV; a.Vb = s->exVb; a.F = s->exF; a.prim = s->exPrim;
    a.obs = s->obs_d; a.err = s->err_d;
    a.N = s->N; a.D = s->D; a.pbc = s->cfg.pbc;
    a.do_first = s->has_first; a.do_last = s->has_last;
    a.k = s->kspring; a.beta = s->exch_beta; a.L = s->L; a.invL = 1.0 / s->L;
    return a;
}

static int rows_per_thread(int N, int& nt) {
    nt = ((N + 31) / 32) * 32;
    if (nt > 1024) nt = 1024;
    int r = (N + nt - 1) / nt;
    int R = 1;
    while (R < r) R <<= 1;
    return R;
}

template <int D>
static int run_recursion(Sim* s, const ExArgs& a, cudaStream_t st) {
    int nt;
    const int R = rows_per_thread(s->N, nt);
    const size_t smem = (size_t)(s->N + 1) * sizeof(double);
#define PIMDB_REC(RR)                                                                                          \
    case RR:                                                                                                   \
        if (smem > 48 * 1024)                                                                                  \
            cudaFuncSetAttribute(k_exch_recursion<D, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        k_exch_recursion<D, RR><<<2, nt, smem, st>>>(a);                                                       \
        break;
    switch (R) {
        PIMDB_REC(1) PIMDB_REC(2) PIMDB_REC(4) PIMDB_REC(8) PIMDB_REC(16) PIMDB_REC(32)
        default:
            s->err = "natoms too large for the single-block exchange recursion (max 32768)";
            return PIMDB_ERR_INVALID_ARGUMENT;
    }
#undef PIMDB_REC
    return PIMDB_OK;
}

template <int D>
static int exchange_impl(Sim* s, cudaStream_t st) {
    ExArgs a = make_args(s);
    int rc = run_recursion<D>(s, a, st);
    if (rc != PIMDB_OK) return rc;
    const int grid = (2 * s->N * 32 + 255) / 256;
    k_exch_forces<D><<<grid, 256, 0, st>>>(a);
    s->launches += 2;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_exchange(Sim* s, cudaStream_t st) {
    if (s->D == 1) return exchange_impl<1>(s, st);
    if (s->D == 2) return exchange_impl<2>(s, st);
    return exchange_impl<3>(s, st);
}

template <int D>
static int estimators_impl(Sim* s) {
    ExArgs a = make_args(s);
    int nt;
    const int R = rows_per_thread(s->N, nt);
    const size_t smem = (size_t)(s->N + 1) * sizeof(double);
#define PIMDB_EST(RR)                                                                                           \
    case RR:                                                                                                    \
        if (smem > 48 * 1024)                                                                                   \
            cudaFuncSetAttribute(k_exch_estimators<D, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        k_exch_estimators<D, RR><<<1, nt, smem, s->stream>>>(a);                                                \
        break;
    switch (R) {
        PIMDB_EST(1) PIMDB_EST(2) PIMDB_EST(4) PIMDB_EST(8) PIMDB_EST(16) PIMDB_EST(32)
        default:
            s->err = "natoms too large for the exchange estimator kernel (max 32768)";
            return PIMDB_ERR_INVALID_ARGUMENT;
    }
#undef PIMDB_EST
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_exchange_estimators(Sim* s) {
    if (s->D == 1) return estimators_impl<1>(s);
    if (s->D == 2) return estimators_impl<2>(s);
    return estimators_impl<3>(s);
}

template <int D>
static int tables_impl(Sim* s, int table) {
    ExArgs a = make_args(s);
    const size_t n = table == PIMDB_EXCH_E ? (size_t)s->N * (s->N + 1) / 2 : (size_t)s->N * s->N;
    if (table == PIMDB_EXCH_E) k_exch_table_E<D><<<grid_for(n, 256), 256, 0, s->stream>>>(a, s->exTab);
    else k_exch_table_prob<D><<<grid_for(n, 256), 256, 0, s->stream>>>(a, s->exTab);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_exchange_tables(Sim* s, int table) {
    if (s->D == 1) return tables_impl<1>(s, table);
    if (s->D == 2) return tables_impl<2>(s, table);
    return tables_impl<3>(s, table);
}

}  // namespace pimdb
