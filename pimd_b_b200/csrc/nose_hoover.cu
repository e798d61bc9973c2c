// Nose-Hoover chain thermostats: nose_hoover (one chain per bead), nose_hoover_np (one chain
// per bead and particle), nose_hoover_np_dim (one chain per degree of freedom). Deterministic, so whole trajectories
// are comparable with the reference and with three of its golden cases.
//
// Reference: src/thermostats/nose_hoover.cpp — constructor :10-30 (Qi = hbar^2 beta / P, Q1 = ndof Qi, required energy
// ndof / beta_P), singleChainStep :102-146 (half-step chain integrator after Tuckerman et al. 2006 / LAMMPS),
// momentaUpdate :69-91 / :162-182 / :207-222, getAdditionToH :37-67 / :184-190 / :224-232.
// The integrator is restated operation by operation, including the reuse of the last `exp_factor` of the downward
// sweep in the update of the first chain element (:126) -- with nchains = 1 that factor is the initial 0.0 in the
// reference and here too.
#include "internal.cuh"
#include "device_utils.cuh"

namespace pimdb {

struct NhArgs {
    double* p;
    double *eta, *eta_dot, *eta_ddot;   // [bead][group][nchains]
    double* part;                       // additionToH partials (obs), [blocks]
    int N, D, Ploc, nchains, mode;      // mode 0: per bead, 1: per particle, 2: per degree of freedom
    size_t S;
    double Q1, Qi, dt2, dt4, dt8, required, inv_beta, inv_mass;
};

// singleChainStep: returns the momentum scaling factor
__device__ __forceinline__ double nh_chain_step(const NhArgs& a, double energy, double* eta, double* ed, double* edd) {
    const int nc = a.nchains;
    double exp_factor = 0.0;
    edd[0] = (energy - a.required) / a.Q1;
    ed[nc - 1] += edd[nc - 1] * a.dt4;
    for (int i = nc - 2; i >= 0; --i) {
        exp_factor = exp(-a.dt8 * ed[i + 1]);
        ed[i] *= exp_factor;
        ed[i] += edd[i] * a.dt4;
        ed[i] *= exp_factor;
    }
    const double scale = exp(-a.dt2 * ed[0]);
    for (int i = 0; i < nc; ++i) eta[i] += a.dt2 * ed[i];
    edd[0] = (energy * scale * scale - a.required) / a.Q1;
    ed[0] *= exp_factor;
    ed[0] += edd[0] * a.dt4;
    ed[0] *= exp_factor;
    double q_former = a.Q1;
    for (int i = 1; i < nc - 1; ++i) {
        exp_factor = exp(-a.dt8 * ed[i + 1]);
        ed[i] *= exp_factor;
        edd[i] = (q_former * ed[i - 1] * ed[i - 1] - a.inv_beta) / a.Qi;
        ed[i] += edd[i] * a.dt4;
        ed[i] *= exp_factor;
        q_former = a.Qi;
    }
    if (nc >= 2) {   // nchains == 1 indexes eta_dot[-1] in the reference; nothing sensible to restate there
        edd[nc - 1] = (a.Qi * ed[nc - 2] * ed[nc - 2] - a.inv_beta) / a.Qi;
        ed[nc - 1] += edd[nc - 1] * a.dt4;
    }
    return scale;
}

// mode 0: one block per owned bead; the bead's kinetic sum is reduced in a fixed order
__global__ void __launch_bounds__(256) k_nh_bead(NhArgs a) {
    __shared__ double sm[32];
    __shared__ double s_scale;
    const int b = blockIdx.x;
    double* pb = a.p + (size_t)b * a.S;
    double e[1] = {0.0};
    for (size_t i = threadIdx.x; i < a.S; i += blockDim.x) e[0] = fma(pb[i], pb[i], e[0]);
    block_sum<1>(e, sm);
    if (threadIdx.x == 0) {
        const size_t o = (size_t)b * a.nchains;
        s_scale = nh_chain_step(a, e[0] * a.inv_mass, a.eta + o, a.eta_dot + o, a.eta_ddot + o);
    }
    __syncthreads();
    const double scale = s_scale;
    for (size_t i = threadIdx.x; i < a.S; i += blockDim.x) pb[i] *= scale;
}

// modes 1 and 2: one thread per chain
__global__ void __launch_bounds__(256) k_nh_local(NhArgs a) {
    const long long groups = a.mode == 1 ? (long long)a.N : (long long)a.N * a.D;
    const long long total = (long long)a.Ploc * groups;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / groups);
        const long long g = idx % groups;
        double* pb = a.p + (size_t)b * a.S;
        const size_t o = (size_t)idx * a.nchains;   // reference order: particle-major (np), (particle*NDIM + axis) (np_dim)
        if (a.mode == 1) {
            const int n = (int)g;
            double e = 0.0;
            for (int c = 0; c < a.D; ++c) { const double v = pb[(size_t)c * a.N + n]; e += v * v; }
            const double scale = nh_chain_step(a, e * a.inv_mass, a.eta + o, a.eta_dot + o, a.eta_ddot + o);
            for (int c = 0; c < a.D; ++c) pb[(size_t)c * a.N + n] *= scale;
        } else {
            const int n = (int)(g / a.D), c = (int)(g % a.D);
            const double v = pb[(size_t)c * a.N + n];
            const double scale = nh_chain_step(a, v * v * a.inv_mass, a.eta + o, a.eta_dot + o, a.eta_ddot + o);
            pb[(size_t)c * a.N + n] = v * scale;
        }
    }
}

// getAdditionToH summed over owned beads (fixed order: one block, strided partials, block reduction)
__global__ void __launch_bounds__(256) k_nh_energy(NhArgs a, double ndof_first, long long nchains_total, double* out) {
    __shared__ double sm[32];
    double e[1] = {0.0};
    for (long long ch = threadIdx.x; ch < nchains_total; ch += blockDim.x) {
        const double* eta = a.eta + ch * a.nchains;
        const double* ed = a.eta_dot + ch * a.nchains;
        double h = 0.5 * a.Q1 * ed[0] * ed[0] + ndof_first * eta[0] * a.inv_beta;
        for (int i = 1; i < a.nchains; ++i) h += 0.5 * a.Qi * ed[i] * ed[i] + eta[i] * a.inv_beta;
        e[0] += h;
    }
    block_sum<1>(e, sm);
    if (threadIdx.x == 0) *out = e[0];
}

static NhArgs make_nh(Sim* s) {
    NhArgs a;
    a.p = s->p;
    a.eta = s->nh_state;
    a.eta_dot = s->nh_state + s->nh_len;
    a.eta_ddot = s->nh_state + 2 * s->nh_len;
    a.part = nullptr;
    a.N = s->N; a.D = s->D; a.Ploc = s->Ploc; a.nchains = s->cfg.nchains; a.S = s->S;
    a.mode = s->cfg.thermostat - PIMDB_THERMO_NOSE_HOOVER;
    const double ndof = a.mode == 0 ? (double)s->D * s->N : (a.mode == 1 ? (double)s->D : 1.0);
    a.Qi = s->beta / s->P;                      // hbar = 1
    a.Q1 = ndof * a.Qi;
    a.dt2 = 0.5 * s->cfg.dt; a.dt4 = 0.25 * s->cfg.dt; a.dt8 = 0.125 * s->cfg.dt;
    a.required = ndof / s->thermo_beta;
    a.inv_beta = 1.0 / s->thermo_beta;
    a.inv_mass = 1.0 / s->cfg.mass;
    return a;
}

int launch_nose_hoover(Sim* s) {
    NhArgs a = make_nh(s);
    if (a.mode == 0) {
        k_nh_bead<<<s->Ploc, 256, 0, s->stream>>>(a);
    } else {
        const size_t total = (size_t)s->Ploc * (a.mode == 1 ? s->N : (size_t)s->N * s->D);
        k_nh_local<<<grid_for(total, 256), 256, 0, s->stream>>>(a);
    }
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

int launch_nose_hoover_energy(Sim* s, double* out_dev) {
    NhArgs a = make_nh(s);
    const double ndof = a.mode == 0 ? (double)s->D * s->N : (a.mode == 1 ? (double)s->D : 1.0);
    const long long chains = (long long)(s->nh_len / s->cfg.nchains);
    k_nh_energy<<<1, 256, 0, s->stream>>>(a, ndof, chains, out_dev);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

}  // namespace pimdb
