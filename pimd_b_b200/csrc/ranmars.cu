// Reference-compatible noise stream (optional, pimdb_config.rng = PIMDB_RNG_RANMARS): one RANMAR generator per bead,
// seeded with seed + bead, drawing one gaussian per particle and axis in particle-major order every thermostat
// half-step -- exactly what LangevinThermostat::momentaUpdate does with Simulation::mars_gen
// (reference src/thermostats/langevin.cpp:15-27, src/simulation.cpp:58-59, libs/random_mars.cpp).
//
// The generator is sequential per bead, so this mode costs ~N*NDIM dependent draws per half-step (one thread per
// bead fills a noise slab that the fused integrator / normal-mode kernels then read instead of evaluating Philox). It
// exists for trajectory-level parity with the reference's thermostatted golden cases, not for speed; the default
// stream stays the counter-based Philox4x32-10 ("Noise stream").
#include "internal.cuh"

namespace pimdb {

// libs/random_mars.cpp:10-56 -- seed expansion into 97 24-bit fractions; one uniform is burnt at the end
static void ranmars_seed(RanMarsState& r, int seed) {
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
        r.u[ii] = s;
    }
    r.u[0] = 0.0;
    r.c = 362436.0 / 16777216.0;
    r.i97 = 97;
    r.j97 = 33;
    r.have_spare = 0;
    r.spare = 0.0;
}

// libs/random_mars.cpp:62-77 -- lagged-Fibonacci subtract-with-borrow step plus the arithmetic sequence c
__host__ __device__ static inline double ranmars_uniform(RanMarsState& r) {
    const double cd = 7654321.0 / 16777216.0, cm = 16777213.0 / 16777216.0;
    double v = r.u[r.i97] - r.u[r.j97];
    if (v < 0.0) v += 1.0;
    r.u[r.i97] = v;
    if (--r.i97 == 0) r.i97 = 97;
    if (--r.j97 == 0) r.j97 = 97;
    r.c -= cd;
    if (r.c < 0.0) r.c += cm;
    v -= r.c;
    if (v < 0.0) v += 1.0;
    return v;
}

// libs/random_mars.cpp:83-103 -- polar Box-Muller; the FIRST value returned is v2*fac, v1*fac is cached
__device__ static inline double ranmars_gaussian(RanMarsState& r) {
    if (r.have_spare) {
        r.have_spare = 0;
        return r.spare;
    }
    double v1, v2, rsq;
    do {
        v1 = 2.0 * ranmars_uniform(r) - 1.0;
        v2 = 2.0 * ranmars_uniform(r) - 1.0;
        rsq = v1 * v1 + v2 * v2;
    } while (rsq >= 1.0 || rsq == 0.0);
    const double fac = sqrt(-2.0 * log(rsq) / rsq);
    r.spare = v1 * fac;
    r.have_spare = 1;
    return v2 * fac;
}

// one thread per owned bead (or mode): noise[b][particle][axis] in the order the reference consumes it
__global__ void k_ranmars_fill(RanMarsState* st, double* noise, int Ploc, int count) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= Ploc) return;
    RanMarsState r = st[b];
    double* out = noise + (size_t)b * count;
    for (int i = 0; i < count; ++i) out[i] = ranmars_gaussian(r);
    st[b] = r;
}

int ranmars_create(Sim* s) {
    const unsigned long long last = s->cfg.seed + (unsigned long long)(s->b1 > 0 ? s->b1 - 1 : 0);
    if (s->cfg.seed == 0 || last > 900000000ull) {   // libs/random_mars.cpp:14-15
        s->err = "Invalid seed for Marsaglia random # generator";
        return PIMDB_ERR_INVALID_ARGUMENT;
    }
    std::vector<RanMarsState> h(s->Ploc);
    for (int b = 0; b < s->Ploc; ++b) {
        ranmars_seed(h[b], (int)(s->cfg.seed + (unsigned long long)(s->b0 + b)));
        (void)ranmars_uniform(h[b]);
    }
    PIMDB_CUDA_TRY(s, cudaMalloc(&s->rm_state, sizeof(RanMarsState) * s->Ploc));
    PIMDB_CUDA_TRY(s, cudaMemcpy(s->rm_state, h.data(), sizeof(RanMarsState) * s->Ploc, cudaMemcpyHostToDevice));
    PIMDB_CUDA_TRY(s, cudaMalloc(&s->rm_noise, sizeof(double) * s->S * s->Ploc));
    PIMDB_CUDA_TRY(s, cudaMemset(s->rm_noise, 0, sizeof(double) * s->S * s->Ploc));
    return PIMDB_OK;
}

int launch_ranmars_fill(Sim* s) {
    k_ranmars_fill<<<(s->Ploc + 31) / 32, 32, 0, s->stream>>>(s->rm_state, s->rm_noise, s->Ploc, (int)s->S);
    s->launches += 1;
    PIMDB_CUDA_TRY(s, cudaGetLastError());
    return PIMDB_OK;
}

}  // namespace pimdb
