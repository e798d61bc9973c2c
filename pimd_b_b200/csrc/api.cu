// C ABI of libpimdb200.so (include/pimdb200.h): lifecycle, state transfer, step sequencing, CUDA-graph capture,
// deferred error reporting. No CPU fallback: every entry point needs a usable sm_100-class device.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <unistd.h>

#include "internal.cuh"

namespace pimdb {
int launch_pair_chunk(Sim* s, int bead_lo, int nb, bool with_obs, bool early = false, int scratch_lo = 0);
int pair_resident_warps_per_sm();
int launch_assemble_chunk(Sim* s, int bead_lo, int nb, bool with_pair);
int launch_exchange_part(Sim* s, cudaStream_t st, int part);
}  // namespace pimdb

using namespace pimdb;

static thread_local std::string g_create_error;
static int settle_momenta(Sim* s);
static void allow_early_launch(Sim* s);
static int maybe_download_x(Sim* s);
static int join_download_x(Sim* s);

#define API_TRY(expr)                      \
    do {                                   \
        int _rc = (expr);                  \
        if (_rc != PIMDB_OK) return _rc;   \
    } while (0)

static int fail(Sim* s, int code, const std::string& msg) {
    if (s) s->err = msg;
    else g_create_error = msg;
    return code;
}

// ----------------------------------------------------------------------------------------------------
// validation: the checks of the reference's Params (src/params.cpp:8-260) with the same messages
static int validate(const pimdb_config* c, std::string& msg) {
    char buf[256];
    if (c->nbeads < 1) {
        snprintf(buf, sizeof buf, "The specified number of beads (%d) is less than one!", c->nbeads);
        msg = buf; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (c->natoms < 1) {
        snprintf(buf, sizeof buf, "The specified number of particles (%d) is smaller than one!", c->natoms);
        msg = buf; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (c->ndim < 1 || c->ndim > 3) { msg = "NDIM must be 1, 2 or 3"; return PIMDB_ERR_INVALID_ARGUMENT; }
    if (c->propagator != PIMDB_PROP_CARTESIAN && c->propagator != PIMDB_PROP_NORMAL_MODES) {
        msg = "The specified time propagator is not supported!"; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (c->bosonic && c->propagator == PIMDB_PROP_NORMAL_MODES) {
        msg = "Normal modes propogation is currently not available for bosons!"; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (c->thermostat < PIMDB_THERMO_NONE || c->thermostat > PIMDB_THERMO_NOSE_HOOVER_NP_DIM) {
        msg = "The specified thermostat is not supported!"; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (c->thermostat >= PIMDB_THERMO_NOSE_HOOVER) {
        if (c->nchains < 1) {
            snprintf(buf, sizeof buf, "The specified number of Nose-Hoover chains (%d) is less than one!", c->nchains);
            msg = buf; return PIMDB_ERR_INVALID_ARGUMENT;
        }
    }
    if (c->nmthermostat && c->thermostat == PIMDB_THERMO_NONE) {
        msg = "nmthermostat cannot be used in nve ensemble!"; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (!(c->temperature > 0.0)) {
        snprintf(buf, sizeof buf, "The specified temperature (%4.3f kelvin) is unphysical!", c->temperature);
        msg = buf; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (!(c->mass > 0.0)) {
        snprintf(buf, sizeof buf, "The provided mass (%4.3f) is unphysical!", c->mass);
        msg = buf; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (!(c->size > 0.0)) {
        snprintf(buf, sizeof buf, "The provided system size (%4.3f) is unphysical!", c->size);
        msg = buf; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (!(c->dt > 0.0)) { msg = "The time step must be positive"; return PIMDB_ERR_INVALID_ARGUMENT; }
    switch (c->int_potential) {
        case PIMDB_POT_FREE: case PIMDB_POT_AZIZ: case PIMDB_POT_HARMONIC: case PIMDB_POT_DIPOLE: break;
        default: msg = "The specified interaction potential is not supported!"; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    switch (c->ext_potential) {
        case PIMDB_POT_FREE: case PIMDB_POT_HARMONIC: case PIMDB_POT_DOUBLE_WELL: case PIMDB_POT_COSINE: break;
        default: msg = "The specified external potential is not supported!"; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (c->bead_begin < 0 || c->bead_end > c->nbeads || c->bead_begin >= c->bead_end) {
        msg = "bead range [bead_begin, bead_end) must be a non-empty sub-range of [0, nbeads)";
        return PIMDB_ERR_INVALID_ARGUMENT;
    }
    const bool all_local = (c->bead_begin == 0 && c->bead_end == c->nbeads);
    if (!all_local && c->nmthermostat && (c->thermostat >= PIMDB_THERMO_NOSE_HOOVER || c->rng == PIMDB_RNG_RANMARS)) {
        // (on a bead shard every rank transforms all modes from gathered slabs; the chains / the sequential generators of the
        // modes would have to be replicated as well)
        msg = "on a bead shard the normal-mode thermostat coupling supports the Langevin thermostat with the default noise stream";
        return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (c->natoms > 65535 * kTile) { msg = "natoms too large"; return PIMDB_ERR_INVALID_ARGUMENT; }
    if (c->exchange_alg != PIMDB_EXCH_QUADRATIC && c->exchange_alg != PIMDB_EXCH_FACTORIAL) {
        msg = "unknown bosonic exchange algorithm"; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    if (c->exchange_alg == PIMDB_EXCH_FACTORIAL && c->bosonic && c->nbeads > 1 && c->natoms > 10) {
        msg = "the factorial exchange algorithm sums over all N! permutations: natoms <= 10"; return PIMDB_ERR_INVALID_ARGUMENT;
    }
    return PIMDB_OK;
}

// reference src/normal_modes.cpp:48-79 -- forward rows C[k][.] and the inverse rows each rank builds for itself
static void build_nm_tables(const Sim* s, std::vector<double>& mats, std::vector<double>& tab) {
    const int P = s->P;
    const double pi = 3.14159265358979323846;  // std::numbers::pi
    mats.assign((size_t)2 * P * P, 0.0);
    tab.assign((size_t)3 * P, 0.0);
    double* fwd = mats.data();
    double* inv = mats.data() + (size_t)P * P;
    for (int k = 0; k < P; ++k) {
        const double fund = 2 * pi / P * k;
        double* row = fwd + (size_t)k * P;
        if (k == 0) {
            for (int j = 0; j < P; ++j) row[j] = 1 / std::sqrt((double)P);
        } else if (k < 0.5 * P) {
            for (int j = 0; j < P; ++j) row[j] = std::sqrt(2.0 / P) * std::cos(fund * j);
        } else if (k == 0.5 * P) {
            for (int j = 0; j < P; ++j) row[j] = 1 / std::sqrt((double)P) * (j % 2 == 0 ? 1.0 : -1.0);
        } else {
            for (int j = 0; j < P; ++j) row[j] = -std::sqrt(2.0 / P) * std::sin(fund * j);
        }
        double* irow = inv + (size_t)k * P;   // held by the rank of bead k in the reference
        const double pref = std::sqrt(2.0 / P);
        irow[0] = 1 / std::sqrt((double)P);
        for (int i = 1; i < 0.5 * P; ++i) irow[i] = pref * std::cos(fund * i);
        if (P % 2 == 0) irow[P / 2] = 1 / std::sqrt((double)P) * (k % 2 == 0 ? 1.0 : -1.0);
        for (int i = (int)std::ceil(0.5 * (P + 1)); i < P; ++i) irow[i] = -pref * std::sin(fund * i);
        // free ring-polymer frequencies, src/propagators/normal_modes_propagator.cpp:13-16
        const double freq = 2 * s->omega_p * std::sin(k * pi / P);
        tab[k] = std::cos(freq * s->cfg.dt);
        tab[P + k] = std::sin(freq * s->cfg.dt);
        tab[2 * P + k] = s->cfg.mass * freq;
    }
}

static void free_all(Sim* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
    if (s->graph) cudaGraphDestroy(s->graph);
    if (s->graph_dl_exec) cudaGraphExecDestroy(s->graph_dl_exec);
    if (s->graph_dl) cudaGraphDestroy(s->graph_dl);
    if (s->stream_c) cudaStreamDestroy(s->stream_c);
    if (s->stream_n) cudaStreamDestroy(s->stream_n);
    if (s->ev_nz_fork) cudaEventDestroy(s->ev_nz_fork);
    if (s->ev_nz_join) cudaEventDestroy(s->ev_nz_join);
    cudaFree(s->nz); cudaFree(s->nz_tag);
    if (s->ev_dl_fork) cudaEventDestroy(s->ev_dl_fork);
    if (s->ev_dl_join) cudaEventDestroy(s->ev_dl_join);
    for (auto* v : {&s->ev_pair, &s->ev_step, &s->ev_integ})
        for (auto& e : *v) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    cudaFree(s->x); cudaFree(s->p); cudaFree(s->f); cudaFree(s->fs); cudaFree(s->fp);
    cudaFree(s->stage_d); cudaFreeHost(s->stage_h);
    cudaFree(s->tile_ij); cudaFree(s->pair_scratch);
    cudaFree(s->exA); cudaFree(s->exV); cudaFree(s->exVb); cudaFree(s->exF); cudaFree(s->exTab);
    cudaFree(s->exC); cudaFree(s->exK); cudaFree(s->exB); cudaFree(s->exG); cudaFree(s->exGok); cudaFree(s->exSync); cudaFree(s->tl); cudaFree(s->rm_state); cudaFree(s->rm_noise); cudaFree(s->exWm); cudaFree(s->exWe);
    cudaFree(s->com_part); cudaFree(s->com); cudaFree(s->tickets); cudaFree(s->draw);
    cudaFree(s->obs_d); cudaFreeHost(s->obs_h); cudaFree(s->obs_part);
    cudaFreeHost(s->err_h);
    cudaFree(s->nmC); cudaFree(s->nmFreq); cudaFree(s->nh_state);
    for (void* m : s->ipc_opened) cudaIpcCloseMemHandle(m);
    cudaFree(s->mailbox); cudaFree(s->peer_seq); cudaFree(s->stamps); cudaFree(s->fact_work);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->ev_copy) cudaEventDestroy(s->ev_copy);
    for (cudaEvent_t e : s->ev_slice) cudaEventDestroy(e);
    if (s->stream_x) cudaStreamDestroy(s->stream_x);
    if (s->stream_r) cudaStreamDestroy(s->stream_r);
    if (s->ev_join2) cudaEventDestroy(s->ev_join2);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int pimdb_abi_version(void) { return PIMDB_ABI_VERSION; }

extern "C" const char* pimdb_last_error(const pimdb_sim* sim) {
    if (!sim) return g_create_error.c_str();
    return reinterpret_cast<const Sim*>(sim)->err.c_str();
}

#define CREATE_TRY(expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            g_create_error = std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr; \
            free_all(s);                                                                         \
            return PIMDB_ERR_CUDA;                                                               \
        }                                                                                        \
    } while (0)

extern "C" int pimdb_create(const pimdb_config* cfg, pimdb_sim** out) {
    if (!out) return fail(nullptr, PIMDB_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (!cfg) return fail(nullptr, PIMDB_ERR_INVALID_ARGUMENT, "cfg is NULL");
    std::string msg;
    int rc = validate(cfg, msg);
    if (rc != PIMDB_OK) return fail(nullptr, rc, msg);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, PIMDB_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, PIMDB_ERR_CUDA, "invalid CUDA device ordinal");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess)
        return fail(nullptr, PIMDB_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, PIMDB_ERR_CUDA, "libpimdb200 is built for sm_100a (B200) only; device is sm_" +
                                                 std::to_string(prop.major) + std::to_string(prop.minor));

    Sim* s = new (std::nothrow) Sim();
    if (!s) return fail(nullptr, PIMDB_ERR_RUNTIME, "out of host memory");
    s->cfg = *cfg;
    s->device = cfg->device;
    s->sm_count = prop.multiProcessorCount;
    s->N = cfg->natoms; s->P = cfg->nbeads; s->D = cfg->ndim;
    s->b0 = cfg->bead_begin; s->b1 = cfg->bead_end; s->Ploc = s->b1 - s->b0;
    s->all_local = (s->b0 == 0 && s->b1 == s->P);
    s->has_first = (s->b0 == 0);
    s->has_last = (s->b1 == s->P);
    s->bosonic = cfg->bosonic && s->P > 1;                       // src/simulation.cpp:690
    s->factorial = s->bosonic && cfg->exchange_alg == PIMDB_EXCH_FACTORIAL;
    s->S = (size_t)s->D * s->N;
    // src/simulation.cpp:37-52 (i-PI convention)
    s->beta = 1.0 / cfg->temperature;
    s->thermo_beta = s->beta / s->P;
    s->exch_beta = s->beta / s->P;                               // bosonic_exchange_base.cpp:16-18
    s->omega_p = s->P / s->beta;
    s->kspring = cfg->mass * s->omega_p * s->omega_p;
    s->L = cfg->size;
    // src/simulation.cpp:84-90
    s->rc = (cfg->int_potential == PIMDB_POT_FREE) ? 0.0 : cfg->cutoff;
    if (cfg->pbc) s->rc = std::fmin(s->rc, 0.5 * cfg->size);
    s->pair_on = (s->rc != 0.0) && cfg->int_potential != PIMDB_POT_FREE;
    s->kext = (cfg->ext_potential == PIMDB_POT_HARMONIC) ? cfg->mass * cfg->ext_omega * cfg->ext_omega : 0.0;
    s->pair_par = (cfg->int_potential == PIMDB_POT_HARMONIC) ? cfg->mass * cfg->int_omega * cfg->int_omega
                                                              : cfg->int_strength;
    // src/thermostats/langevin.cpp:10-13
    s->c1 = std::exp(-0.5 * cfg->gamma * cfg->dt);
    s->c2 = std::sqrt((1 - s->c1 * s->c1) * cfg->mass / s->thermo_beta);
    s->T = (s->N + kTile - 1) / kTile;
    s->TP = s->T * (s->T + 1) / 2;

    CREATE_TRY(cudaSetDevice(s->device));
    int lo = 0, hi = 0;
    CREATE_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    s->prio_hi = hi; s->prio_lo = lo;
    CREATE_TRY(cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, lo));
    CREATE_TRY(cudaStreamCreateWithPriority(&s->stream_x, cudaStreamNonBlocking, hi));
    CREATE_TRY(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
    CREATE_TRY(cudaStreamCreateWithPriority(&s->stream_r, cudaStreamNonBlocking, hi));
    CREATE_TRY(cudaStreamCreateWithPriority(&s->stream_c, cudaStreamNonBlocking, hi));
    CREATE_TRY(cudaEventCreateWithFlags(&s->ev_dl_fork, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&s->ev_dl_join, cudaEventDisableTiming));
    CREATE_TRY(cudaEventCreateWithFlags(&s->ev_join2, cudaEventDisableTiming));

    const size_t slab_bytes = s->S * sizeof(double);
    const size_t own_bytes = slab_bytes * s->Ploc;
    CREATE_TRY(cudaMalloc(&s->x, slab_bytes * (s->Ploc + 2)));
    CREATE_TRY(cudaMalloc(&s->p, own_bytes));
    CREATE_TRY(cudaMalloc(&s->f, own_bytes));
    CREATE_TRY(cudaMalloc(&s->fs, own_bytes));
    CREATE_TRY(cudaMalloc(&s->fp, own_bytes));
    CREATE_TRY(cudaMalloc(&s->stage_d, 3 * own_bytes));
    CREATE_TRY(cudaHostAlloc(&s->stage_h, 3 * own_bytes, cudaHostAllocDefault));
    CREATE_TRY(cudaEventCreateWithFlags(&s->ev_copy, cudaEventDisableTiming));
    CREATE_TRY(cudaMemset(s->x, 0, slab_bytes * (s->Ploc + 2)));
    CREATE_TRY(cudaMemset(s->p, 0, own_bytes));
    CREATE_TRY(cudaMemset(s->f, 0, own_bytes));    // forces start at zero like the reference's (App. A-1)
    CREATE_TRY(cudaMemset(s->fs, 0, own_bytes));
    CREATE_TRY(cudaMemset(s->fp, 0, own_bytes));

    if (s->pair_on) {
        std::vector<ushort2> tiles;
        tiles.reserve(s->TP);
        for (int I = 0; I < s->T; ++I)       // off-diagonal tile pairs first, the diagonal ones (half the work) last: see k_pair_tiles
            for (int J = I + 1; J < s->T; ++J) tiles.push_back(make_ushort2((unsigned short)I, (unsigned short)J));
        for (int I = 0; I < s->T; ++I) tiles.push_back(make_ushort2((unsigned short)I, (unsigned short)I));
        CREATE_TRY(cudaMalloc(&s->tile_ij, tiles.size() * sizeof(ushort2)));
        CREATE_TRY(cudaMemcpy(s->tile_ij, tiles.data(), tiles.size() * sizeof(ushort2), cudaMemcpyHostToDevice));
        const size_t per_bead = (size_t)s->T * s->T * s->D * kTile * sizeof(double);
        size_t cap_mb = 4096;
        if (const char* e = getenv("PIMDB_PAIR_SCRATCH_MB")) cap_mb = (size_t)std::max(1, atoi(e));
        size_t chunk = (cap_mb << 20) / per_bead;
        if (chunk < 1) chunk = 1;
        if (chunk > (size_t)s->Ploc) chunk = s->Ploc;
        s->bead_chunk = (int)chunk;
        CREATE_TRY(cudaMalloc(&s->pair_scratch, per_bead * chunk));
    }
    if (s->bosonic) {
        const size_t NN = (size_t)s->N * s->N;
        CREATE_TRY(cudaMalloc(&s->exA, sizeof(double) * (2 * s->N + 2)));   // A[N] | Inv[N+1]
        // Up to N = 8192 the Boltzmann factors exist only as block-scaled tiles (8 bytes per entry) + diagonal-block inverses; the
        // 16-byte extended-range tables (2 N^2 entries, 2.1 GB at N = 8192) are allocated for larger N and for the scalar
        // cross-check recurrences (PIMDB_EXCH_NOBLOCKED=1) only
        const bool scalar_path = !s->factorial && (s->N > 8192 || getenv("PIMDB_EXCH_NOBLOCKED"));
        if (s->factorial) CREATE_TRY(cudaMalloc(&s->fact_work, sizeof(double) * (5 * NN + s->N)));
        if (scalar_path || (!s->factorial && s->N <= 512)) CREATE_TRY(cudaMalloc(&s->exC, sizeof(int4) * 2 * NN));   // (N <= 512: kept beside the tiles)
        if (!scalar_path && !s->factorial) {  // block-scaled factor tiles + diagonal-block inverses of the blocked recurrence (exchange.cu)
            const size_t nbk = (size_t)((s->N + 31) / 32);
            CREATE_TRY(cudaMalloc(&s->exK, sizeof(double) * 2 * nbk * nbk * 1024));      // 32 x 32 tiles, forward | backward
            CREATE_TRY(cudaMemset(s->exK, 0, sizeof(double) * 2 * nbk * nbk * 1024));
            CREATE_TRY(cudaMalloc(&s->exB, sizeof(int) * 2 * nbk * s->N));
            CREATE_TRY(cudaMalloc(&s->exG, sizeof(double) * 2 * nbk * (1024 + 32)));     // G f|b, h f|b
            CREATE_TRY(cudaMemset(s->exG, 0, sizeof(double) * 2 * nbk * (1024 + 32)));
            CREATE_TRY(cudaMalloc(&s->exGok, sizeof(int) * 4 * (size_t)((s->N + 31) / 32)));   // Gok f|b, block status f|b
            CREATE_TRY(cudaMemset(s->exGok, 0, sizeof(int) * 4 * (size_t)((s->N + 31) / 32)));
        }
        if (getenv("PIMDB_TIMELINE")) {
            CREATE_TRY(cudaMalloc(&s->tl, sizeof(unsigned long long) * 64));
            std::vector<unsigned long long> init(64);
            for (int i = 0; i < 32; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0ull; }
            CREATE_TRY(cudaMemcpy(s->tl, init.data(), sizeof(unsigned long long) * 64, cudaMemcpyHostToDevice));
        }
        CREATE_TRY(cudaMalloc(&s->exSync, sizeof(int) * 4));
        CREATE_TRY(cudaMemset(s->exSync, 0, sizeof(int) * 4));
        CREATE_TRY(cudaMalloc(&s->exWm, sizeof(double) * 2 * (s->N + 1)));
        CREATE_TRY(cudaMalloc(&s->exWe, sizeof(int) * 2 * (s->N + 1)));
        CREATE_TRY(cudaMemset(s->exWm, 0, sizeof(double) * 2 * (s->N + 1)));
        CREATE_TRY(cudaMemset(s->exWe, 0, sizeof(int) * 2 * (s->N + 1)));
        CREATE_TRY(cudaMalloc(&s->exV, sizeof(double) * (s->N + 1)));
        CREATE_TRY(cudaMalloc(&s->exVb, sizeof(double) * (s->N + 1)));
        CREATE_TRY(cudaMalloc(&s->exF, sizeof(double) * 2 * s->S));
        CREATE_TRY(cudaMemset(s->exV, 0, sizeof(double) * (s->N + 1)));
        CREATE_TRY(cudaMemset(s->exVb, 0, sizeof(double) * (s->N + 1)));
        CREATE_TRY(cudaMemset(s->exF, 0, sizeof(double) * 2 * s->S));
    }
    if (cfg->rng == PIMDB_RNG_RANMARS && cfg->thermostat == PIMDB_THERMO_LANGEVIN) {
        const int rc = ranmars_create(s);
        if (rc != PIMDB_OK) { g_create_error = s->err; free_all(s); return rc; }
    }
    if (!s->all_local) {   // mailbox + counters of the peer-memory sharding protocol (pimdb_peer_export / _attach)
        size_t mb_bytes = sizeof(PeerMailbox) + sizeof(unsigned long long) * 2 * 2 * 2 * s->S;   // + the self-validating halo inbox
        if (cfg->propagator == PIMDB_PROP_NORMAL_MODES || cfg->nmthermostat)
            mb_bytes += sizeof(double) * 2 * 2 * (size_t)s->P * s->S;                            // + the all-gather buffer
        CREATE_TRY(cudaMalloc(&s->mailbox, mb_bytes));
        CREATE_TRY(cudaMemset(s->mailbox, 0, mb_bytes));
        CREATE_TRY(cudaMalloc(&s->peer_seq, sizeof(unsigned int) * 8));
        CREATE_TRY(cudaMemset(s->peer_seq, 0, sizeof(unsigned int) * 8));
    }
    if (getenv("PIMDB_TIMELINE")) {
        CREATE_TRY(cudaMalloc(&s->stamps, sizeof(unsigned long long) * 64));
        CREATE_TRY(cudaMemset(s->stamps, 0, sizeof(unsigned long long) * 64));
    }
    CREATE_TRY(cudaMalloc(&s->com_part, sizeof(double) * 2 * 4 * kMaxPartials));
    CREATE_TRY(cudaMemset(s->com_part, 0, sizeof(double) * 2 * 4 * kMaxPartials));
    CREATE_TRY(cudaMalloc(&s->com, sizeof(double) * 4));
    CREATE_TRY(cudaMemset(s->com, 0, sizeof(double) * 4));
    CREATE_TRY(cudaMalloc(&s->tickets, sizeof(unsigned int) * 4));
    CREATE_TRY(cudaMemset(s->tickets, 0, sizeof(unsigned int) * 4));
    CREATE_TRY(cudaMalloc(&s->draw, sizeof(unsigned long long)));
    CREATE_TRY(cudaMemset(s->draw, 0, sizeof(unsigned long long)));
    if (cfg->thermostat == PIMDB_THERMO_LANGEVIN && !cfg->nmthermostat && cfg->propagator == PIMDB_PROP_CARTESIAN &&
        cfg->rng != PIMDB_RNG_RANMARS && !getenv("PIMDB_NO_NOISE_PREFETCH") &&
        (size_t)s->Ploc * s->D * s->N <= (size_t)(getenv("PIMDB_NOISE_PREFETCH_MAX") ? atol(getenv("PIMDB_NOISE_PREFETCH_MAX")) : 262144)) {
        // (small systems only, where the thermostat launches are latency-bound: measured on B200, C3 -- 98 k momenta -- gains
        // 1.5-2 us of a 64 us iteration; at C4 -- 786 k -- it is neutral and at C5 -- 6.3 M -- the extra 200 MB of traffic per
        // iteration beside the HBM-bound exchange kernels costs 2.5 %)
        const size_t slot = (size_t)s->Ploc * s->D * s->N * sizeof(double);
        CREATE_TRY(cudaMalloc(&s->nz, 2 * slot));
        CREATE_TRY(cudaMalloc(&s->nz_tag, 2 * sizeof(unsigned long long)));
        CREATE_TRY(cudaMemset(s->nz_tag, 0xff, 2 * sizeof(unsigned long long)));     // no draw has this index
        CREATE_TRY(cudaStreamCreateWithPriority(&s->stream_n, cudaStreamNonBlocking, s->prio_lo));
        CREATE_TRY(cudaEventCreateWithFlags(&s->ev_nz_fork, cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&s->ev_nz_join, cudaEventDisableTiming));
        s->nz_on = true;
    }
    s->no_ticketless = getenv("PIMDB_NO_TICKETLESS") != nullptr;
    CREATE_TRY(cudaMalloc(&s->obs_d, sizeof(DevObs)));
    CREATE_TRY(cudaMemset(s->obs_d, 0, sizeof(DevObs)));
    CREATE_TRY(cudaHostAlloc(&s->obs_h, sizeof(DevObs), cudaHostAllocDefault));
    CREATE_TRY(cudaMalloc(&s->obs_part, sizeof(double) * 8 * kMaxPartials));
    CREATE_TRY(cudaHostAlloc(&s->err_h, sizeof(int), cudaHostAllocMapped));
    *s->err_h = 0;
    CREATE_TRY(cudaHostGetDevicePointer(&s->err_d, s->err_h, 0));
    if (cfg->propagator == PIMDB_PROP_NORMAL_MODES || cfg->nmthermostat) {
        std::vector<double> mats, tab;
        build_nm_tables(s, mats, tab);
        CREATE_TRY(cudaMalloc(&s->nmC, mats.size() * sizeof(double)));
        CREATE_TRY(cudaMalloc(&s->nmFreq, tab.size() * sizeof(double)));
        CREATE_TRY(cudaMemcpy(s->nmC, mats.data(), mats.size() * sizeof(double), cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMemcpy(s->nmFreq, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (cfg->thermostat >= PIMDB_THERMO_NOSE_HOOVER) {
        const size_t groups = cfg->thermostat == PIMDB_THERMO_NOSE_HOOVER ? 1
                              : cfg->thermostat == PIMDB_THERMO_NOSE_HOOVER_NP ? (size_t)s->N : (size_t)s->N * s->D;
        s->nh_len = (size_t)s->Ploc * groups * cfg->nchains;
        CREATE_TRY(cudaMalloc(&s->nh_state, 3 * s->nh_len * sizeof(double)));
        CREATE_TRY(cudaMemset(s->nh_state, 0, 3 * s->nh_len * sizeof(double)));   // eta = eta_dot = eta_dot_dot = 0 (nose_hoover.cpp:22-24)
    }
    CREATE_TRY(cudaDeviceSynchronize());
    *out = reinterpret_cast<pimdb_sim*>(s);
    return PIMDB_OK;
}

extern "C" void pimdb_destroy(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    cudaStreamSynchronize(s->stream_x);
    cudaStreamSynchronize(s->stream_r);
    free_all(s);
}

// ----------------------------------------------------------------------------------------------------
static int check_deferred(Sim* s) {
    if (*s->err_h != 0) {
        const int e = *s->err_h;
        *s->err_h = e & kErrPeerTimeout;   // a lost peer is permanent: every later synchronising call reports it again
        if (e & kErrPeerTimeout)
            return fail(s, PIMDB_ERR_RUNTIME, "bead shard timed out waiting for a peer GPU (halo slice / momentum sums); every rank must "
                                              "make the same sequence of calls");
        if (e & kErrSyncTimeout)   // the bounded waits between the blocks of the cluster recurrence (never seen; an internal error)
            return fail(s, PIMDB_ERR_RUNTIME, "exchange recurrence: a wait between the blocks of its thread-block cluster ran out");
        // same wording as the reference's std::overflow_error (quadratic_bosonic_exchange.cpp:92-97,119-124)
        return fail(s, PIMDB_ERR_OVERFLOW,
                    std::string("Invalid sig_denom / e_shift in bosonic exchange potential (non-finite ") +
                        ((e & kErrOverflowFwd) ? "V" : "V_backwards") + ")");
    }
    return PIMDB_OK;
}

extern "C" int pimdb_synchronize(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    PIMDB_CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    return check_deferred(s);
}

static double* array_ptr(Sim* s, int which, bool& halo) {
    halo = false;
    switch (which) {
        case PIMDB_X: halo = true; return s->x;
        case PIMDB_P: return s->p;
        case PIMDB_F: return s->f;
        case PIMDB_F_SPRING: return s->fs;
        case PIMDB_F_PHYS: return s->fp;
        default: return nullptr;
    }
}

// true when the caller's buffer is page-locked (cudaMallocHost / cudaHostRegister / torch pin_memory)
static bool host_is_pinned(const void* ptr) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// f_spring / f_phys after a step that assembled the forces inside its closing integrator kernel: rebuilt on request from
// the pair partials and exterior forces of that evaluation, which are still in place
static int refresh_split_forces(Sim* s) {
    if (!s->split_stale) return PIMDB_OK;
    return launch_assemble_chunk(s, 0, s->Ploc, s->pair_on);
}

// One upload = one asynchronous PCIe copy per array (straight from a page-locked caller buffer, through the pinned
// staging buffer otherwise) + ONE transpose kernel for all of them.
// `wait_for_copy`: return only when the PCIe copies out of the caller's buffers are done (pimdb_set_state: the caller may
// reuse its buffer at once). pimdb_upload_state does not wait for page-locked buffers -- the usual contract of an
// asynchronous copy, stated in the header -- so that the host can enqueue the step behind the upload without a bubble.
static int upload_arrays(Sim* s, int n, const int* which, const double* const* host, bool wait_for_copy = true) {
    const size_t count = s->S * s->Ploc, bytes = count * sizeof(double);
    double* dst[3]; bool halo[3];
    if (n < 1 || n > 3) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "1 to 3 arrays per transfer");
    bool all_pinned = true;
    for (int i = 0; i < n; ++i) {
        dst[i] = array_ptr(s, which[i], halo[i]);
        if (!dst[i] || !host[i]) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "unknown state array");
        all_pinned = all_pinned && host_is_pinned(host[i]);
        if (which[i] == PIMDB_P) { s->p_shift_pending = false; s->z_owed = false; }   // the caller replaces the momenta
    }
    if (!all_pinned) PIMDB_CUDA_TRY(s, cudaStreamSynchronize(s->stream));   // the pinned staging buffer is about to be rewritten
    bool packed = true;                                                      // arrays adjacent in host memory: ONE copy
    for (int i = 1; i < n; ++i) packed = packed && host[i] == host[i - 1] + count;
    if (!all_pinned) {
        for (int i = 0; i < n; ++i) memcpy(s->stage_h + i * count, host[i], bytes);
        PIMDB_CUDA_TRY(s, cudaMemcpyAsync(s->stage_d, s->stage_h, n * bytes, cudaMemcpyHostToDevice, s->stream));
    } else if (packed) {
        PIMDB_CUDA_TRY(s, cudaMemcpyAsync(s->stage_d, host[0], n * bytes, cudaMemcpyHostToDevice, s->stream));
    } else {
        for (int i = 0; i < n; ++i)
            PIMDB_CUDA_TRY(s, cudaMemcpyAsync(s->stage_d + i * count, host[i], bytes, cudaMemcpyHostToDevice, s->stream));
    }
    // The caller may reuse its buffers as soon as the call returns (pageable and page-locked alike): wait for the copies,
    // not for the transpose.
    PIMDB_CUDA_TRY(s, cudaEventRecord(s->ev_copy, s->stream));
    bool x_changed = false, p_changed = false;
    for (int i = 0; i < n; ++i) { x_changed = x_changed || which[i] == PIMDB_X; p_changed = p_changed || which[i] == PIMDB_P; }
    // (a full ring: the transpose fills its halo slabs too; new momenta on a handle that owns every bead: it also leaves zero
    // momentum sums pending -- p - 0.0 == p bit for bit -- which is the state a captured iteration starts from, see
    // make_entry_state_uniform: no separate memset in front of the next pimdb_step)
    const bool zero_com = p_changed && s->cfg.fixcom && s->all_local && !s->peer_on;
    API_TRY(launch_aos_to_soa(s, n, dst, halo, x_changed && s->all_local, zero_com));
    if (zero_com) { s->p_shift_pending = true; s->com_buf = 0; }
    if (x_changed && s->peer_on) API_TRY(launch_peer_push_halos(s));
    if (wait_for_copy || !all_pinned) PIMDB_CUDA_TRY(s, cudaEventSynchronize(s->ev_copy));
    return PIMDB_OK;
}

static int download_arrays(Sim* s, int n, const int* which, double* const* host) {
    const size_t count = s->S * s->Ploc, bytes = count * sizeof(double);
    const double* src[3]; bool halo[3];
    if (n < 1 || n > 3) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "1 to 3 arrays per transfer");
    bool all_pinned = true;
    for (int i = 0; i < n; ++i) {
        src[i] = array_ptr(s, which[i], halo[i]);
        if (!src[i] || !host[i]) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "unknown state array");
        all_pinned = all_pinned && host_is_pinned(host[i]);
        if (which[i] == PIMDB_P) API_TRY(settle_momenta(s));
        if (which[i] == PIMDB_F_SPRING || which[i] == PIMDB_F_PHYS) API_TRY(refresh_split_forces(s));
    }
    API_TRY(launch_soa_to_aos(s, n, src, halo));
    bool packed = true;
    for (int i = 1; i < n; ++i) packed = packed && host[i] == host[i - 1] + count;
    if (!all_pinned || packed) {
        PIMDB_CUDA_TRY(s, cudaMemcpyAsync(all_pinned ? host[0] : s->stage_h, s->stage_d, n * bytes, cudaMemcpyDeviceToHost, s->stream));
    } else {
        for (int i = 0; i < n; ++i)
            PIMDB_CUDA_TRY(s, cudaMemcpyAsync(host[i], s->stage_d + i * count, bytes, cudaMemcpyDeviceToHost, s->stream));
    }
    PIMDB_CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    if (!all_pinned)
        for (int i = 0; i < n; ++i) memcpy(host[i], s->stage_h + i * count, bytes);
    return check_deferred(s);
}

extern "C" int pimdb_set_state(pimdb_sim* sim, int which, const double* host) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !host) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    return upload_arrays(s, 1, &which, &host);
}

extern "C" int pimdb_get_state(pimdb_sim* sim, int which, double* host) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !host) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    return download_arrays(s, 1, &which, &host);
}

extern "C" int pimdb_upload_state(pimdb_sim* sim, const double* x, const double* p) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || (!x && !p)) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    int which[2]; const double* host[2]; int n = 0;
    if (x) { which[n] = PIMDB_X; host[n++] = x; }
    if (p) { which[n] = PIMDB_P; host[n++] = p; }
    return upload_arrays(s, n, which, host, false);
}

extern "C" int pimdb_download_state(pimdb_sim* sim, double* x, double* p, double* f) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || (!x && !p && !f)) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    int which[3]; double* host[3]; int n = 0;
    if (x) { which[n] = PIMDB_X; host[n++] = x; }
    if (p) { which[n] = PIMDB_P; host[n++] = p; }
    if (f) { which[n] = PIMDB_F; host[n++] = f; }
    return download_arrays(s, n, which, host);
}

// ----------------------------------------------------------------------------------------------------
// force evaluation: exchange on the high-priority side stream, pair tiles + assembly on the main stream
// `assemble_later`: the caller's next k_integrate launch forms f itself (OP_ASSEMBLE) -- one pass over the pair partials
// and one launch fewer on the step's critical path.
// `after_integrate`: the previous launch on the main stream was one of our integrator kernels (a captured step): the first
// kernel of the exchange chain may be launched early behind it.
static int enqueue_forces_inner(Sim* s, bool assemble_later, bool after_integrate) {
    const bool ex = s->bosonic && (s->has_first || s->has_last);
    bool pair_early = false;
    if (ex) {
        // The exchange chain is: factor tiles + block inverses -> recurrences (a few large blocks, latency-bound) -> exterior
        // forces, beside the pair tiles. What matters is the ORDER in which the grids get their SMs: the recurrence blocks
        // must be resident before the pair tiles flood the GPU, or they wait for pair-tile blocks to retire (one wave is
        // ~19 us at C3). Round 1 launched the recurrence first and let it spin on a device counter for the tiles; now every
        // dependency is real. In a captured step all three grids sit on the main stream, chained by programmatic dependent
        // launches (PTX griddepcontrol):
        //     tiles  ->  recurrences (scheduled once every tile block has started; block in griddepcontrol.wait until the
        //                tile grid has completed)  ->  pair tiles (scheduled once every recurrence block has started; they
        //                read nothing the exchange kernels write, so they never wait)
        // and the exterior forces follow the recurrences on the high-priority side stream. Nothing can time out, and a
        // profiler / MPS / time slicing that serialises the kernels only removes the overlap.
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(s->stream, &cap);
        static const bool no_chain = getenv("PIMDB_EXCH_NOCHAIN") != nullptr;     // plain order (A/B timing)
        // (only while the tile grid is a single wave, N <= 512: the pair tiles are scheduled once every tile block and then every
        // recurrence block has STARTED, and a tile grid of many waves would hold them back for most of its run time)
        const bool chain = cap == cudaStreamCaptureStatusActive && s->exK && s->N <= 512 && !no_chain;
        if (!chain) {
            // Tile grids of many waves (N > 512), or eager launches: factor tiles on the main stream AHEAD of the pair tiles, the
            // rest of the chain on the side stream. (A 1024-thread tile block needs more registers than one retiring pair-tile
            // block frees, so behind a running pair-tile grid it is never placed, whatever its priority: launched after the
            // pair tiles, the tile grid of C4 started when the pair tiles had finished -- measured.)
            API_TRY(launch_exchange_part(s, s->stream, 0));
            PIMDB_CUDA_TRY(s, cudaEventRecord(s->ev_fork, s->stream));
            PIMDB_CUDA_TRY(s, cudaStreamWaitEvent(s->stream_x, s->ev_fork, 0));
            API_TRY(launch_exchange_part(s, s->stream_x, 1));
        } else {
        if (after_integrate) allow_early_launch(s);
        const int rc0 = launch_exchange_part(s, s->stream, 0);
        s->pdl_next = false;
        API_TRY(rc0);
        {
            s->pdl_recur = true;
            const int rc2 = launch_exchange_part(s, s->stream, 2);
            s->pdl_recur = false;
            API_TRY(rc2);
            PIMDB_CUDA_TRY(s, cudaEventRecord(s->ev_fork, s->stream));
            PIMDB_CUDA_TRY(s, cudaStreamWaitEvent(s->stream_x, s->ev_fork, 0));
            API_TRY(launch_exchange_part(s, s->stream_x, 3));
            pair_early = true;
        }
        }
        PIMDB_CUDA_TRY(s, cudaEventRecord(s->ev_join, s->stream_x));
    }
    bool joined = !ex;
    auto join = [&]() -> int {
        PIMDB_CUDA_TRY(s, cudaStreamWaitEvent(s->stream, s->ev_join, 0));
        joined = true;
        return PIMDB_OK;
    };
    if (s->pair_on) {
        // A captured step cuts a LARGE pair-tile grid into several launches of >= ~4 waves each, chained by programmatic
        // launches without a wait (they are independent): the next launch is scheduled as soon as every block of the
        // previous one has started, so the SMs never drain in between -- but each boundary is a point where the hardware
        // picks among the pending grids again, and the exterior-force kernel of the exchange chain (higher priority, ready
        // long before the pair tiles are through) gets its blocks placed there instead of behind the whole pair-tile
        // queue (measured at C4: it started when the last pair-tile block had been dispatched, 880 us after it was ready).
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(s->stream, &cap);
        const bool capturing = cap == cudaStreamCaptureStatusActive;
        // (slices of ~6 waves -- cutting costs ~4 % of the pair tiles through the joints -- but at least two once the grid is
        // 4 waves long, so that a bead shard of C4 on 8 GPUs has a joint as well; C3, 2.5 waves, stays one launch)
        const double waves = (double)s->Ploc * s->TP / ((double)pair_resident_warps_per_sm() * s->sm_count);
        int disp = s->bead_chunk;
        if (capturing && ex && waves >= 4.0 && !getenv("PIMDB_PAIR_ONE_LAUNCH")) {
            const int nslices = std::max(2, std::min(16, (int)(waves / 6.0)));
            disp = std::max(1, (s->Ploc + nslices - 1) / nslices);
        }
        for (int lo = 0; lo < s->Ploc; lo += s->bead_chunk) {
            const int nb = std::min(s->bead_chunk, s->Ploc - lo);
            int nslice = 0;
            for (int d0 = 0; d0 < nb;) {
                int nd = std::min(disp, nb - d0);
                if (nb - (d0 + nd) < (disp + 1) / 2) nd = nb - d0;          // no small last slice: it joins the one before
                const bool early = (lo == 0 && d0 == 0) ? pair_early : (capturing && d0 > 0);
                API_TRY(launch_pair_chunk(s, lo + d0, nd, false, early, d0));
                d0 += nd;
                if (d0 < nb) {
                    // a programmatic edge orders the next slice behind the START of this one only: what comes after the last
                    // slice gets a full dependency on every slice
                    if ((int)s->ev_slice.size() <= nslice) {
                        cudaEvent_t e;
                        PIMDB_CUDA_TRY(s, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                        s->ev_slice.push_back(e);
                    }
                    PIMDB_CUDA_TRY(s, cudaEventRecord(s->ev_slice[nslice++], s->stream));
                }
            }
            for (int k = 0; k < nslice; ++k) PIMDB_CUDA_TRY(s, cudaStreamWaitEvent(s->stream, s->ev_slice[k], 0));
            if (!joined) API_TRY(join());
            if (!assemble_later) API_TRY(launch_assemble_chunk(s, lo, nb, true));
        }
    } else {
        if (!joined) API_TRY(join());
        if (!assemble_later) API_TRY(launch_assemble_chunk(s, 0, s->Ploc, false));
    }
    return PIMDB_OK;
}

// Inside an iteration with the counter-based Langevin thermostat, the noise of the next two thermostat half steps (the one
// that closes this iteration and the one that opens the next) is drawn beside the force kernels, on its own low-priority
// stream, launched last so that it takes whatever room the other grids leave (integrator.cu k_noise_prefetch).
static int enqueue_forces(Sim* s, bool assemble_later = false, bool after_integrate = false) {
    const bool prefetch = s->nz_on && s->nz_want;
    if (prefetch) PIMDB_CUDA_TRY(s, cudaEventRecord(s->ev_nz_fork, s->stream));
    API_TRY(enqueue_forces_inner(s, assemble_later, after_integrate));
    if (prefetch) {
        PIMDB_CUDA_TRY(s, cudaStreamWaitEvent(s->stream_n, s->ev_nz_fork, 0));
        API_TRY(launch_noise_prefetch(s, s->stream_n, 2, s->nz_first_off));
        PIMDB_CUDA_TRY(s, cudaEventRecord(s->ev_nz_join, s->stream_n));
        PIMDB_CUDA_TRY(s, cudaStreamWaitEvent(s->stream, s->ev_nz_join, 0));
    }
    return PIMDB_OK;
}

// The closing kick of a Cartesian step can assemble the forces itself when the pair partials of all owned beads are in
// the scratch slab at once. (PIMDB_NO_FUSED_ASSEMBLE=1: separate k_assemble, for A/B timing.)
static bool fuse_assembly(const Sim* s) {
    static const bool off = getenv("PIMDB_NO_FUSED_ASSEMBLE") != nullptr;
    return !off && s->cfg.propagator == PIMDB_PROP_CARTESIAN && (!s->pair_on || s->bead_chunk >= s->Ploc);
}

// Bead shards over peer memory: one iteration is O Z B A forces B O Z (O thermostat half step, Z zeroMomentum). Z is the
// projection p -> p - mean(p) and a Langevin O is affine with the same coefficients for every degree of freedom, so
// Z O Z = Z O (and Z Z = Z without a thermostat): the closing Z of an iteration is subsumed by the first Z of the next,
// and between iterations its exchange of momentum sums is skipped. It is carried out before anything reads the momenta
// (settle_momenta). Other thermostats do not commute with Z like that and keep both.
static bool lazy_closing_com(const Sim* s) {
    return s->peer_on && s->cfg.fixcom && !s->cfg.nmthermostat &&
           (s->cfg.thermostat == PIMDB_THERMO_LANGEVIN || s->cfg.thermostat == PIMDB_THERMO_NONE);
}

// Fuses consecutive element-wise stages into as few k_integrate launches as their fixed in-kernel order
// (SUBCM -> O_PRE -> B -> O_POST -> A, SUM last) allows.
struct Fuser {
    Sim* s;
    unsigned ops = 0;
    int stage = -1;
    int rc = PIMDB_OK;
    explicit Fuser(Sim* sim) : s(sim) {}
    bool chained = false;      // the previous launch on the main stream was made by this Fuser (nothing in between)
    void flush() {
        if (ops && rc == PIMDB_OK) {
            if (chained) allow_early_launch(s);
            rc = launch_integrate(s, ops);
            chained = true;
        }
        ops = 0;
        stage = -1;
    }
    void put(unsigned op, int st) {
        if (st <= stage) flush();
        ops |= op;
        stage = st;
    }
    void subcm() { put(OP_SUBCM, 0); }
    void langevin() {
        if (ops & (OP_O_PRE | OP_O_POST | OP_A | OP_SUM)) flush();
        if (ops & (OP_B | OP_B_PHYS)) put(OP_O_POST, 3);
        else put(OP_O_PRE, 1);
    }
    void kick(bool phys_only, bool assemble = false) {
        if (assemble || (ops & (OP_B | OP_B_PHYS))) flush();
        put((phys_only ? OP_B_PHYS : OP_B) | (assemble ? OP_ASSEMBLE : 0u), 2);
    }
    void drift() { put(OP_A | ((s->all_local || s->peer_on) ? OP_HALO : 0u), 4); }
    void sum() { put(OP_SUM, 5); }
};

static void thermostat_into(Sim* s, Fuser& fz) {
    if (s->cfg.thermostat >= PIMDB_THERMO_NOSE_HOOVER) {
        // NMCoupling (src/thermostats/thermostat_coupling.cpp:29-47): the chains of "bead" k act on the momenta of
        // normal mode k -- transform, run the same chain kernels on the mode momenta, transform back
        fz.flush();
        if (s->cfg.nmthermostat && fz.rc == PIMDB_OK) fz.rc = launch_nm_momenta(s, true);
        if (fz.rc == PIMDB_OK) fz.rc = launch_nose_hoover(s);
        if (s->cfg.nmthermostat && fz.rc == PIMDB_OK) fz.rc = launch_nm_momenta(s, false);
        fz.chained = false;
        return;
    }
    if (s->cfg.thermostat != PIMDB_THERMO_LANGEVIN) return;
    if (s->cfg.nmthermostat) {
        fz.flush();
        if (fz.rc == PIMDB_OK) fz.rc = launch_nm_thermostat(s);
        fz.chained = false;
    } else {
        fz.langevin();
    }
}

static void propagator_into(Sim* s, Fuser& fz) {
    if (s->cfg.propagator == PIMDB_PROP_CARTESIAN) {
        fz.kick(false);
        fz.drift();
        fz.flush();
        if (!s->all_local && !s->peer_on) return;   // host-driven sharding: the host exchanges halos, then calls phase 2
        if (fz.rc == PIMDB_OK) fz.rc = maybe_download_x(s);
        if (s->dl_forked) fz.chained = false;       // (an event record sits between the integrator kernel and the tile kernel)
        const bool fuse = fuse_assembly(s);
        if (fz.rc == PIMDB_OK) fz.rc = enqueue_forces(s, fuse, fz.chained);
        fz.chained = false;
        fz.kick(false, fuse);
    } else {
        fz.flush();
        fz.chained = false;
        if (fz.rc == PIMDB_OK) fz.rc = launch_nm_propagate(s);   // half kick (physical forces) + exact ring rotation
        if (fz.rc == PIMDB_OK) fz.rc = s->peer_on ? launch_peer_push_halos(s) : launch_fill_halos(s);
        if (fz.rc == PIMDB_OK) fz.rc = maybe_download_x(s);
        if (fz.rc == PIMDB_OK) fz.rc = enqueue_forces(s);
        fz.kick(true);
    }
}

// A centre-of-mass shift whose sums are known (com[]) but which has not been subtracted from p yet. pimdb_step leaves
// the second zeroMomentum of an iteration in this state so that the subtraction rides in the first kernel of the next
// iteration instead of costing a launch; every other entry point settles it first, so callers never see it.
static int settle_momenta(Sim* s) {
    if (s->z_owed) {   // peer mode: the closing zeroMomentum that consecutive iterations skip (collective: every rank gets here)
        s->z_owed = false;
        API_TRY(launch_integrate(s, OP_SUM));
        s->p_shift_pending = true;
    }
    if (!s->p_shift_pending) return PIMDB_OK;
    s->p_shift_pending = false;
    return launch_integrate(s, OP_SUBCM);
}

// body of Simulation::run, src/simulation.cpp:246-259
// Inside a captured step, a kernel that directly follows another kernel of ours on the main stream can be launched with
// programmatic stream serialisation (it calls griddepcontrol.wait first thing): its launch latency, ~1 us per kernel
// boundary, overlaps the tail of its predecessor. Eager launches keep the plain order.
static void allow_early_launch(Sim* s) {
    static const bool off = getenv("PIMDB_NO_PDL") != nullptr;   // A/B timing
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s->stream, &cap);
    s->pdl_next = !off && cap == cudaStreamCaptureStatusActive;
}

// pimdb_step_download: the coordinates are final once the drift has run, long before the forces are; their transpose and
// their PCIe copy go to a side stream here and overlap the force evaluation. Joined at the end of the iteration.
static int maybe_download_x(Sim* s) {
    if (!s->dl_hook || s->dl_forked) return PIMDB_OK;
    PIMDB_CUDA_TRY(s, cudaEventRecord(s->ev_dl_fork, s->stream));
    PIMDB_CUDA_TRY(s, cudaStreamWaitEvent(s->stream_c, s->ev_dl_fork, 0));
    const double* src[1] = {s->x};
    const bool halo[1] = {true};
    API_TRY(launch_soa_to_aos(s, 1, src, halo, s->stream_c));
    PIMDB_CUDA_TRY(s, cudaMemcpyAsync(s->dl_hook, s->stage_d, s->S * s->Ploc * sizeof(double), cudaMemcpyDeviceToHost, s->stream_c));
    PIMDB_CUDA_TRY(s, cudaEventRecord(s->ev_dl_join, s->stream_c));
    s->dl_forked = true;
    return PIMDB_OK;
}
static int join_download_x(Sim* s) {
    if (!s->dl_forked) return PIMDB_OK;
    s->dl_forked = false;
    PIMDB_CUDA_TRY(s, cudaStreamWaitEvent(s->stream, s->ev_dl_join, 0));
    return PIMDB_OK;
}

// Bead shard, fixcom, Cartesian Langevin (or no) thermostat: the boundary slices leave one kernel early (OP_HALO_EARLY).
static bool early_halo_push(const Sim* s) {
    static const bool off = getenv("PIMDB_PEER_LATE_HALO") != nullptr;   // A/B timing
    return !off && lazy_closing_com(s) && s->cfg.propagator == PIMDB_PROP_CARTESIAN;
}

// One handle owns every bead, velocity Verlet + Langevin on the counter-based stream + fixcom, forces assembled by the closing
// kick: the iteration whose launches enqueue_step_inner writes out by hand (PIMDB_NO_TICKETLESS=1: the general path, A/B timing)
static bool ticketless_langevin_step(const Sim* s) {
    return !s->no_ticketless && s->all_local && !s->peer_on && s->cfg.fixcom && s->cfg.propagator == PIMDB_PROP_CARTESIAN &&
           s->cfg.thermostat == PIMDB_THERMO_LANGEVIN && !s->cfg.nmthermostat && !s->rm_state && fuse_assembly(s);
}
static int enqueue_step_inner(Sim* s, bool defer_last_com);
static int enqueue_step(Sim* s, bool defer_last_com) {
    s->nz_want = true;
    const int rc = enqueue_step_inner(s, defer_last_com);
    s->nz_want = false;
    return rc;
}
static int enqueue_step_inner(Sim* s, bool defer_last_com) {
    if (early_halo_push(s)) {
        //   [O | SUM -> peers | boundary slices -> neighbours]  [wait sums | SUBCM | B | A | fix received slices]  forces  [assemble | B | O]
        const unsigned o_pre = s->cfg.thermostat == PIMDB_THERMO_LANGEVIN ? OP_O_PRE : 0u;
        const unsigned o_post = s->cfg.thermostat == PIMDB_THERMO_LANGEVIN ? OP_O_POST : 0u;
        s->p_shift_pending = false; s->z_owed = false;
        // (the noise draw counter is advanced by the launch in the middle, which has no O stage, as in the one-handle iteration
        // below: the closing launch then needs no last block at all -- its early credit belongs to the flag protocol of the
        // late halo push, which this path does not use)
        const bool bump = o_pre != 0 && !s->rm_state && !s->no_ticketless;
        s->li_no_ticket = bump;
        API_TRY(launch_integrate(s, o_pre | OP_SUM | OP_HALO_EARLY));
        allow_early_launch(s);
        s->li_draw_bump = bump ? 2 : 0;
        API_TRY(launch_integrate(s, OP_SUBCM | OP_B | OP_A | OP_HALO_FIX));
        API_TRY(maybe_download_x(s));
        const bool fuse = fuse_assembly(s);
        s->nz_first_off = bump ? -1 : 0;      // (the counter already stands at "next opening": the closing draw is the one before)
        const int rcf = enqueue_forces(s, fuse, !s->dl_forked);
        s->nz_first_off = 0;
        API_TRY(rcf);
        s->li_no_ticket = bump; s->li_draw_off = bump ? -1 : 0;
        API_TRY(launch_integrate(s, (fuse ? OP_ASSEMBLE : 0u) | OP_B | o_post));
        API_TRY(join_download_x(s));
        s->z_owed = true;
        return PIMDB_OK;
    }
    if (ticketless_langevin_step(s)) {
        //   [SUBCM | O | SUM]  [SUBCM | B | A | halos]  forces  [assemble | B | O | SUM]
        // The launches the general path below makes for this configuration, stage for stage; what differs is who advances the
        // noise draw counter. There, the last block of every launch with an O stage (a ticket, a fence and an atomic round trip
        // on the tail of two kernels per iteration); here the launch between them, which has no O stage and so no reader of the
        // counter, adds 2 in its prologue: the opening O reads counter + 0 before it, the closing O counter - 1 after it.
        s->li_no_ticket = true;
        API_TRY(launch_integrate(s, (s->p_shift_pending ? OP_SUBCM : 0u) | OP_O_PRE | OP_SUM));
        s->p_shift_pending = false;
        allow_early_launch(s);
        s->li_draw_bump = 2;
        API_TRY(launch_integrate(s, OP_SUBCM | OP_B | OP_A | OP_HALO));
        API_TRY(maybe_download_x(s));
        s->nz_first_off = -1;
        const int rcf = enqueue_forces(s, true, !s->dl_forked);
        s->nz_first_off = 0;
        API_TRY(rcf);
        s->li_no_ticket = true; s->li_draw_off = -1;
        API_TRY(launch_integrate(s, OP_ASSEMBLE | OP_B | OP_O_POST | OP_SUM));
        if (defer_last_com) s->p_shift_pending = true;
        else API_TRY(launch_integrate(s, OP_SUBCM));
        return join_download_x(s);
    }
    Fuser fz(s);
    const bool lazy = lazy_closing_com(s);
    if (lazy) { s->p_shift_pending = false; s->z_owed = false; }   // whatever Z is outstanding is subsumed by this iteration's first
    else if (s->p_shift_pending) { fz.subcm(); s->p_shift_pending = false; }
    thermostat_into(s, fz);
    if (s->cfg.fixcom) { fz.sum(); fz.subcm(); }
    propagator_into(s, fz);
    thermostat_into(s, fz);
    if (s->cfg.fixcom) {
        if (lazy) s->z_owed = true;
        else {
            fz.sum();
            if (defer_last_com) s->p_shift_pending = true;
            else fz.subcm();
        }
    }
    fz.flush();
    if (fz.rc == PIMDB_OK) fz.rc = join_download_x(s);
    return fz.rc;
}

// pimdb_step replays one captured iteration, so every iteration must start in the same state: with fixcom that is
// "a shift is pending". A pending shift of zero is a no-op (p - 0.0 == p bit for bit).
static int make_entry_state_uniform(Sim* s) {
    if (lazy_closing_com(s)) return PIMDB_OK;   // an iteration starts the same way whether or not a Z is outstanding
    if (s->cfg.fixcom && !s->p_shift_pending) {
        if (s->peer_on) API_TRY(launch_integrate(s, OP_ZERO_SUM));   // every rank publishes zero sums
        else if (s->all_local) {      // block partials, summed by the consumer (integrator.cu): zero partials in array 0
            PIMDB_CUDA_TRY(s, cudaMemsetAsync(s->com_part, 0, sizeof(double) * 4 * kMaxPartials, s->stream));
            s->com_buf = 0;
        } else PIMDB_CUDA_TRY(s, cudaMemsetAsync(s->com, 0, sizeof(double) * 4, s->stream));
        s->p_shift_pending = true;
    }
    if (s->cfg.fixcom && s->all_local && !s->peer_on && s->com_buf != 0) {
        // the captured iteration names the partial arrays statically: the pending sums always enter it in array 0
        PIMDB_CUDA_TRY(s, cudaMemcpyAsync(s->com_part, s->com_part + 4 * kMaxPartials, sizeof(double) * 4 * kMaxPartials,
                                          cudaMemcpyDeviceToDevice, s->stream));
        s->com_buf = 0;
    }
    return PIMDB_OK;
}

static int require_all_local(Sim* s, const char* what) {
    if (!s->all_local && !s->peer_on)
        return fail(s, PIMDB_ERR_INVALID_ARGUMENT, std::string(what) + " needs all beads on this handle, or a bead shard attached to its "
                                                   "peers (pimdb_peer_attach); host-driven sharding uses pimdb_step_phase");
    return PIMDB_OK;
}

extern "C" int pimdb_update_neighbors(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    API_TRY(require_all_local(s, "pimdb_update_neighbors"));
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    return s->peer_on ? launch_peer_push_halos(s) : launch_fill_halos(s);
}

extern "C" int pimdb_update_forces(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    return enqueue_forces(s);
}

extern "C" int pimdb_moment_step(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    API_TRY(settle_momenta(s));
    return launch_integrate(s, OP_B);
}

extern "C" int pimdb_coords_step(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    API_TRY(settle_momenta(s));
    return launch_integrate(s, OP_A | ((s->all_local || s->peer_on) ? OP_HALO : 0u));
}

extern "C" int pimdb_propagator_step(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    API_TRY(require_all_local(s, "pimdb_propagator_step"));
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    API_TRY(settle_momenta(s));
    Fuser fz(s);
    propagator_into(s, fz);
    fz.flush();
    return fz.rc;
}

extern "C" int pimdb_thermostat_step(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    API_TRY(settle_momenta(s));
    Fuser fz(s);
    thermostat_into(s, fz);
    fz.flush();
    return fz.rc;
}

extern "C" int pimdb_zero_momentum(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    API_TRY(require_all_local(s, "pimdb_zero_momentum"));
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    API_TRY(settle_momenta(s));
    Fuser fz(s);
    fz.sum();
    fz.subcm();
    fz.flush();
    return fz.rc;
}

extern "C" int pimdb_step(pimdb_sim* sim, int nsteps) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || nsteps < 0) return PIMDB_ERR_INVALID_ARGUMENT;
    API_TRY(require_all_local(s, "pimdb_step"));
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    if (s->timing) {   // eager path with CUDA events around every step and every pair-force launch
        for (int i = 0; i < nsteps; ++i) {
            cudaEvent_t e0, e1;
            PIMDB_CUDA_TRY(s, cudaEventCreate(&e0));
            PIMDB_CUDA_TRY(s, cudaEventCreate(&e1));
            PIMDB_CUDA_TRY(s, cudaEventRecord(e0, s->stream));
            API_TRY(enqueue_step(s, true));
            PIMDB_CUDA_TRY(s, cudaEventRecord(e1, s->stream));
            s->ev_step.emplace_back(e0, e1);
        }
        return PIMDB_OK;
    }
    if (nsteps == 0) return PIMDB_OK;   // nothing to enqueue (and nothing to capture: the entry state below belongs to a real step)
    API_TRY(make_entry_state_uniform(s));
    if (!s->graph_exec) {
        const unsigned long long before = s->launches;
        const bool pending0 = s->p_shift_pending, owed0 = s->z_owed, stale0 = s->split_stale;
        PIMDB_CUDA_TRY(s, cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        s->tl_next = 0;
        s->stamp_next = 0;
        int rc = enqueue_step(s, true);
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamEndCapture(s->stream, &g);
        // nothing has run yet: the host-side flags go back to what they were; the replays below set them
        s->p_shift_pending = pending0; s->z_owed = owed0; s->split_stale = stale0;
        if (rc != PIMDB_OK) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess) s->launches = before;
        PIMDB_CUDA_TRY(s, ce);
        s->graph = g;
        s->graph_kernels = s->launches - before;
        s->launches = before;
        PIMDB_CUDA_TRY(s, cudaGraphInstantiate(&s->graph_exec, s->graph, 0));
    }
    for (int i = 0; i < nsteps; ++i) {
        PIMDB_CUDA_TRY(s, cudaGraphLaunch(s->graph_exec, s->stream));
        s->launches += s->graph_kernels;
    }
    // host-side state after an iteration (what enqueue_step leaves behind)
    if (s->cfg.fixcom) {
        if (lazy_closing_com(s)) { s->z_owed = true; s->p_shift_pending = false; }
        else s->p_shift_pending = true;
    }
    if (s->cfg.propagator == PIMDB_PROP_CARTESIAN && fuse_assembly(s)) s->split_stale = true;
    return PIMDB_OK;
}

// nsteps iterations, then the state on the host. The last iteration is a second captured graph in which the coordinates
// leave for `x` as soon as the drift has produced them (their PCIe copy overlaps the force evaluation); momenta and forces
// follow when the iteration is done. Page-locked destinations only get the overlap; otherwise step + download.
extern "C" int pimdb_step_download(pimdb_sim* sim, int nsteps, double* x, double* p, double* f) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || nsteps < 1) return PIMDB_ERR_INVALID_ARGUMENT;
    API_TRY(require_all_local(s, "pimdb_step_download"));
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    if (!x || s->timing || !host_is_pinned(x)) {
        API_TRY(pimdb_step(sim, nsteps));
        return pimdb_download_state(sim, x, p, f);
    }
    if (nsteps > 1) API_TRY(pimdb_step(sim, nsteps - 1));
    API_TRY(make_entry_state_uniform(s));
    if (!s->graph_dl_exec || s->dl_x_host != x) {
        if (s->graph_dl_exec) { cudaGraphExecDestroy(s->graph_dl_exec); s->graph_dl_exec = nullptr; }
        if (s->graph_dl) { cudaGraphDestroy(s->graph_dl); s->graph_dl = nullptr; }
        const unsigned long long before = s->launches;
        const bool pending0 = s->p_shift_pending, owed0 = s->z_owed, stale0 = s->split_stale;
        PIMDB_CUDA_TRY(s, cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        s->dl_hook = x;
        s->dl_forked = false;
        int rc = enqueue_step(s, true);
        s->dl_hook = nullptr;
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamEndCapture(s->stream, &g);
        s->p_shift_pending = pending0; s->z_owed = owed0; s->split_stale = stale0;
        s->graph_dl_kernels = s->launches - before;
        s->launches = before;
        if (rc != PIMDB_OK) { if (g) cudaGraphDestroy(g); return rc; }
        PIMDB_CUDA_TRY(s, ce);
        s->graph_dl = g;
        s->dl_x_host = x;
        PIMDB_CUDA_TRY(s, cudaGraphInstantiate(&s->graph_dl_exec, s->graph_dl, 0));
    }
    PIMDB_CUDA_TRY(s, cudaGraphLaunch(s->graph_dl_exec, s->stream));
    s->launches += s->graph_dl_kernels;
    if (s->cfg.fixcom) {
        if (lazy_closing_com(s)) { s->z_owed = true; s->p_shift_pending = false; }
        else s->p_shift_pending = true;
    }
    if (s->cfg.propagator == PIMDB_PROP_CARTESIAN && fuse_assembly(s)) s->split_stale = true;
    if (!p && !f) {
        PIMDB_CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        return check_deferred(s);
    }
    return pimdb_download_state(sim, nullptr, p, f);
}

// phases for bead sharding; the host runs the collectives in between (include/pimdb200.h)
extern "C" int pimdb_step_phase(pimdb_sim* sim, int phase) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    if (s->cfg.propagator != PIMDB_PROP_CARTESIAN || s->cfg.nmthermostat)
        return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "bead-sharded phases support the cartesian propagator / thermostat only");
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    API_TRY(settle_momenta(s));
    Fuser fz(s);
    switch (phase) {
        case 0:
            thermostat_into(s, fz);
            if (s->cfg.fixcom) fz.sum();
            break;
        case 1:
            if (s->cfg.fixcom) fz.subcm();
            fz.kick(false);
            fz.put(OP_A, 4);   // no ring-wrap halo write: the host exchanges halos
            break;
        case 2:
            fz.rc = enqueue_forces(s);
            fz.kick(false);
            thermostat_into(s, fz);
            if (s->cfg.fixcom) fz.sum();
            break;
        case 3:
            if (s->cfg.fixcom) fz.subcm();
            break;
        default:
            return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "phase must be 0..3");
    }
    fz.flush();
    return fz.rc;
}

// ----------------------------------------------------------------------------------------------------
// Bead sharding over peer memory. Each handle exports a blob (cudaIpc handles of its coordinate array and of its mailbox
// + what a peer in the same process needs to use the pointers directly); the host gathers the blobs of all ranks (one
// torch.distributed all-gather at start-up, or plainly in a single process driving several GPUs) and hands every handle
// the whole table. From then on halo slices and momentum sums travel inside the step's own kernels (integrator.cu), and
// pimdb_step works on a shard exactly as on a full ring: one CUDA graph per rank, no host call between its kernels.
namespace {
struct PeerBlob {
    unsigned int magic, version;
    long long pid;
    unsigned long long boot_tag;       // distinguishes handles of different processes with recycled pids (create-time stamp)
    int device, natoms, nbeads, ndim, bead_begin, bead_end;
    char uuid[16];                     // of the exporting handle's GPU: two shards on ONE GPU compete for its SMs
    void* x;                           // device pointers, meaningful inside the exporting process only
    void* mailbox;
    cudaIpcMemHandle_t x_ipc, mailbox_ipc;
};
static_assert(sizeof(PeerBlob) <= PIMDB_PEER_BLOB_BYTES, "blob does not fit");
constexpr unsigned int kBlobMagic = 0x50494d44u;   // "PIMD"
}  // namespace

extern "C" int pimdb_peer_export(pimdb_sim* sim, void* blob_out) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !blob_out) return PIMDB_ERR_INVALID_ARGUMENT;
    if (s->all_local) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "this handle owns every bead: there is nothing to share");
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    PeerBlob b;
    memset(&b, 0, sizeof b);
    b.magic = kBlobMagic; b.version = PIMDB_ABI_VERSION;
    b.pid = (long long)getpid();
    {
        cudaDeviceProp prop;
        PIMDB_CUDA_TRY(s, cudaGetDeviceProperties(&prop, s->device));
        memcpy(b.uuid, prop.uuid.bytes, 16);
    }
    b.device = s->device; b.natoms = s->N; b.nbeads = s->P; b.ndim = s->D; b.bead_begin = s->b0; b.bead_end = s->b1;
    b.x = s->x; b.mailbox = s->mailbox;
    PIMDB_CUDA_TRY(s, cudaIpcGetMemHandle(&b.x_ipc, s->x));
    PIMDB_CUDA_TRY(s, cudaIpcGetMemHandle(&b.mailbox_ipc, s->mailbox));
    memset(blob_out, 0, PIMDB_PEER_BLOB_BYTES);
    memcpy(blob_out, &b, sizeof b);
    return PIMDB_OK;
}

static int map_peer(Sim* s, const PeerBlob& b, bool want_x, void** x_out, void** mailbox_out) {
    *x_out = nullptr;
    if (b.pid == (long long)getpid()) {          // same process: the pointers are valid as they are
        if (b.device != s->device) {
            int can = 0;
            PIMDB_CUDA_TRY(s, cudaDeviceCanAccessPeer(&can, s->device, b.device));
            if (!can) return fail(s, PIMDB_ERR_CUDA, "GPUs " + std::to_string(s->device) + " and " + std::to_string(b.device) + " cannot access each other's memory");
            const cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) PIMDB_CUDA_TRY(s, e);
            cudaGetLastError();
        }
        *mailbox_out = b.mailbox;
        if (want_x) *x_out = b.x;
        return PIMDB_OK;
    }
    PIMDB_CUDA_TRY(s, cudaIpcOpenMemHandle(mailbox_out, b.mailbox_ipc, cudaIpcMemLazyEnablePeerAccess));
    s->ipc_opened.push_back(*mailbox_out);
    if (want_x) {
        PIMDB_CUDA_TRY(s, cudaIpcOpenMemHandle(x_out, b.x_ipc, cudaIpcMemLazyEnablePeerAccess));
        s->ipc_opened.push_back(*x_out);
    }
    return PIMDB_OK;
}

extern "C" int pimdb_peer_attach(pimdb_sim* sim, int world, int rank, const void* blobs) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !blobs) return PIMDB_ERR_INVALID_ARGUMENT;
    if (s->peer_on) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "handle is already attached to its peers");
    if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world)
        return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "peer sharding supports 2.." + std::to_string(kMaxPeers) + " ranks");
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    std::vector<PeerBlob> tab(world);
    int expect = 0;
    for (int r = 0; r < world; ++r) {
        memcpy(&tab[r], (const char*)blobs + (size_t)r * PIMDB_PEER_BLOB_BYTES, sizeof(PeerBlob));
        const PeerBlob& b = tab[r];
        if (b.magic != kBlobMagic || b.version != PIMDB_ABI_VERSION) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "not a pimdb peer blob (rank " + std::to_string(r) + ")");
        if (b.natoms != s->N || b.nbeads != s->P || b.ndim != s->D) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "peer " + std::to_string(r) + " holds a different system");
        if (b.bead_begin != expect || b.bead_end <= b.bead_begin) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "bead ranges must tile [0, nbeads) in rank order");
        expect = b.bead_end;
    }
    s->peer_shares_gpu = false;
    for (int r = 0; r < world; ++r)
        if (r != rank && memcmp(tab[r].uuid, tab[rank].uuid, 16) == 0) s->peer_shares_gpu = true;
    if (expect != s->P || tab[rank].bead_begin != s->b0 || tab[rank].bead_end != s->b1)
        return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "bead ranges must tile [0, nbeads) in rank order (and blobs[rank] must be this handle's)");
    PeerDev pd{};
    pd.world = world; pd.rank = rank;
    pd.prev = (rank + world - 1) % world; pd.next = (rank + 1) % world;
    pd.mine = s->mailbox; pd.seq = s->peer_seq;
    pd.ll_mine = reinterpret_cast<unsigned long long*>(s->mailbox + 1);
    const bool nm = s->cfg.propagator == PIMDB_PROP_NORMAL_MODES || s->cfg.nmthermostat;
    if (nm) pd.gather_mine = pd.gather_to[rank] = reinterpret_cast<double*>(pd.ll_mine + 2 * 2 * 2 * s->S);
    unsigned long long ms = 20000;
    if (const char* e = getenv("PIMDB_PEER_TIMEOUT_MS")) ms = (unsigned long long)std::max(1, atoi(e));
    pd.timeout_ns = ms * 1000000ull;
    for (int r = 0; r < world; ++r) {
        if (r == rank) { pd.box[r] = s->mailbox; continue; }
        void *x = nullptr, *mb = nullptr;
        API_TRY(map_peer(s, tab[r], r == pd.prev || r == pd.next, &x, &mb));
        pd.box[r] = reinterpret_cast<PeerMailbox*>(mb);
        const size_t ploc_r = (size_t)(tab[r].bead_end - tab[r].bead_begin);
        if (r == pd.prev) pd.halo_to_prev = reinterpret_cast<double*>(x) + (ploc_r + 1) * s->S;   // its trailing halo slab
        if (r == pd.next) pd.halo_to_next = reinterpret_cast<double*>(x);                          // its leading halo slab
        if (r == pd.prev) pd.ll_to_prev = reinterpret_cast<unsigned long long*>(pd.box[r] + 1);
        if (r == pd.next) pd.ll_to_next = reinterpret_cast<unsigned long long*>(pd.box[r] + 1);
        if (nm) pd.gather_to[r] = reinterpret_cast<double*>(reinterpret_cast<unsigned long long*>(pd.box[r] + 1) + 2 * 2 * 2 * s->S);
    }
    s->peer = pd;
    s->peer_on = true;
    if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
    if (s->graph) { cudaGraphDestroy(s->graph); s->graph = nullptr; }
    // Simulation::updateNeighboringCoordinates for whatever coordinates the handle holds now
    return launch_peer_push_halos(s);
}

// Enqueue (asynchronously) what the next read of the momenta would have to do first. A single host thread that drives
// several shards calls this on every handle before it reads any of them: the closing zeroMomentum is collective, and a
// blocking read of one handle would otherwise wait for sums that the other handles have not been asked to publish yet.
extern "C" int pimdb_settle(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    return settle_momenta(s);
}

extern "C" int pimdb_peer_attached(const pimdb_sim* sim) {
    const Sim* s = reinterpret_cast<const Sim*>(sim);
    return s && s->peer_on ? 1 : 0;
}

// ----------------------------------------------------------------------------------------------------
extern "C" int pimdb_exchange_prepare(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    if (!s->bosonic) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "simulation is not bosonic");
    if (!s->has_first && !s->has_last) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "this handle owns no exterior bead");
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    return launch_exchange(s, s->stream);
}

extern "C" int pimdb_exchange_get(pimdb_sim* sim, int table, double* out, size_t n) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !out) return PIMDB_ERR_INVALID_ARGUMENT;
    if (!s->bosonic) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "simulation is not bosonic");
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    const size_t N = s->N;
    size_t need = 0;
    const double* src = nullptr;
    switch (table) {
        case PIMDB_EXCH_V: need = N + 1; src = s->exV; break;
        case PIMDB_EXCH_VB: need = N + 1; src = s->exVb; break;
        case PIMDB_EXCH_E: need = N * (N + 1) / 2; break;
        case PIMDB_EXCH_PROB: need = N * N; break;
        default: return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "unknown exchange table");
    }
    if (n < need) return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "output buffer too small");
    if (s->factorial && table != PIMDB_EXCH_V)
        return fail(s, PIMDB_ERR_INVALID_ARGUMENT, "the factorial exchange class has no recursion tables (only V[N], its effective potential)");
    if (!src) {
        if (s->exTabCap < need) {
            cudaFree(s->exTab);
            s->exTab = nullptr; s->exTabCap = 0;
            PIMDB_CUDA_TRY(s, cudaMalloc(&s->exTab, need * sizeof(double)));
            s->exTabCap = need;
        }
        API_TRY(launch_exchange_tables(s, table));
        src = s->exTab;
    }
    PIMDB_CUDA_TRY(s, cudaMemcpyAsync(out, src, need * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    PIMDB_CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    return check_deferred(s);
}

// Observable::calculate for energy / classical / bosonic (src/observables/*.cpp), partials of the owned beads
extern "C" int pimdb_observables_calc(pimdb_sim* sim, pimdb_observables* out) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !out) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    API_TRY(settle_momenta(s));
    API_TRY(launch_peer_wait_halos(s));
    if (s->pair_on) API_TRY(refresh_split_forces(s));   // the pair estimators reuse the scratch slab that holds the pair partials
    PIMDB_CUDA_TRY(s, cudaMemsetAsync(s->obs_d, 0, sizeof(DevObs), s->stream));
    API_TRY(launch_obs_elementwise(s));
    if (s->pair_on) {
        for (int lo = 0; lo < s->Ploc; lo += s->bead_chunk)
            API_TRY(launch_pair_chunk(s, lo, std::min(s->bead_chunk, s->Ploc - lo), true));
    }
    if (s->bosonic && s->has_first) {
        API_TRY(launch_exchange(s, s->stream));   // tables for the current positions
        if (!s->factorial) API_TRY(launch_exchange_estimators(s));   // (the factorial class leaves its estimators with the forces)
    }
    if (s->cfg.thermostat >= PIMDB_THERMO_NOSE_HOOVER) API_TRY(launch_nose_hoover_energy(s, &s->obs_d->nh_energy));
    PIMDB_CUDA_TRY(s, cudaMemcpyAsync(s->obs_h, s->obs_d, sizeof(DevObs), cudaMemcpyDeviceToHost, s->stream));
    PIMDB_CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    API_TRY(check_deferred(s));
    const DevObs& o = *s->obs_h;
    const double P = s->P, N = s->N, D = s->D;
    memset(out, 0, sizeof *out);
    // energy.cpp:30-48: per bead NDIM*N/(2 beta); classical links subtract their spring energy / P;
    // the bead-0 owner of a bosonic system adds primEstimator() = e[N]/P instead
    out->kinetic = s->Ploc * (0.5 * D * N / s->beta) - o.spring_e[0] / P;
    if (s->bosonic && s->has_first) out->kinetic += o.prim_est / P;
    const bool ext_on = s->cfg.ext_potential != PIMDB_POT_FREE;
    const bool int_on = s->cfg.int_potential != PIMDB_POT_FREE;
    // energy.cpp:56-111
    if (ext_on || int_on) {
        out->potential = ((ext_on ? o.ext_v : 0.0) + o.pair_v) / P;
        out->virial = ((ext_on ? o.ext_vir : 0.0) + o.pair_vir) * (0.5 / P);
    }
    if (ext_on && int_on) {
        out->ext_pot = o.ext_v / P;
        out->int_pot = o.pair_v / P;
    }
    // classical.cpp:32-78
    out->cl_kinetic = o.p2 * (0.5 / s->cfg.mass);
    out->temperature = 2.0 * out->cl_kinetic / (D * N * P) / P;
    out->cl_spring = o.spring_e[0] + ((s->bosonic && s->has_first) ? o.v_n : 0.0);
    out->nh_energy = o.nh_energy;   // classical.cpp:24-26
    // gsf_action.cpp:21-73 (alpha = 0, IPI convention: spring constant / P): an odd bead contributes
    // -beta (-V_b/(3P) + alpha F2_b/(9 k P^2)) and V_b/(P/2) to pot_gsf, an even bead -beta (V_b/(3P) + (1-alpha) F2_b/(9 k P^2))
    if (!int_on) {
        const double alpha = 0.0, sp = s->kspring / P;
        const double w_odd = (-1.0) * (o.gsf[0] / (3 * P)) + alpha * (o.gsf[2] / (9 * sp * P * P));
        const double w_even = o.gsf[1] / (3 * P) + (1 - alpha) * (o.gsf[3] / (9 * sp * P * P));
        out->w_gsf = (-1.0) * s->beta * (w_odd + w_even);
        out->pot_gsf = o.gsf[0] / (0.5 * P);
    } else {
        out->w_gsf = out->pot_gsf = std::nan("");
    }
    // bosonic.cpp:17-22, quadratic_bosonic_exchange.cpp:222-240
    if (s->bosonic && s->has_first && !s->factorial) {   // (factorial_bosonic_exchange.cpp:234-251: both "not implemented", 0)
        out->prob_dist = std::exp(-s->exch_beta * (o.e_diag_sum - o.v_n) - std::lgamma(N + 1.0));
        out->prob_all = std::exp(-s->exch_beta * (o.e_full - o.v_n));
    }
    return PIMDB_OK;
}

// ----------------------------------------------------------------------------------------------------
extern "C" void* pimdb_get_stream(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    return s ? (void*)s->stream : nullptr;
}

extern "C" int pimdb_set_stream(pimdb_sim* sim, void* cuda_stream) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    PIMDB_CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    s->stream = (cudaStream_t)cuda_stream;
    s->own_stream = false;
    return PIMDB_OK;
}

extern "C" void* pimdb_halo_ptr(pimdb_sim* sim, int which, size_t* count) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return nullptr;
    if (count) *count = s->S;
    switch (which) {
        case 0: return s->x + s->S;
        case 1: return s->x + (size_t)s->Ploc * s->S;
        case 2: return s->x;
        case 3: return s->x + (size_t)(s->Ploc + 1) * s->S;
        default: return nullptr;
    }
}

extern "C" void* pimdb_com_ptr(pimdb_sim* sim) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    return s ? (void*)s->com : nullptr;
}

extern "C" unsigned long long pimdb_launch_count(const pimdb_sim* sim) {
    const Sim* s = reinterpret_cast<const Sim*>(sim);
    return s ? s->launches : 0ull;
}

extern "C" int pimdb_timing_enable(pimdb_sim* sim, int on) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    PIMDB_CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    for (auto* v : {&s->ev_pair, &s->ev_step, &s->ev_integ}) {
        for (auto& e : *v) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        v->clear();
    }
    s->integ_bytes = 0.0;
    s->timing = on != 0;
    return PIMDB_OK;
}

// what = 0: pair-force kernel launches, 1: whole steps, 2: fused integrator launches
extern "C" int pimdb_timing_get(pimdb_sim* sim, int what, double* ms_avg, unsigned long long* count) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !ms_avg || !count) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    PIMDB_CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    auto& v = what == 0 ? s->ev_pair : (what == 2 ? s->ev_integ : s->ev_step);
    double tot = 0.0;
    for (auto& e : v) {
        float ms = 0.f;
        PIMDB_CUDA_TRY(s, cudaEventElapsedTime(&ms, e.first, e.second));
        tot += ms;
    }
    *count = v.size();
    *ms_avg = v.empty() ? 0.0 : tot / (double)v.size();
    return PIMDB_OK;
}

// algorithmic bytes per launch of the fused integrator launches timed so far (what pimdb_timing_get(2) averages over)
extern "C" int pimdb_timing_integrator_bytes(pimdb_sim* sim, double* bytes_per_launch) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !bytes_per_launch) return PIMDB_ERR_INVALID_ARGUMENT;
    *bytes_per_launch = s->ev_integ.empty() ? 0.0 : s->integ_bytes / (double)s->ev_integ.size();
    return PIMDB_OK;
}

// ----------------------------------------------------------------------------------------------------
// FP64 FMA micro-benchmark: the denominator of the pair-force roofline (MEASURED_PEAKS.json has no FP64 figure).
// 8 independent accumulators per thread, 4096 dependent rounds, 148 x 8 blocks of 256 threads.
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int rounds, double a, double b) {
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < rounds; ++i) {
        r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
        r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
}

extern "C" int pimdb_bench_fp64_peak(int device, double* tflops) {
    if (!tflops) return PIMDB_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return PIMDB_ERR_CUDA;
    const int blocks = kNumSM * 8, threads = 256, rounds = 4096;
    double* buf = nullptr;
    if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) return PIMDB_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k_dfma_peak<<<blocks, threads>>>(buf, rounds, 0.999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(buf); return PIMDB_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * rounds * (double)blocks * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) * 1e-12);   // first launch = warm-up
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf);
    *tflops = best;
    return PIMDB_OK;
}

// ----------------------------------------------------------------------------------------------------
// Profiling aid (PIMDB_TIMELINE=1, not declared in pimdb200.h): [first block start, last block end] in ns of the
// kernels of the captured step, in launch order, accumulated (min / max) since the last call; the slots are reset.
extern "C" int pimdb_debug_timeline(pimdb_sim* sim, unsigned long long* out) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !out || !s->tl) return 0;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    cudaMemcpy(out, s->tl, sizeof(unsigned long long) * 64, cudaMemcpyDeviceToHost);
    std::vector<unsigned long long> init(64);
    for (int i = 0; i < 32; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0ull; }
    cudaMemcpy(s->tl, init.data(), sizeof(unsigned long long) * 64, cudaMemcpyHostToDevice);
    return s->tl_next;
}

// ... and which kernel each of those slots belongs to (see tl_slot): out[32]
extern "C" int pimdb_debug_timeline_kinds(pimdb_sim* sim, unsigned char* out) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !out || !s->tl) return 0;
    memcpy(out, s->tl_kind, 32);
    return s->tl_next;
}

// Profiling aid (PIMDB_TIMELINE=1): the %globaltimer stamps of the phases of the k_integrate launches of the captured step
// (block 0: 0 start, 1 counters read, 2 sums in, 3 halo / credit waits done, 4 main loop done, 5 ticket taken; last block: 6, 7 end).
extern "C" int pimdb_debug_integrate_stamps(pimdb_sim* sim, unsigned long long* out) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !out || !s->stamps) return 0;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    cudaMemcpy(out, s->stamps, sizeof(unsigned long long) * 64, cudaMemcpyDeviceToHost);
    return s->stamp_next;
}

// ----------------------------------------------------------------------------------------------------
// Test aid (not part of the reference surface, not declared in pimdb200.h): how the blocked exchange recurrence
// solved each 32-row block in its last run, in step order. out[0 .. nb) forward, out[nb .. 2nb) backward:
// 1 = one matrix-vector product with the precomputed block inverse, 2 = exact sequential extended-range steps,
// 0 = block without steps. Returns the number of blocks per direction (0 when the blocked kernel does not apply).
extern "C" int pimdb_debug_exchange_blocks(pimdb_sim* sim, int* out) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !out || !s->bosonic || !s->exGok) return 0;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    cudaStreamSynchronize(s->stream_x);
    const int nb = (s->N + 31) / 32;
    if (cudaMemcpy(out, s->exGok + 2 * nb, sizeof(int) * 2 * nb, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    if (!getenv("PIMDB_EXCH_REASONS"))
        for (int i = 0; i < 2 * nb; ++i) out[i] &= 15;     // status only; the high bits say why a block went exact
    return nb;
}

// ----------------------------------------------------------------------------------------------------
// Profiling aid (not part of the reference surface, not declared in pimdb200.h): the pair-tile kernel of every owned bead,
// `reps` launches back to back on the handle's stream with nothing beside them. The caller brackets the call with CUDA
// events: (elapsed / reps) is the kernel's duration when it has the GPU to itself (bench.py, roofline.alone).
extern "C" int pimdb_debug_pair_tiles_only(pimdb_sim* sim, int reps) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !s->pair_on || reps < 1) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    const bool timing = s->timing;
    s->timing = false;
    int rc = PIMDB_OK;
    for (int r = 0; r < reps && rc == PIMDB_OK; ++r)
        for (int lo = 0; lo < s->Ploc && rc == PIMDB_OK; lo += s->bead_chunk)
            rc = launch_pair_chunk(s, lo, std::min(s->bead_chunk, s->Ploc - lo), false);
    s->timing = timing;
    return rc;
}

// Profiling aid (not part of the reference surface, not declared in pimdb200.h): average warm duration in
// microseconds of the two halves of the exchange chain, measured with CUDA events on the handle's stream.
//   out[0] = prefix sums + Boltzmann factors, out[1] = recurrences + exterior forces
extern "C" int pimdb_debug_exchange_timing(pimdb_sim* sim, int reps, double* out) {
    Sim* s = reinterpret_cast<Sim*>(sim);
    if (!s || !out || !s->bosonic) return PIMDB_ERR_INVALID_ARGUMENT;
    PIMDB_CUDA_TRY(s, cudaSetDevice(s->device));
    cudaEvent_t e[3];
    for (auto& ev : e) cudaEventCreate(&ev);
    double acc[2] = {0.0, 0.0};
    for (int r = 0; r < reps + 2; ++r) {
        cudaEventRecord(e[0], s->stream);
        API_TRY(launch_exchange_part(s, s->stream, 0));
        cudaEventRecord(e[1], s->stream);
        API_TRY(launch_exchange_part(s, s->stream, 1));
        cudaEventRecord(e[2], s->stream);
        PIMDB_CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        if (r >= 2) {
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, e[0], e[1]);
            cudaEventElapsedTime(&b, e[1], e[2]);
            acc[0] += a; acc[1] += b;
        }
    }
    for (auto& ev : e) cudaEventDestroy(ev);
    out[0] = acc[0] / reps * 1e3;
    out[1] = acc[1] / reps * 1e3;
    if (getenv("PIMDB_EXCH_DEBUG")) {   // per-warp clock64 stamps of one more run: out[2 + 3*(dir*32 + warp) + {0,1,2}]
        constexpr int kDbg = 64 * 3 + 2048 + 32;   // per-warp totals | per-(warp, block) stamps | owner-phase start
        if (!s->dbg_buf) cudaMalloc(&s->dbg_buf, sizeof(long long) * kDbg);
        cudaMemset(s->dbg_buf, 0, sizeof(long long) * kDbg);
        API_TRY(launch_exchange_part(s, s->stream, 0));
        API_TRY(launch_exchange_part(s, s->stream, 1));
        PIMDB_CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        std::vector<long long> h(kDbg);
        cudaMemcpy(h.data(), s->dbg_buf, sizeof(long long) * kDbg, cudaMemcpyDeviceToHost);
        const int nout = getenv("PIMDB_EXCH_DEBUG_FULL") ? kDbg : 64 * 3;   // the caller sizes `out` accordingly
        for (int i = 0; i < nout; ++i) out[2 + i] = (double)h[i];
        cudaFree(s->dbg_buf);
        s->dbg_buf = nullptr;
    }
    return PIMDB_OK;
}
