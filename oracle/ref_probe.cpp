// TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// ref_probe: a tiny driver that links the UNMODIFIED reference objects (everything under
// /root/reference/src except pimdb.cpp) and evaluates the reference's own hot path on positions /
// momenta that we hand it as raw doubles, so the CUDA path and the C restatement (oracle/pimd_oracle.c)
// can be compared with the real thing at full FP64 precision instead of the 13 printed digits of the
// reference's state dumps.
//
//   PIMDB_NP=<P> ref_probe <config.ini> <in_dir> <out_dir> forces
//   PIMDB_NP=<P> ref_probe <config.ini> <in_dir> <out_dir> traj <K> [dump_every]
//
// in_dir/x.bin  : [P][N][NDIM] doubles, atomic units (bead-major, AoS like the reference's dVec)
// in_dir/p.bin  : same shape, momenta (optional; if absent the reference's own initial momenta stay)
//
// "forces" reproduces what VelocityVerletPropagator::step does between the A and the second B step
// (src/propagators/velocity_verlet.cpp:15-18): updateNeighboringCoordinates(); updateForces(); and then
// evaluates every enabled observable exactly like Simulation::run + ObservablesLogger::log do
// (src/simulation.cpp:272-279, src/observables/observable.cpp:92-116).
// Outputs (out_dir): f.bin, f_spring.bin, f_phys.bin  [P][N][NDIM]; obs.txt ("name value" per line, bead-summed);
// bosonic runs add exch_V.bin [N+1], exch_Vb.bin [N+1], exch_E.bin [N(N+1)/2] (reference serial order),
// exch_prob.bin [N][N] and exch_scalars.txt.
//
// "traj K" runs K iterations of the body of Simulation::run (src/simulation.cpp:246-259: thermostat half
// step, optional COM removal, propagator step, thermostat half step, optional COM removal) and dumps
// x/p/f after every `dump_every` iterations as x_<it>.bin etc. The first half-kick uses zero forces, as in the
// reference (forces are never evaluated before the first propagator step).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <random>
#include <ranges>
#include <sstream>
#include <string>
#include <vector>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include "mpi.h"

// The exchange tables (V_backwards, connection_probabilities) have no public getter in the reference.
// The probe reads them through the class definition with access control switched off; the compiled
// reference objects are untouched (access specifiers do not change layout).
#define private public
#define protected public
#include "params.h"
#include "simulation.h"
#include "observables.h"
#include "propagators.h"
#include "thermostats.h"
#include "normal_modes.h"
#undef private
#undef protected

static bool read_slab(const std::string& path, int bead, int n, dVec& dst) {
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    size_t bytes = size_t(n) * NDIM * sizeof(double);
    ssize_t got = pread(fd, dst.data(), bytes, off_t(bead) * off_t(bytes));
    close(fd);
    if (got != (ssize_t)bytes) { fprintf(stderr, "ref_probe: short read on %s\n", path.c_str()); exit(2); }
    return true;
}

static void write_slab(const std::string& path, int bead, int n, const double* src) {
    int fd = open(path.c_str(), O_WRONLY | O_CREAT, 0644);
    if (fd < 0) { perror(path.c_str()); exit(2); }
    size_t bytes = size_t(n) * NDIM * sizeof(double);
    if (pwrite(fd, src, bytes, off_t(bead) * off_t(bytes)) != (ssize_t)bytes) { perror("pwrite"); exit(2); }
    close(fd);
}

static void write_vec(const std::string& path, const std::vector<double>& v) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); exit(2); }
    fwrite(v.data(), sizeof(double), v.size(), f);
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: PIMDB_NP=P %s config.ini in_dir out_dir forces|traj [K] [dump_every]\n", argv[0]);
        return 2;
    }
    MPI_Init(&argc, &argv);
    int rank, size;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    const std::string ini = argv[1], in_dir = argv[2], out_dir = argv[3], mode = argv[4];
    int rc = 0;
    try {
        std::streambuf* old = std::cout.rdbuf();
        std::ostringstream sink;
        std::cout.rdbuf(sink.rdbuf());  // silence the reference's status lines
        Params params(ini, rank);
        Simulation sim(rank, size, params, 1u);
        std::cout.rdbuf(old);
        if (sim.nbeads != size) {
            if (rank == 0) fprintf(stderr, "ref_probe: PIMDB_NP (%d) != nbeads (%d)\n", size, sim.nbeads);
            MPI_Finalize();
            return 2;
        }
        const int N = sim.natoms;
        read_slab(in_dir + "/x.bin", rank, N, sim.coord);
        read_slab(in_dir + "/p.bin", rank, N, sim.momenta);
        sim.updateNeighboringCoordinates();
        mkdir(out_dir.c_str(), 0755);

        if (mode == "forces") {
            sim.updateForces();
            dVec fs(N), fp(N);
            sim.updateSpringForces(fs);
            sim.updatePhysicalForces(fp);
            write_slab(out_dir + "/f.bin", rank, N, sim.forces.data());
            write_slab(out_dir + "/f_spring.bin", rank, N, fs.data());
            write_slab(out_dir + "/f_phys.bin", rank, N, fp.data());

            // Observables exactly as the run loop does: calculate on every bead, sum over beads.
            std::ofstream obs;
            if (rank == 0) obs.open(out_dir + "/obs.txt");
            for (const auto& o : sim.observables) {
                o->resetValues();
                o->calculate();
            }
            for (const auto& o : sim.observables) {
                for (auto it = o->quantities.begin(); it != o->quantities.end(); ++it) {
                    double local = it.value(), total = 0.0;
                    MPI_Allreduce(&local, &total, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
                    if (rank == 0) {
                        char buf[64];
                        snprintf(buf, sizeof buf, "%.17g", total);
                        obs << it.key() << ' ' << buf << '\n';
                    }
                }
            }
            if (sim.bosonic && rank == 0) {
                auto* ex = dynamic_cast<BosonicExchange*>(sim.bosonic_exchange.get());
                if (ex) {
                    write_vec(out_dir + "/exch_V.bin", ex->V);
                    write_vec(out_dir + "/exch_Vb.bin", ex->V_backwards);
                    if (N <= 2048) {
                        write_vec(out_dir + "/exch_E.bin", ex->E_kn);
                        write_vec(out_dir + "/exch_prob.bin", ex->connection_probabilities);
                    }
                    FILE* f = fopen((out_dir + "/exch_scalars.txt").c_str(), "w");
                    fprintf(f, "effective_potential %.17g\n", ex->effectivePotential());
                    fprintf(f, "prim_estimator %.17g\n", ex->primEstimator());
                    fprintf(f, "prob_dist %.17g\n", ex->getDistinctProbability());
                    fprintf(f, "prob_all %.17g\n", ex->getLongestProbability());
                    fclose(f);
                }
            }
        } else if (mode == "traj") {
            const int K = argc > 5 ? atoi(argv[5]) : 1;
            const int every = argc > 6 ? atoi(argv[6]) : K;
            for (int it = 1; it <= K; ++it) {
                sim.setStep(it - 1);
                sim.thermostat->step();
                if (sim.fixcom) sim.zeroMomentum();
                sim.propagator->step();
                sim.thermostat->step();
                if (sim.fixcom) sim.zeroMomentum();
                if (it % every == 0 || it == K) {
                    write_slab(out_dir + "/x_" + std::to_string(it) + ".bin", rank, N, sim.coord.data());
                    write_slab(out_dir + "/p_" + std::to_string(it) + ".bin", rank, N, sim.momenta.data());
                    write_slab(out_dir + "/f_" + std::to_string(it) + ".bin", rank, N, sim.forces.data());
                }
            }
        } else {
            if (rank == 0) fprintf(stderr, "ref_probe: unknown mode %s\n", mode.c_str());
            rc = 2;
        }
    } catch (const std::exception& ex) {
        fprintf(stderr, "ref_probe[rank %d]: exception: %s\n", rank, ex.what());
        rc = 3;
        // A throw on one rank would leave the others in a barrier; bail out hard like mpirun would.
        fflush(stderr);
        _exit(rc);
    }
    MPI_Finalize();
    return rc;
}
