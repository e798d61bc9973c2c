// pimdb_gpu: the reference's `pimdb` entry point (src/pimdb.cpp:31-68) on top of the B200 hot path.
//   pimdb_gpu [-in config.ini] [--dim D] [--device K] [--gpus G] [--rng philox|ranmars] [--factorial] [--bosonic_alg]
// --gpus G shards the beads over G GPUs (devices K .. K+G-1) coupled through peer memory: the counterpart of the
// reference's `mpirun -np P pimdb` (README.md:200-203), in one process.
// --rng ranmars (or PIMDB_RNG=ranmars) draws the Langevin noise from the reference's own generator, one sequential
// RANMAR stream per bead (libs/random_mars.cpp): thermostatted runs then follow the reference's trajectories.
// Same INI schema, same output/ files (simulation.out, position_b.xyz, velocity_b.dat, force_b.dat, report.txt),
// same error reporting ("[X] <kind>: <message>", exit code 0 like the reference). NDIM is a run-time flag here
// (compile-time in the reference, CMakeLists.txt:44-48).
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "pimdb_host.hpp"

int main(int argc, char** argv) {
    std::string config = "config.ini";
    int ndim = 3, device = 0, ngpus = 1;
    std::string rng = std::getenv("PIMDB_RNG") ? std::getenv("PIMDB_RNG") : "philox";
    bool info = false, factorial = false;
    try {
        for (int i = 1; i < argc; ++i) {
            if (!std::strcmp(argv[i], "--dim")) {
                if (i + 1 < argc && std::isdigit((unsigned char)argv[i + 1][0])) ndim = std::atoi(argv[++i]);
                else { std::cout << "Program runs 1-, 2- and 3-dimensional systems (select with --dim D)\n"; info = true; }
            } else if (!std::strcmp(argv[i], "--bosonic_alg")) {
                std::cout << "Program runs the quadratic bosonic algorithm (--factorial selects the factorial one, natoms <= 10).\n";
                info = true;
            } else if (!std::strcmp(argv[i], "--factorial")) {
                factorial = true;
            } else if (!std::strcmp(argv[i], "--device")) {
                if (i + 1 < argc) device = std::atoi(argv[++i]);
            } else if (!std::strcmp(argv[i], "--gpus")) {
                if (i + 1 < argc) ngpus = std::atoi(argv[++i]);
            } else if (!std::strcmp(argv[i], "--rng")) {
                if (i + 1 < argc) rng = argv[++i];
            } else if (!std::strcmp(argv[i], "-in")) {
                if (i + 1 < argc) config = argv[++i];
                else throw std::invalid_argument("-in option requires a filename argument");
            }
        }
        if (!info) {
            std::cout << "[*] Initializing the simulation parameters\n";
            pimdb_host::Params params(config, ndim);
            if (rng == "ranmars") params.cfg.rng = PIMDB_RNG_RANMARS;
            else if (rng != "philox") throw std::invalid_argument("--rng takes philox or ranmars");
            if (factorial) params.cfg.exchange_alg = PIMDB_EXCH_FACTORIAL;   // the reference: -DFACTORIAL_BOSONIC_ALGORITHM at build time
            pimdb_host::Simulation sim(params, device, ngpus);
            sim.run();
        }
    } catch (const std::invalid_argument& ex) {
        std::cout << "[X] Invalid argument error: " << ex.what() << '\n';
    } catch (const std::overflow_error& ex) {
        std::cout << "[X] Overflow error: " << ex.what() << '\n';
    } catch (const std::exception& ex) {
        std::cout << "[X] Error: " << ex.what() << '\n';
    }
    return 0;
}
