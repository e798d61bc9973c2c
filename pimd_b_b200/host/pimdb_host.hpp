// Host-side C++20 mirror of the reference's plugin surface, forwarding to the C ABI (include/pimdb200.h).
//
// The reference builds five kinds of objects in Simulation's factories and calls them from Simulation::run
// (include/potentials/potential.h:6-29, include/bosonic_exchange/bosonic_exchange_base.h:16-54,
// include/propagators/propagator.h:7-18, include/thermostats/thermostat.h:9-19,
// include/observables/observable.h:14-52, include/simulation.h:20-131). The classes below keep those names,
// method names, argument meaning and exception types, but the numerics run in libpimdb200.so on the GPU for all
// beads at once (there is no per-rank state any more): a Potential is a descriptor that fills the device
// configuration, a Propagator / Thermostat / BosonicExchange forwards its step to the handle owned by Simulation.
#pragma once

#include <cstddef>
#include <fstream>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pimdb200.h"

namespace pimdb_host {

// ------------------------------------------------------------------ units (values of include/units.h:27-153)
namespace Units {
double convertToInternal(const std::string& family, const std::string& unit, double number);  // src/units.cpp:42-60
double convertToUser(const std::string& family, const std::string& unit, double number);      // src/units.cpp:62-64
}  // namespace Units

// ------------------------------------------------------------------ INI (the subset of inih behaviour the reference relies on)
class IniFile {
public:
    explicit IniFile(const std::string& filename);
    bool ok() const { return ok_; }
    std::string Get(const std::string& section, const std::string& name, const std::string& def) const;
    long GetInteger(const std::string& section, const std::string& name, long def) const;   // strtol semantics
    double GetReal(const std::string& section, const std::string& name, double def) const;  // strtod semantics
    bool GetBoolean(const std::string& section, const std::string& name, bool def) const;
    bool Has(const std::string& section, const std::string& name) const;

private:
    bool ok_ = false;
    std::map<std::string, std::string> values_;  // "section=name" lower-cased
};

// ------------------------------------------------------------------ Params (src/params.cpp:8-260)
struct Params {
    explicit Params(const std::string& filename, int ndim);
    pimdb_config cfg{};              // everything the device needs, atomic units
    long steps = 100000, sfreq = 1000;
    double threshold = 0.1;
    std::string init_pos_type = "random", init_pos_spec;   // random | grid | xyz(<fmt>)
    std::string init_vel_type = "random", init_vel_spec;   // random | manual | manual(<fmt>)
    std::string interaction_name = "free", external_name = "free", propagator_type = "cartesian",
                thermostat_type = "langevin";
    std::string out_positions = "off", out_velocities = "off", out_forces = "off";
    std::string obs_energy = "kelvin", obs_classical = "off", obs_bosonic = "off", obs_gsf = "off";
    static double getQuantity(const std::string& family, const std::string& input);  // "<number> <unit>"
};

class Simulation;

// ------------------------------------------------------------------ Potential (include/potentials/potential.h)
// A descriptor: name + parameters. V / gradV of the reference are evaluated inside the force kernels.
class Potential {
public:
    Potential(std::string name, int id) : name_(std::move(name)), id_(id) {}
    virtual ~Potential() = default;
    const std::string& name() const { return name_; }
    int id() const { return id_; }

private:
    std::string name_;
    int id_;
};

// ------------------------------------------------------------------ BosonicExchangeBase (bosonic_exchange_base.h:16-54)
class BosonicExchange {
public:
    explicit BosonicExchange(Simulation& sim) : sim_(sim) {}
    void prepare();                                        // evaluateBosonicEnergies on the device
    void exteriorSpringForce(std::vector<double>& f, int bead);  // [N][NDIM] spring force on bead 0 or P-1
    double effectivePotential();                           // V[N]
    double primEstimator();
    double getDistinctProbability();
    double getLongestProbability();
    double getVn(int n);
    std::vector<double> getV();

private:
    Simulation& sim_;
};

// ------------------------------------------------------------------ Propagator (include/propagators/propagator.h)
class Propagator {
public:
    explicit Propagator(Simulation& sim) : sim(sim) {}
    virtual ~Propagator() = default;
    virtual void step();   // VelocityVerletPropagator::step or NormalModesPropagator::step, by configuration
    void momentStep();
    void coordsStep();

protected:
    Simulation& sim;
};
using VelocityVerletPropagator = Propagator;
using NormalModesPropagator = Propagator;

// ------------------------------------------------------------------ Thermostat (include/thermostats/thermostat.h)
class Thermostat {
public:
    explicit Thermostat(Simulation& sim) : sim(sim) {}
    virtual ~Thermostat() = default;
    void step();
    virtual double getAdditionToH() { return 0.0; }

protected:
    Simulation& sim;
};
using LangevinThermostat = Thermostat;

// ------------------------------------------------------------------ Observable (include/observables/observable.h)
class Observable {
public:
    Observable(Simulation& sim, std::string out_unit) : sim(sim), out_unit(std::move(out_unit)) {}
    virtual ~Observable() = default;
    virtual void calculate() = 0;
    void initialize(const std::vector<std::string>& labels);
    void resetValues();
    std::vector<std::pair<std::string, double>> quantities;   // insertion-ordered like tsl::ordered_map

protected:
    double& q(const std::string& label);
    Simulation& sim;
    std::string out_unit;
};

class EnergyObservable : public Observable {    // src/observables/energy.cpp
public:
    EnergyObservable(Simulation& sim, const std::string& out_unit);
    void calculate() override;
};
class ClassicalObservable : public Observable { // src/observables/classical.cpp
public:
    ClassicalObservable(Simulation& sim, const std::string& out_unit);
    void calculate() override;
};
class BosonicObservable : public Observable {   // src/observables/bosonic.cpp
public:
    BosonicObservable(Simulation& sim, const std::string& out_unit);
    void calculate() override;
};

class GSFActionObservable : public Observable {  // src/observables/gsf_action.cpp (free interaction only)
public:
    GSFActionObservable(Simulation& sim, const std::string& out_unit);
    void calculate() override;
};

class ObservablesLogger {                        // src/observables/observable.cpp:61-116
public:
    ObservablesLogger(const std::string& filename, const std::vector<std::unique_ptr<Observable>>& observables);
    void log(long step);

private:
    std::ofstream file;
    const std::vector<std::unique_ptr<Observable>>& observables;
};

// ------------------------------------------------------------------ State dumps (src/states/*.cpp)
class State {
public:
    State(Simulation& sim, std::string kind, long freq, const std::string& out_unit);
    void initialize();
    void output(long step);

private:
    Simulation& sim;
    std::string kind, family;
    long freq;
    double factor;
    std::vector<std::ofstream> files;
};

// ------------------------------------------------------------------ Simulation (include/simulation.h:20-131)
class Simulation {
public:
    // ngpus > 1: the beads are sharded over that many GPUs (devices device, device+1, ...), one handle each, coupled through
    // peer memory (pimdb_peer_export / pimdb_peer_attach) -- what `mpirun -np P pimdb` is to the reference (README.md:200-203)
    Simulation(Params& params, int device = 0, int ngpus = 1);
    ~Simulation();
    Simulation(const Simulation&) = delete;

    void run();                                   // src/simulation.cpp:222-290
    void updateForces();                          // :353-374
    void updateNeighboringCoordinates();          // :379-382
    void zeroMomentum();                          // :581-603

    // host copies of the device state, [P][N][NDIM]; pull*/push* move them across the ABI
    std::vector<double> coord, momenta, forces;
    void pushCoord();
    void pushMomenta();
    void pullCoord();
    void pullMomenta();
    void pullForces();
    const pimdb_observables& deviceObservables(); // cached per MD step

    int natoms, nbeads, ndim;
    long steps, sfreq;
    double threshold, dt, mass, beta, temperature, size;
    bool bosonic, fixcom, pbc;
    std::string external_potential_name, interaction_potential_name, thermostat_type, propagator_type;

    std::unique_ptr<Potential> ext_potential, int_potential;
    std::unique_ptr<BosonicExchange> bosonic_exchange;
    std::unique_ptr<Propagator> propagator;
    std::unique_ptr<Thermostat> thermostat;
    std::vector<std::unique_ptr<Observable>> observables;
    std::vector<std::unique_ptr<State>> states;

    pimdb_sim* handle = nullptr;                  // the shard that owns bead 0 (the only one on a single GPU)
    std::vector<pimdb_sim*> handles;              // all shards, in bead order
    std::vector<int> bead_begin;                  // first bead of every shard (+ nbeads at the end)
    void check(int rc, pimdb_sim* h = nullptr) const;   // status code -> the reference's exception type
    void forEach(int (*fn)(pimdb_sim*));          // the same (asynchronous) call on every shard
    void pushArray(int which, const std::vector<double>& host);
    void pullArray(int which, std::vector<double>& host);
    long getStep() const { return md_step; }

private:
    void initializePositions(const Params& p);    // :708-728 (random | grid | xyz)
    void initializeMomenta(const Params& p);      // :736-753 (random | manual)
    void printReport(double wall_time) const;     // :517-572
    Params& params;
    long md_step = 0;
    long obs_step = -1;
    pimdb_observables obs_cache{};
};

}  // namespace pimdb_host
