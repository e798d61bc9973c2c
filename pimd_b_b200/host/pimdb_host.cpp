// Host-side C++20 mirror of the reference's plugin surface (see pimdb_host.hpp). No numerics of the hot path live
// here: everything per-step is a call into libpimdb200.so. What is restated on the host is what the reference does
// once per run around the path: INI/units parsing, initial conditions, the output writers.
#include "pimdb_host.hpp"

#include <algorithm>
#include <array>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <filesystem>
#include <format>
#include <iostream>
#include <sstream>

namespace pimdb_host {

// ====================================================================================== units
namespace {
const std::vector<std::pair<std::string, double>>& prefixes() {   // include/units.h:27-49
    static const std::vector<std::pair<std::string, double>> p = {
        {"yotta", 1e24}, {"zetta", 1e21}, {"exa", 1e18},   {"peta", 1e15},  {"tera", 1e12},  {"giga", 1e9},
        {"mega", 1e6},   {"kilo", 1e3},   {"hecto", 1e2},  {"deci", 1e-1},  {"centi", 1e-2}, {"milli", 1e-3},
        {"micro", 1e-6}, {"nano", 1e-9},  {"pico", 1e-12}, {"femto", 1e-15}, {"atto", 1e-18}, {"zepto", 1e-21},
        {"yocto", 1e-24}};
    return p;
}
const std::map<std::string, std::map<std::string, double>>& unit_map() {   // include/units.h:52-153
    static const double amu = 1822.8885;
    static const std::map<std::string, std::map<std::string, double>> m = {
        {"undefined", {{"", 1.0}, {"automatic", 1.0}, {"atomic_unit", 1.0}}},
        {"energy", {{"", 1.0}, {"automatic", 1.0}, {"atomic_unit", 1.0}, {"electronvolt", 0.036749326},
                    {"j/mol", 0.00000038087989}, {"cal/mol", 0.0000015946679}, {"kelvin", 3.1668152e-06}}},
        {"temperature", {{"", 1.0}, {"automatic", 1.0}, {"atomic_unit", 1.0}, {"kelvin", 3.1668152e-06}}},
        {"time", {{"", 1.0}, {"automatic", 1.0}, {"atomic_unit", 1.0}, {"second", 4.1341373e16}}},
        {"frequency", {{"", 1.0}, {"automatic", 1.0}, {"atomic_unit", 1.0}, {"inversecm", 4.5563353e-06},
                       {"hertz*rad", 2.4188843e-17}, {"hertz", 1.5198298e-16}}},
        {"length", {{"", 1.0}, {"automatic", 1.0}, {"atomic_unit", 1.0}, {"angstrom", 1.8897261},
                    {"meter", 1.8897261e10}, {"radian", 1.0}, {"degree", 0.017453292519943295}}},
        {"velocity", {{"", 1.0}, {"automatic", 1.0}, {"atomic_unit", 1.0}, {"angstrom/ps", 4.5710289e-5},
                      {"m/s", 4.5710289e-7}}},
        {"momentum", {{"", 1.0}, {"automatic", 1.0}, {"atomic_unit", 1.0}}},
        {"mass", {{"", 1.0}, {"automatic", 1.0}, {"atomic_unit", 1.0}, {"dalton", amu}, {"amu", amu},
                  {"electronmass", 1.0}}},
        {"force", {{"", 1.0}, {"automatic", 1.0}, {"atomic_unit", 1.0}, {"newton", 12137805}, {"ev/ang", 0.019446904}}},
    };
    return m;
}
}  // namespace

double Units::convertToInternal(const std::string& family, const std::string& unit, double number) {
    if (family == "number") return number;
    auto fam = unit_map().find(family);
    if (fam == unit_map().end()) throw std::invalid_argument(family + " is an undefined units kind.");
    // metric prefixes are peeled greedily from the left, like the reference's regex "(p1|p2|...)*(.*)"
    std::string base = unit;
    double scale = 1.0;
    bool peeled = true;
    while (peeled) {
        peeled = false;
        for (const auto& [name, value] : prefixes()) {
            if (base.rfind(name, 0) == 0) {
                base = base.substr(name.size());
                scale = value;   // the reference keeps the last matched prefix only
                peeled = true;
                break;
            }
        }
    }
    auto it = fam->second.find(base);
    if (it == fam->second.end()) throw std::invalid_argument(base + " is an undefined unit for kind " + family + ".");
    return number * it->second * scale;
}

double Units::convertToUser(const std::string& family, const std::string& unit, double number) {
    return number / convertToInternal(family, unit, 1.0);
}

// ====================================================================================== INI
namespace {
std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return s;
}
std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
}  // namespace

IniFile::IniFile(const std::string& filename) {
    std::ifstream in(filename);
    if (!in.is_open()) return;
    ok_ = true;
    std::string line, section;
    bool first = true;
    while (std::getline(in, line)) {
        if (first && line.size() >= 3 && (unsigned char)line[0] == 0xEF) line = line.substr(3);   // UTF-8 BOM
        first = false;
        std::string t = trim(line);
        if (t.empty() || t[0] == ';' || t[0] == '#') continue;
        if (t[0] == '[') {
            size_t e = t.find(']');
            if (e != std::string::npos) section = lower(trim(t.substr(1, e - 1)));
            continue;
        }
        size_t eq = t.find_first_of("=:");
        if (eq == std::string::npos) continue;
        std::string name = lower(trim(t.substr(0, eq)));
        std::string value = t.substr(eq + 1);
        // inline comments: " ;" preceded by whitespace (libs/ini.h:115-119)
        for (size_t i = 1; i < value.size(); ++i)
            if (value[i] == ';' && std::isspace((unsigned char)value[i - 1])) { value = value.substr(0, i); break; }
        values_[section + "=" + name] = trim(value);
    }
}
bool IniFile::Has(const std::string& s, const std::string& n) const { return values_.count(lower(s) + "=" + lower(n)) > 0; }
std::string IniFile::Get(const std::string& s, const std::string& n, const std::string& def) const {
    auto it = values_.find(lower(s) + "=" + lower(n));
    return it == values_.end() ? def : it->second;
}
long IniFile::GetInteger(const std::string& s, const std::string& n, long def) const {
    std::string v = Get(s, n, "");
    char* end = nullptr;
    long r = std::strtol(v.c_str(), &end, 0);
    return end > v.c_str() ? r : def;
}
double IniFile::GetReal(const std::string& s, const std::string& n, double def) const {
    std::string v = Get(s, n, "");
    char* end = nullptr;
    double r = std::strtod(v.c_str(), &end);
    return end > v.c_str() ? r : def;
}
bool IniFile::GetBoolean(const std::string& s, const std::string& n, bool def) const {
    std::string v = lower(Get(s, n, ""));
    if (v == "true" || v == "yes" || v == "on" || v == "1") return true;
    if (v == "false" || v == "no" || v == "off" || v == "0") return false;
    return def;
}

// ====================================================================================== Params
double Params::getQuantity(const std::string& family, const std::string& input) {
    std::istringstream iss(input);
    double value;
    std::string unit;
    if (!(iss >> value >> std::ws >> unit)) throw std::invalid_argument("Invalid input format");
    return Units::convertToInternal(family, unit, value);
}

namespace {
// "token(value)" | "token" (src/params.cpp:310-330)
bool parseTokenParentheses(const std::string& input, std::string& token, std::string& value) {
    std::string t = trim(input);
    size_t open = t.find('(');
    if (open == std::string::npos) {
        token = t; value.clear();
        return !t.empty() && std::all_of(t.begin(), t.end(), [](unsigned char c) { return std::isalnum(c) || c == '_'; });
    }
    if (t.back() != ')') return false;
    token = t.substr(0, open);
    value = t.substr(open + 1, t.size() - open - 2);
    return !token.empty();
}
int potential_id(const std::string& n) {
    if (n == "free") return PIMDB_POT_FREE;
    if (n == "aziz") return PIMDB_POT_AZIZ;
    if (n == "harmonic") return PIMDB_POT_HARMONIC;
    if (n == "dipole") return PIMDB_POT_DIPOLE;
    if (n == "double_well") return PIMDB_POT_DOUBLE_WELL;
    if (n == "cosine") return PIMDB_POT_COSINE;
    return -1;
}
bool in(const std::string& s, std::initializer_list<const char*> l) {
    return std::any_of(l.begin(), l.end(), [&](const char* c) { return s == c; });
}
}  // namespace

Params::Params(const std::string& filename, int ndim) {
    IniFile r(filename);
    if (!r.ok()) throw std::invalid_argument(std::format("Unable to read the configuration file {}", filename));
    const std::string SIM = "simulation", SYS = "system", IP = "interaction_potential", EP = "external_potential";
    pimdb_config& c = cfg;
    c.ndim = ndim;
    c.dt = getQuantity("time", r.Get(SIM, "dt", "1.0 femtosecond"));
    threshold = r.GetReal(SIM, "threshold", 0.1);
    c.gamma = r.GetReal(SIM, "gamma", -1.0);
    if (c.gamma < 0) c.gamma = 1 / (100.0 * c.dt);
    c.nchains = (int)r.GetInteger(SIM, "nchains", 4);
    if (c.nchains < 1)
        throw std::invalid_argument(std::format("The specified number of Nose-Hoover chains ({}) is less than one!", c.nchains));
    steps = static_cast<long>(std::stod(r.Get(SIM, "steps", "1e5")));
    sfreq = r.GetInteger(SIM, "sfreq", 1000);
    c.nbeads = (int)r.GetInteger(SIM, "nbeads", 4);
    if (c.nbeads < 1)
        throw std::invalid_argument(std::format("The specified number of beads ({}) is less than one!", c.nbeads));
    c.seed = static_cast<unsigned int>(std::stod(r.Get(SIM, "seed", "1234")));
    c.bosonic = r.GetBoolean(SIM, "bosonic", false);
    c.fixcom = r.GetBoolean(SIM, "fixcom", true);
    c.pbc = r.GetBoolean(SIM, "pbc", false);
    c.nmthermostat = r.GetBoolean(SIM, "nmthermostat", false);
    if (!parseTokenParentheses(r.Get(SIM, "initial_position", "random"), init_pos_type, init_pos_spec))
        throw std::invalid_argument("The coordinate initialization method format is invalid!");
    if (!in(init_pos_type, {"random", "xyz", "grid"}))
        throw std::invalid_argument(std::format("The specified coordinate initialization method ({}) is not supported!", init_pos_type));
    if (!parseTokenParentheses(r.Get(SIM, "initial_velocity", "random"), init_vel_type, init_vel_spec))
        throw std::invalid_argument("The velocity initialization method format is invalid!");
    if (!in(init_vel_type, {"random", "manual"}))
        throw std::invalid_argument(std::format("The specified velocity initialization method ({}) is not supported!", init_vel_type));
    propagator_type = r.Get(SIM, "propagator", "cartesian");
    if (c.bosonic && propagator_type == "normal_modes")
        throw std::invalid_argument("Normal modes propogation is currently not available for bosons!");
    if (!in(propagator_type, {"cartesian", "normal_modes"}))
        throw std::invalid_argument(std::format("The specified time propagator ({}) is not supported!", propagator_type));
    c.propagator = propagator_type == "normal_modes" ? PIMDB_PROP_NORMAL_MODES : PIMDB_PROP_CARTESIAN;
    thermostat_type = r.Get(SIM, "thermostat", "error");
    if (thermostat_type == "error") throw std::invalid_argument("Thermostat must be specified!");
    if (c.nmthermostat && thermostat_type == "none")
        throw std::invalid_argument("nmthermostat cannot be used in nve ensemble!");
    if (r.Has(SIM, "nchains") && in(thermostat_type, {"none", "langevin"}))
        throw std::invalid_argument("nchains can only be used with Nose-Hoover thermostats!");
    if (thermostat_type == "langevin") c.thermostat = PIMDB_THERMO_LANGEVIN;
    else if (thermostat_type == "none") c.thermostat = PIMDB_THERMO_NONE;
    else if (thermostat_type == "nose_hoover") c.thermostat = PIMDB_THERMO_NOSE_HOOVER;
    else if (thermostat_type == "nose_hoover_np") c.thermostat = PIMDB_THERMO_NOSE_HOOVER_NP;
    else if (thermostat_type == "nose_hoover_np_dim") c.thermostat = PIMDB_THERMO_NOSE_HOOVER_NP_DIM;
    else throw std::invalid_argument(std::format("The specified thermostat ({}) is not supported!", propagator_type));

    c.temperature = getQuantity("temperature", r.Get(SYS, "temperature", "1.0 kelvin"));
    if (c.temperature <= 0.0)
        throw std::invalid_argument(std::format("The specified temperature ({0:4.3f} kelvin) is unphysical!", c.temperature));
    c.natoms = (int)r.GetInteger(SYS, "natoms", 1);
    if (c.natoms < 1)
        throw std::invalid_argument(std::format("The specified number of particles ({}) is smaller than one!", c.natoms));
    c.mass = getQuantity("mass", r.Get(SYS, "mass", "1.0 dalton"));
    if (c.mass <= 0.0) throw std::invalid_argument(std::format("The provided mass ({0:4.3f}) is unphysical!", c.mass));
    c.size = getQuantity("length", r.Get(SYS, "size", "1.0 picometer"));
    if (c.size <= 0.0) throw std::invalid_argument(std::format("The provided system size ({0:4.3f}) is unphysical!", c.size));

    interaction_name = r.Get(IP, "name", "free");
    if (!in(interaction_name, {"aziz", "free", "harmonic", "dipole"}))
        throw std::invalid_argument(std::format("The specified interaction potential ({}) is not supported!", interaction_name));
    c.int_potential = potential_id(interaction_name);
    c.cutoff = getQuantity("length", r.Get(IP, "cutoff", "-1.0 angstrom"));
    if (interaction_name == "free") c.cutoff = 0.0;
    else if (interaction_name == "harmonic") c.int_omega = getQuantity("energy", r.Get(IP, "omega", "1.0 millielectronvolt"));
    else if (interaction_name == "dipole") c.int_strength = r.GetReal(IP, "strength", 1.0);

    external_name = r.Get(EP, "name", "free");
    if (!in(external_name, {"free", "harmonic", "double_well", "cosine"}))
        throw std::invalid_argument(std::format("The specified external potential ({}) is not supported!", external_name));
    c.ext_potential = potential_id(external_name);
    if (external_name == "harmonic") c.ext_omega = getQuantity("energy", r.Get(EP, "omega", "1.0 millielectronvolt"));
    else if (external_name == "double_well") {
        c.ext_strength = getQuantity("energy", r.Get(EP, "strength", "1.0 millielectronvolt"));
        c.ext_location = getQuantity("length", r.Get(EP, "location", "1.0 angstrom"));
    } else if (external_name == "cosine") {
        c.ext_amplitude = getQuantity("energy", r.Get(EP, "amplitude", "1.0 millielectronvolt"));
        c.ext_phase = r.GetReal(EP, "phase", 1.0);
    }
    out_positions = r.Get("output", "positions", "off");
    out_velocities = r.Get("output", "velocities", "off");
    out_forces = r.Get("output", "forces", "off");
    obs_energy = r.Get("observables", "energy", "kelvin");
    obs_classical = r.Get("observables", "classical", "off");
    obs_bosonic = r.Get("observables", "bosonic", "off");
    obs_gsf = r.Get("observables", "gsf", "off");
    c.bead_begin = 0;
    c.bead_end = c.nbeads;
}

// ====================================================================================== plugins -> C ABI
void BosonicExchange::prepare() {
    sim_.check(pimdb_exchange_prepare(sim_.handles.front()), sim_.handles.front());
    if (sim_.handles.size() > 1) sim_.check(pimdb_exchange_prepare(sim_.handles.back()), sim_.handles.back());
}

void BosonicExchange::exteriorSpringForce(std::vector<double>& f, int bead) {
    std::vector<double> all((size_t)sim_.nbeads * sim_.natoms * sim_.ndim);
    sim_.updateForces();
    sim_.pullArray(PIMDB_F_SPRING, all);
    const size_t slab = (size_t)sim_.natoms * sim_.ndim;
    f.assign(all.begin() + bead * slab, all.begin() + (bead + 1) * slab);
}
std::vector<double> BosonicExchange::getV() {
    std::vector<double> v(sim_.natoms + 1);
    sim_.check(pimdb_exchange_get(sim_.handle, PIMDB_EXCH_V, v.data(), v.size()));
    return v;
}
double BosonicExchange::getVn(int n) { return getV().at(n); }
double BosonicExchange::effectivePotential() { return getV().back(); }
double BosonicExchange::primEstimator() {
    // kinetic = sum_beads NDIM N/(2 beta) - sum_{classical links} E/P + primEstimator(); undo the first two terms
    const pimdb_observables& o = sim_.deviceObservables();
    const double classical_links = o.cl_spring - effectivePotential();
    return o.kinetic - sim_.nbeads * (0.5 * sim_.ndim * sim_.natoms / sim_.beta) + classical_links / sim_.nbeads;
}
double BosonicExchange::getDistinctProbability() { return sim_.deviceObservables().prob_dist; }
double BosonicExchange::getLongestProbability() { return sim_.deviceObservables().prob_all; }

void Propagator::step() { sim.forEach(pimdb_propagator_step); }
void Propagator::momentStep() { sim.forEach(pimdb_moment_step); }
void Propagator::coordsStep() { sim.forEach(pimdb_coords_step); }
void Thermostat::step() { sim.forEach(pimdb_thermostat_step); }

// ====================================================================================== observables
void Observable::initialize(const std::vector<std::string>& labels) {
    for (const auto& l : labels) quantities.emplace_back(l, 0.0);
}
void Observable::resetValues() {
    for (auto& kv : quantities) kv.second = 0.0;
}
double& Observable::q(const std::string& label) {
    for (auto& kv : quantities)
        if (kv.first == label) return kv.second;
    throw std::invalid_argument("Unknown observable quantity " + label);
}

EnergyObservable::EnergyObservable(Simulation& s, const std::string& u) : Observable(s, u) {   // energy.cpp:10-19
    if (sim.external_potential_name == "free" && sim.interaction_potential_name == "free") initialize({"kinetic"});
    else if (sim.external_potential_name == "free" || sim.interaction_potential_name == "free")
        initialize({"kinetic", "potential", "virial"});
    else initialize({"kinetic", "potential", "ext_pot", "int_pot", "virial"});
}
void EnergyObservable::calculate() {
    const pimdb_observables& o = sim.deviceObservables();
    auto conv = [&](double v) { return Units::convertToUser("energy", out_unit, v); };
    q("kinetic") = conv(o.kinetic);
    const bool e = sim.external_potential_name != "free", i = sim.interaction_potential_name != "free";
    if (e && i) { q("ext_pot") = conv(o.ext_pot); q("int_pot") = conv(o.int_pot); }
    if (e || i) { q("potential") = conv(o.potential); q("virial") = conv(o.virial); }
}
ClassicalObservable::ClassicalObservable(Simulation& s, const std::string& u) : Observable(s, u) {   // classical.cpp:11-19
    if (sim.thermostat_type.rfind("nose_hoover", 0) == 0) initialize({"temperature", "cl_kinetic", "cl_spring", "nh_energy"});
    else initialize({"temperature", "cl_kinetic", "cl_spring"});
}
void ClassicalObservable::calculate() {
    const pimdb_observables& o = sim.deviceObservables();
    q("temperature") = Units::convertToUser("temperature", "kelvin", o.temperature);
    q("cl_kinetic") = Units::convertToUser("energy", out_unit, o.cl_kinetic);
    q("cl_spring") = Units::convertToUser("energy", out_unit, o.cl_spring);
    if (sim.thermostat_type.rfind("nose_hoover", 0) == 0)
        q("nh_energy") = Units::convertToUser("energy", out_unit, o.nh_energy);   // Thermostat::getAdditionToH summed over beads
}
BosonicObservable::BosonicObservable(Simulation& s, const std::string& u) : Observable(s, u) {
    initialize({"prob_dist", "prob_all"});
}
void BosonicObservable::calculate() {
    const pimdb_observables& o = sim.deviceObservables();
    q("prob_dist") = o.prob_dist;
    q("prob_all") = o.prob_all;
}

GSFActionObservable::GSFActionObservable(Simulation& s, const std::string& u) : Observable(s, u) {   // gsf_action.cpp:9-13
    initialize({"w_gsf", "pot_gsf"});
}
void GSFActionObservable::calculate() {
    const pimdb_observables& o = sim.deviceObservables();
    q("w_gsf") = o.w_gsf;                                                    // dimensionless (gsf_action.cpp:72)
    q("pot_gsf") = Units::convertToUser("energy", out_unit, o.pot_gsf);      // gsf_action.cpp:66
}

ObservablesLogger::ObservablesLogger(const std::string& filename, const std::vector<std::unique_ptr<Observable>>& obs)
    : observables(obs) {
    file.open("output/" + filename, std::ios::out | std::ios::app);
    if (!file.is_open()) throw std::ios_base::failure(std::format("Failed to open {}.", filename));
    file << std::format("{:^16s}", "step");
    for (const auto& o : observables)
        for (const auto& kv : o->quantities) file << std::vformat(" {:^16s}", std::make_format_args(kv.first));
    file << '\n';
}
void ObservablesLogger::log(long step) {
    file << std::format("{:^16.8e}", static_cast<double>(step));
    for (const auto& o : observables)
        for (const auto& kv : o->quantities) file << std::format(" {:^16.8e}", kv.second);
    file << '\n';
    file.flush();
}

// ====================================================================================== state dumps
State::State(Simulation& s, std::string k, long f, const std::string& unit) : sim(s), kind(std::move(k)), freq(f) {
    family = kind == "position" ? "length" : kind;
    try {
        factor = Units::convertToUser(family, unit, 1.0);
    } catch (const std::invalid_argument&) {
        throw std::invalid_argument("Invalid output unit for " + kind + " state.");
    }
}
void State::initialize() {
    const char* pat = kind == "position" ? "output/position_{}.xyz" : kind == "velocity" ? "output/velocity_{}.dat"
                                                                                           : "output/force_{}.dat";
    for (int b = 0; b < sim.nbeads; ++b)
        files.emplace_back(std::vformat(pat, std::make_format_args(b)), std::ios::out | std::ios::app);
}
void State::output(long step) {
    if (step % freq != 0) return;
    const std::vector<double>* src;
    if (kind == "position") { sim.pullCoord(); src = &sim.coord; }
    else if (kind == "velocity") { sim.pullMomenta(); src = &sim.momenta; }
    else { sim.pullForces(); src = &sim.forces; }
    const double scale = kind == "velocity" ? factor / sim.mass : factor;
    const size_t slab = (size_t)sim.natoms * sim.ndim;
    for (int b = 0; b < sim.nbeads; ++b) {
        auto& out = files[b];
        out << std::format("{}\n", sim.natoms) << std::format("Step {}\n", step);
        for (int i = 0; i < sim.natoms; ++i) {
            if (kind == "position") out << "1";
            else out << (i + 1) << " 1";
            for (int a = 0; a < sim.ndim; ++a)
                out << std::format(" {:^20.12e}", (*src)[b * slab + (size_t)i * sim.ndim + a] * scale);
            if (sim.ndim == 1) out << " 0.0 0.0";
            else if (sim.ndim == 2) out << " 0.0";
            out << "\n";
        }
    }
}

// ====================================================================================== Simulation
void Simulation::check(int rc, pimdb_sim* h) const {
    if (rc == PIMDB_OK) return;
    const char* m = pimdb_last_error(h ? h : handle);
    std::string msg = m ? m : "pimdb error";
    if (rc == PIMDB_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    if (rc == PIMDB_ERR_OVERFLOW) throw std::overflow_error(msg);
    throw std::runtime_error(msg);
}

Simulation::Simulation(Params& p, int device, int ngpus) : params(p) {
    pimdb_config& c = p.cfg;
    c.device = device;
    natoms = c.natoms; nbeads = c.nbeads; ndim = c.ndim;
    steps = p.steps; sfreq = p.sfreq; threshold = p.threshold * p.steps;
    dt = c.dt; mass = c.mass; temperature = c.temperature; size = c.size;
    beta = 1.0 / temperature;
    bosonic = c.bosonic && nbeads > 1; fixcom = c.fixcom; pbc = c.pbc;
    external_potential_name = p.external_name;
    interaction_potential_name = p.interaction_name;
    thermostat_type = p.thermostat_type;
    propagator_type = p.propagator_type;
    if (ngpus < 1 || ngpus > nbeads) throw std::invalid_argument("--gpus must be between 1 and the number of beads");
    // contiguous, balanced bead ranges: the first nbeads % ngpus shards own one bead more
    for (int g = 0; g <= ngpus; ++g) bead_begin.push_back(g * (nbeads / ngpus) + std::min(g, nbeads % ngpus));
    for (int g = 0; g < ngpus; ++g) {
        pimdb_config cg = c;
        // PIMDB_SHARD_SAME_DEVICE=1: every shard on `device` (tests on a one-GPU box; the shards still talk through their mailboxes)
        cg.device = std::getenv("PIMDB_SHARD_SAME_DEVICE") ? device : device + g;
        cg.bead_begin = bead_begin[g];
        cg.bead_end = bead_begin[g + 1];
        pimdb_sim* h = nullptr;
        const int rc = pimdb_create(&cg, &h);
        if (rc != PIMDB_OK) {
            std::string msg = pimdb_last_error(nullptr);
            for (pimdb_sim* made : handles) pimdb_destroy(made);
            handles.clear();
            if (rc == PIMDB_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
            throw std::runtime_error(msg);
        }
        handles.push_back(h);
    }
    handle = handles.front();
    if (ngpus > 1) {   // one process drives every shard: the blobs are simply concatenated
        std::vector<char> blobs((size_t)ngpus * PIMDB_PEER_BLOB_BYTES);
        for (int g = 0; g < ngpus; ++g) check(pimdb_peer_export(handles[g], blobs.data() + (size_t)g * PIMDB_PEER_BLOB_BYTES), handles[g]);
        for (int g = 0; g < ngpus; ++g) check(pimdb_peer_attach(handles[g], ngpus, g, blobs.data()), handles[g]);
    }
    ext_potential = std::make_unique<Potential>(external_potential_name, c.ext_potential);
    int_potential = std::make_unique<Potential>(interaction_potential_name, c.int_potential);
    propagator = std::make_unique<Propagator>(*this);
    thermostat = std::make_unique<Thermostat>(*this);
    if (bosonic) bosonic_exchange = std::make_unique<BosonicExchange>(*this);
    const size_t n = (size_t)nbeads * natoms * ndim;
    coord.assign(n, 0.0); momenta.assign(n, 0.0); forces.assign(n, 0.0);
    initializePositions(p);
    initializeMomenta(p);
    pushCoord();
    pushMomenta();
    // src/simulation.cpp:780-818
    auto add_state = [&](const std::string& units, const std::string& name) {
        if (units == "off" || units == "false") return;
        const std::string u = (units == "on" || units == "true" || units == "none") ? "atomic_unit" : units;
        states.push_back(std::make_unique<State>(*this, name, sfreq, u));
    };
    add_state(p.out_positions, "position");
    add_state(p.out_velocities, "velocity");
    add_state(p.out_forces, "force");
    auto unit_of = [](const std::string& u) { return u == "none" ? std::string("") : u; };
    if (p.obs_energy != "off") observables.push_back(std::make_unique<EnergyObservable>(*this, unit_of(p.obs_energy)));
    if (p.obs_classical != "off") observables.push_back(std::make_unique<ClassicalObservable>(*this, unit_of(p.obs_classical)));
    if (bosonic && p.obs_bosonic != "off") observables.push_back(std::make_unique<BosonicObservable>(*this, unit_of(p.obs_bosonic)));
    if (p.obs_gsf != "off") {
        // With an interaction potential the reference's GSF observable adds a one-row gradient to an N-row array and
        // reads past its end (src/observables/gsf_action.cpp:36): there is nothing well-defined to reproduce.
        if (interaction_potential_name != "free")
            throw std::invalid_argument("The gsf observable is only supported with a free interaction potential");
        observables.push_back(std::make_unique<GSFActionObservable>(*this, unit_of(p.obs_gsf)));
    }
}

Simulation::~Simulation() {
    for (pimdb_sim* h : handles) pimdb_destroy(h);
}

// Every entry point but the reads only enqueues work, so one host thread keeps all shards busy: the call goes to every
// handle in turn and the devices sort out the hand-shakes among themselves.
void Simulation::forEach(int (*fn)(pimdb_sim*)) {
    for (pimdb_sim* h : handles) check(fn(h), h);
}
void Simulation::pushArray(int which, const std::vector<double>& host) {
    const size_t slab = (size_t)natoms * ndim;
    for (size_t g = 0; g < handles.size(); ++g) check(pimdb_set_state(handles[g], which, host.data() + bead_begin[g] * slab), handles[g]);
}
void Simulation::pullArray(int which, std::vector<double>& host) {
    const size_t slab = (size_t)natoms * ndim;
    forEach(pimdb_settle);   // the reads block and their deferred part is collective: enqueue it everywhere first
    for (size_t g = 0; g < handles.size(); ++g) check(pimdb_get_state(handles[g], which, host.data() + bead_begin[g] * slab), handles[g]);
}
void Simulation::pushCoord() { pushArray(PIMDB_X, coord); }
void Simulation::pushMomenta() { pushArray(PIMDB_P, momenta); }
void Simulation::pullCoord() { pullArray(PIMDB_X, coord); }
void Simulation::pullMomenta() { pullArray(PIMDB_P, momenta); }
void Simulation::pullForces() { pullArray(PIMDB_F, forces); }
void Simulation::updateForces() { forEach(pimdb_update_forces); }
void Simulation::updateNeighboringCoordinates() { forEach(pimdb_update_neighbors); }
void Simulation::zeroMomentum() {
    if (handles.size() == 1) { check(pimdb_zero_momentum(handle)); return; }
    throw std::runtime_error("zeroMomentum as a separate call needs all beads on one GPU (sharded runs fold it into the step)");
}

// ObservablesLogger::log sums the per-rank values with MPI_Allreduce (src/observables/observable.cpp:92-116): here the
// per-shard partial structs are added field by field.
const pimdb_observables& Simulation::deviceObservables() {
    if (obs_step != md_step) {
        forEach(pimdb_settle);
        pimdb_observables total{};
        static_assert(sizeof(pimdb_observables) % sizeof(double) == 0, "a struct of doubles");
        for (pimdb_sim* h : handles) {
            pimdb_observables part{};
            check(pimdb_observables_calc(h, &part), h);
            for (size_t i = 0; i < sizeof(pimdb_observables) / sizeof(double); ++i)
                reinterpret_cast<double*>(&total)[i] += reinterpret_cast<const double*>(&part)[i];
        }
        obs_cache = total;
        obs_step = md_step;
    }
    return obs_cache;
}

// src/simulation.cpp:708-728 with genRandomPositions :118-126 and uniformParticleGrid :133-193; one generator per bead
// seeded seed+bead, positions first, then momenta from the same generator (:58, :76-77).
void Simulation::initializePositions(const Params& p) {
    const size_t slab = (size_t)natoms * ndim;
    const double ang = Units::convertToInternal("length", "angstrom", 1.0);
    if (p.init_pos_type == "xyz") {
        std::string fmt = p.init_pos_spec;
        int zero = 0;
        const bool formatted = std::vformat(fmt, std::make_format_args(zero)) != fmt;
        const int first = formatted ? (std::filesystem::exists(std::vformat(fmt, std::make_format_args(zero))) ? 0 : 1) : 0;
        for (int b = 0; b < nbeads; ++b) {
            int arg = b + first;
            const std::string name = formatted ? std::vformat(fmt, std::make_format_args(arg)) : fmt;
            std::ifstream in(name);
            if (!in.is_open()) throw std::runtime_error(std::format("Cannot open the xyz file named {}.", name));
            int n = 0;
            in >> n;
            if (n != natoms)
                throw std::runtime_error(std::format(
                    "The number of atoms in the xyz file ({}) does not match the requested number of atoms.", name));
            std::string line;
            std::getline(in, line);
            std::getline(in, line);
            for (int i = 0; i < natoms; ++i) {
                std::string symbol;
                in >> symbol;
                for (int a = 0; a < ndim; ++a) {
                    double v;
                    in >> v;
                    coord[b * slab + (size_t)i * ndim + a] = v * ang;
                }
            }
        }
    } else if (p.init_pos_type == "grid") {
        const double EPS = 1.0e-7;
        const double volume = std::pow(size, ndim);
        const double init_side = std::pow((1.0 * natoms / volume), -1.0 / (1.0 * ndim));
        std::array<int, 3> num{1, 1, 1};
        std::array<double, 3> cell{0, 0, 0};
        int total = 1;
        for (int i = 0; i < ndim; ++i) {
            num[i] = std::max(static_cast<int>(std::ceil((size / init_side) - EPS)), 1);
            cell[i] = size / (1.0 * num[i]);
            total *= num[i];
        }
        if (total < natoms) throw std::runtime_error("Number of grid boxes is less than the number of particles");
        for (int n = 0; n < total && n < natoms; ++n) {
            for (int i = 0; i < ndim; ++i) {
                int scale = 1;
                for (int j = i + 1; j < ndim; ++j) scale *= num[j];
                const int gi = (n / scale) % num[i];
                double pos = (gi + 0.5) * cell[i] - 0.5 * size;
                if (pbc) pos -= size * std::floor(pos / size + 0.5);
                for (int b = 0; b < nbeads; ++b) coord[b * slab + (size_t)n * ndim + i] = pos;
            }
        }
    } else {
        for (int b = 0; b < nbeads; ++b) {
            std::mt19937 gen((unsigned)(p.cfg.seed + b));
            std::uniform_real_distribution<double> u(-0.5 * size, 0.5 * size);
            for (size_t q = 0; q < slab; ++q) coord[b * slab + q] = u(gen);
        }
    }
}

void Simulation::initializeMomenta(const Params& p) {
    const size_t slab = (size_t)natoms * ndim;
    if (p.init_vel_type == "manual") {
        const double vel = Units::convertToInternal("velocity", "angstrom/ps", 1.0);
        std::string fmt = p.init_vel_spec.empty() ? std::string("init/vel_{:02}.dat") : p.init_vel_spec;
        int zero = 0;
        int first = 1;   // LAMMPS convention for the default file names (:739)
        if (!p.init_vel_spec.empty())
            first = std::filesystem::exists(std::vformat(fmt, std::make_format_args(zero))) ? 0 : 1;
        for (int b = 0; b < nbeads; ++b) {
            int arg = b + first;
            const std::string name = std::vformat(fmt, std::make_format_args(arg));
            std::ifstream in(name);
            if (!in.is_open()) throw std::runtime_error(std::format("Cannot open the velocity file named {}.", name));
            int n = 0;
            in >> n;
            if (n != natoms)
                throw std::runtime_error(std::format(
                    "The number of atoms in the velocity file ({}) does not match the requested number of atoms.", name));
            std::string line;
            std::getline(in, line);
            std::getline(in, line);
            for (int i = 0; i < natoms; ++i) {
                std::string tok;
                in >> tok >> tok;
                for (int a = 0; a < ndim; ++a) {
                    double v;
                    in >> v;
                    momenta[b * slab + (size_t)i * ndim + a] = mass * (v * vel);
                }
            }
        }
        return;   // no COM removal after loading (:737-742)
    }
    // genMomentum :200-217 — the generator continues after the positions when those were random too
    const double thermo_beta = beta / nbeads;
    for (int b = 0; b < nbeads; ++b) {
        std::mt19937 gen((unsigned)(p.cfg.seed + b));
        if (p.init_pos_type == "random") {
            std::uniform_real_distribution<double> u(-0.5 * size, 0.5 * size);
            for (size_t q = 0; q < slab; ++q) (void)u(gen);
        }
        for (size_t q = 0; q < slab; ++q) {
            std::normal_distribution<double> normal(0.0, 1 / std::sqrt(thermo_beta * mass));   // fresh per sample (:213-217)
            momenta[b * slab + q] = mass * normal(gen);
        }
    }
    // zeroMomentum after random generation, whatever fixcom says (:750-751); same summation order as the reference
    std::array<double, 3> cm{0, 0, 0};
    for (int b = 0; b < nbeads; ++b) {
        std::array<double, 3> part{0, 0, 0};
        for (int i = 0; i < natoms; ++i)
            for (int a = 0; a < ndim; ++a) part[a] += momenta[b * slab + (size_t)i * ndim + a];
        for (int a = 0; a < ndim; ++a) cm[a] += part[a] / (natoms * nbeads);
    }
    for (int b = 0; b < nbeads; ++b)
        for (int i = 0; i < natoms; ++i)
            for (int a = 0; a < ndim; ++a) momenta[b * slab + (size_t)i * ndim + a] -= cm[a];
}

// Simulation::run, src/simulation.cpp:222-290. The device advances whole batches of steps between the iterations
// that dump state or log observables (the reference evaluates the observables every step but prints every sfreq).
void Simulation::run() {
    std::cout << "[*] Running the simulation\n";
    std::filesystem::create_directory("output");
    ObservablesLogger logger("simulation.out", observables);
    for (auto& s : states) s->initialize();
    const auto t0 = std::chrono::steady_clock::now();
    long step = 0;
    while (step <= steps) {
        md_step = step;
        const bool event = (step % sfreq == 0);
        if (event) {
            for (auto& o : observables) o->resetValues();
            for (auto& s : states) s->output(step);
            for (pimdb_sim* h : handles) check(pimdb_step(h, 1), h);
            md_step = step + 1;   // observables follow the update of this iteration (App. A-3)
            if (!(step < threshold)) {
                for (auto& o : observables) o->calculate();
                logger.log(step);
            }
            ++step;
        } else {
            const long next_event = std::min(steps + 1, (step / sfreq + 1) * sfreq);
            for (pimdb_sim* h : handles) check(pimdb_step(h, (int)(next_event - step)), h);
            step = next_event;
        }
    }
    forEach(pimdb_synchronize);
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << std::format("[*] Simulation finished running successfully (Runtime = {:.3} sec)\n", wall);
    printReport(wall);
}

void Simulation::printReport(double wall_time) const {
    std::ofstream rep("output/report.txt", std::ios::out | std::ios::app);
    auto line = [&](const std::string& k, const auto& v) { rep << std::format("{:<40}\t:\t{}\n", k, v); };
    rep << "---------\nParameters\n---------\n";
    if (bosonic) { line("Statistics", "Bosonic"); line("Bosonic algorithm", params.cfg.exchange_alg == PIMDB_EXCH_FACTORIAL ? "Naive" : "Feldman-Hirshberg"); }
    else line("Statistics", "Boltzmannonic");
    line("Time propagation algorithm", propagator_type);
    line("Periodic boundary conditions", pbc);
    line("Dimension", ndim);
    line("Seed", params.cfg.seed);
    line("Coordinate initialization method", params.init_pos_type);
    line("Number of atoms", natoms);
    line("Number of beads", nbeads);
    line("Temperature", std::format("{} kelvin", Units::convertToUser("temperature", "kelvin", temperature)));
    line("Linear size of the system", std::format("{} angstroms", Units::convertToUser("length", "angstrom", size)));
    line("Mass", std::format("{} amu", Units::convertToUser("mass", "dalton", mass)));
    line("Total number of MD steps", steps);
    line("Interaction potential name", interaction_potential_name);
    line("External potential name", external_potential_name);
    rep << "---------\nFeatures\n---------\n";
    line("Minimum image convention", true);
    line("Wrapping of coordinates", true);
    line("Using i-Pi convention", true);
    line("Device path", "libpimdb200 (sm_100a)");
    line("GPUs (bead shards)", handles.size());
    rep << "---------\n";
    line("Wall time (sec)", std::format("{:.3f}", wall_time));
    line("Wall time per step (sec)", std::format("{:.5e}", wall_time / steps));
}

}  // namespace pimdb_host
