"""Host-side mirror of the reference's ``Params`` (INI schema, units, defaults, validation).

The reference reads ``config.ini`` with inih and converts every ``"<number> <unit>"`` quantity to atomic
units (reference: src/params.cpp:8-260, src/units.cpp:42-64, include/units.h:27-153). This module keeps
that schema verbatim -- same sections, keys, defaults, unit names, metric prefixes and error messages --
so an existing ``config.ini`` drives the B200 path unchanged.  It does no numerics.
"""
from __future__ import annotations

import configparser
import re
from dataclasses import dataclass, field, asdict
from typing import Dict, Optional

# include/units.h:15-22
KB = 1.0
HBAR = 1.0
AMU = 1822.8885

# include/units.h:27-49
UNIT_PREFIX: Dict[str, float] = {
    "": 1.0, "yotta": 1e24, "zetta": 1e21, "exa": 1e18, "peta": 1e15, "tera": 1e12, "giga": 1e9,
    "mega": 1e6, "kilo": 1e3, "hecto": 1e2, "deci": 1e-1, "centi": 1e-2, "milli": 1e-3, "micro": 1e-6,
    "nano": 1e-9, "pico": 1e-12, "femto": 1e-15, "atto": 1e-18, "zepto": 1e-21, "yocto": 1e-24,
}

_COMMON = {"": 1.0, "automatic": 1.0, "atomic_unit": 1.0}

# include/units.h:52-153 (only the families the INI schema and the writers use)
UNIT_MAP: Dict[str, Dict[str, float]] = {
    "undefined": dict(_COMMON),
    "energy": {**_COMMON, "electronvolt": 0.036749326, "j/mol": 0.00000038087989,
               "cal/mol": 0.0000015946679, "kelvin": 3.1668152e-06},
    "temperature": {**_COMMON, "kelvin": 3.1668152e-06},
    "time": {**_COMMON, "second": 4.1341373e16},
    "frequency": {**_COMMON, "inversecm": 4.5563353e-06, "hertz*rad": 2.4188843e-17, "hertz": 1.5198298e-16},
    "length": {**_COMMON, "angstrom": 1.8897261, "meter": 1.8897261e10, "radian": 1.0,
               "degree": 0.017453292519943295},
    "velocity": {**_COMMON, "angstrom/ps": 4.5710289e-5, "m/s": 4.5710289e-7},
    "momentum": dict(_COMMON),
    "mass": {**_COMMON, "dalton": AMU, "amu": AMU, "electronmass": 1.0},
    "force": {**_COMMON, "newton": 12137805, "ev/ang": 0.019446904},
}

_PREFIX_RE = re.compile("(" + "|".join(k for k in UNIT_PREFIX if k) + ")*(.*)")


def separate_prefix_unit(unit: str):
    """src/units.cpp:17-40."""
    m = _PREFIX_RE.fullmatch(unit)
    if m:
        return (m.group(1) or ""), m.group(2)
    return "", unit


def convert_to_internal(family: str, unit: str, number: float) -> float:
    """src/units.cpp:42-60 (same error texts)."""
    if family == "number":
        return number
    if family not in UNIT_MAP:
        raise ValueError(f"{family} is an undefined units kind.")
    prefix, base = separate_prefix_unit(unit)
    if base not in UNIT_MAP[family]:
        raise ValueError(f"{base} is an undefined unit for kind {family}.")
    return number * UNIT_MAP[family][base] * UNIT_PREFIX[prefix]


def convert_to_user(family: str, unit: str, number: float) -> float:
    """src/units.cpp:62-64."""
    return number / convert_to_internal(family, unit, 1.0)


def get_quantity(family: str, text: str) -> float:
    """``"<number> <unit>"`` -> atomic units (src/params.cpp:281-308)."""
    parts = text.split()
    if len(parts) < 2:
        raise ValueError("Invalid input format")
    return convert_to_internal(family, parts[1], float(parts[0]))


INT_POTENTIALS = ("free", "aziz", "harmonic", "dipole")           # src/params.cpp:194
EXT_POTENTIALS = ("free", "harmonic", "double_well", "cosine")     # src/params.cpp:195
PROPAGATORS = ("cartesian", "normal_modes")                        # src/params.cpp:144
THERMOSTATS = ("langevin", "nose_hoover", "nose_hoover_np", "nose_hoover_np_dim", "none")  # :160


@dataclass
class SimConfig:
    """All quantities in atomic units (what ``Params`` holds after parsing)."""
    # [simulation]
    dt: float = 1.0e-15 * 4.1341373e16
    threshold: float = 0.1
    gamma: float = -1.0          # <0 -> 1/(100 dt)   (src/params.cpp:17-20)
    nchains: int = 4
    steps: int = 100000
    sfreq: int = 1000
    nbeads: int = 4
    seed: int = 1234
    bosonic: bool = False
    fixcom: bool = True
    pbc: bool = False
    nmthermostat: bool = False
    initial_position: str = "random"
    initial_velocity: str = "random"
    propagator: str = "cartesian"
    thermostat: str = "langevin"
    # [system]
    temperature: float = 3.1668152e-06
    natoms: int = 1
    mass: float = AMU
    size: float = 1.0e-12 * 1.8897261e10
    ndim: int = 3                # compile-time NDIM in the reference (CMakeLists.txt:44-48)
    # [interaction_potential]
    interaction: str = "free"
    cutoff: float = -1.0 * 1.8897261
    int_omega: float = 0.0
    int_strength: float = 1.0
    # [external_potential]
    external: str = "free"
    ext_omega: float = 0.0
    ext_strength: float = 0.0
    ext_location: float = 0.0
    ext_amplitude: float = 0.0
    ext_phase: float = 1.0
    # [output] / [observables]
    out_positions: str = "off"
    out_velocities: str = "off"
    out_forces: str = "off"
    obs_energy: str = "kelvin"
    obs_classical: str = "off"
    obs_bosonic: str = "off"
    obs_gsf: str = "off"
    # not an INI key: noise stream of the Langevin thermostat on the GPU, "philox" (parallel, default) or "ranmars"
    # (the reference's own generator, one sequential stream per bead: trajectory-level parity, slow)
    rng: str = "philox"
    # not an INI key either: the reference chooses its exchange class at compile time (CMakeLists.txt:50-54,
    # -DFACTORIAL_BOSONIC_ALGORITHM): "quadratic" (Feldman-Hirshberg, default) or "factorial" (all N! permutations, N <= 10)
    exchange_alg: str = "quadratic"

    def __post_init__(self):
        if self.gamma < 0:
            self.gamma = 1.0 / (100.0 * self.dt)

    # ---- derived constants (src/simulation.cpp:37-52, 84-90) ----
    @property
    def beta(self) -> float:
        return 1.0 / (KB * self.temperature)

    @property
    def thermo_beta(self) -> float:
        return self.beta / self.nbeads

    @property
    def omega_p(self) -> float:
        return self.nbeads / (self.beta * HBAR)

    @property
    def spring_constant(self) -> float:
        return self.mass * self.omega_p * self.omega_p

    @property
    def cutoff_effective(self) -> float:
        rc = 0.0 if self.interaction == "free" else self.cutoff
        if self.pbc:
            rc = min(rc, 0.5 * self.size)
        return rc

    @property
    def bosonic_active(self) -> bool:
        return bool(self.bosonic and self.nbeads > 1)   # src/simulation.cpp:690

    def validate(self) -> None:
        """The checks of src/params.cpp with the same messages (raised as ValueError ~ std::invalid_argument)."""
        if self.nchains < 1:
            raise ValueError(f"The specified number of Nose-Hoover chains ({self.nchains}) is less than one!")
        if self.nbeads < 1:
            raise ValueError(f"The specified number of beads ({self.nbeads}) is less than one!")
        if self.propagator not in PROPAGATORS:
            raise ValueError(f"The specified time propagator ({self.propagator}) is not supported!")
        if self.bosonic and self.propagator == "normal_modes":
            raise ValueError("Normal modes propogation is currently not available for bosons!")
        if self.thermostat not in THERMOSTATS:
            raise ValueError(f"The specified thermostat ({self.propagator}) is not supported!")
        if self.nmthermostat and self.thermostat == "none":
            raise ValueError("nmthermostat cannot be used in nve ensemble!")
        if self.temperature <= 0.0:
            raise ValueError(f"The specified temperature ({self.temperature:4.3f} kelvin) is unphysical!")
        if self.natoms < 1:
            raise ValueError(f"The specified number of particles ({self.natoms}) is smaller than one!")
        if self.mass <= 0.0:
            raise ValueError(f"The provided mass ({self.mass:4.3f}) is unphysical!")
        if self.size <= 0.0:
            raise ValueError(f"The provided system size ({self.size:4.3f}) is unphysical!")
        if self.interaction not in INT_POTENTIALS:
            raise ValueError(f"The specified interaction potential ({self.interaction}) is not supported!")
        if self.external not in EXT_POTENTIALS:
            raise ValueError(f"The specified external potential ({self.external}) is not supported!")
        if self.ndim not in (1, 2, 3):
            raise ValueError(f"NDIM must be 1, 2 or 3 (got {self.ndim})")

    # ---- INI round trip ----
    def to_ini(self, energy_unit: str = "atomic_unit") -> str:
        """An INI the *reference* parses to exactly these values (everything in ``atomic_unit``, %.17g)."""
        g = lambda v: f"{v:.17g}"
        b = lambda v: "true" if v else "false"
        lines = [
            "[simulation]",
            f"dt = {g(self.dt)} atomic_unit",
            f"steps = {self.steps}",
            f"sfreq = {self.sfreq}",
            f"threshold = {g(self.threshold)}",
            f"gamma = {g(self.gamma)}",
            f"nbeads = {self.nbeads}",
            f"seed = {self.seed}",
            f"bosonic = {b(self.bosonic)}",
            f"fixcom = {b(self.fixcom)}",
            f"pbc = {b(self.pbc)}",
            f"nmthermostat = {b(self.nmthermostat)}",
            f"initial_position = {self.initial_position}",
            f"initial_velocity = {self.initial_velocity}",
            f"propagator = {self.propagator}",
            f"thermostat = {self.thermostat}",
        ]
        if self.thermostat.startswith("nose_hoover"):
            lines.append(f"nchains = {self.nchains}")
        lines += [
            "[system]",
            f"temperature = {g(self.temperature)} atomic_unit",
            f"natoms = {self.natoms}",
            f"mass = {g(self.mass)} atomic_unit",
            f"size = {g(self.size)} atomic_unit",
            "[interaction_potential]",
            f"name = {self.interaction}",
            f"cutoff = {g(self.cutoff)} atomic_unit",
        ]
        if self.interaction == "harmonic":
            lines.append(f"omega = {g(self.int_omega)} atomic_unit")
        if self.interaction == "dipole":
            lines.append(f"strength = {g(self.int_strength)}")
        lines += ["[external_potential]", f"name = {self.external}"]
        if self.external == "harmonic":
            lines.append(f"omega = {g(self.ext_omega)} atomic_unit")
        if self.external == "double_well":
            lines += [f"strength = {g(self.ext_strength)} atomic_unit", f"location = {g(self.ext_location)} atomic_unit"]
        if self.external == "cosine":
            lines += [f"amplitude = {g(self.ext_amplitude)} atomic_unit", f"phase = {g(self.ext_phase)}"]
        lines += [
            "[output]",
            f"positions = {self.out_positions}",
            f"velocities = {self.out_velocities}",
            f"forces = {self.out_forces}",
            "[observables]",
            f"energy = {energy_unit if self.obs_energy != 'off' else 'off'}",
            f"classical = {energy_unit if self.obs_classical != 'off' else 'off'}",
            f"bosonic = {self.obs_bosonic}",
            f"gsf = {self.obs_gsf}",
        ]
        return "\n".join(lines) + "\n"

    def as_dict(self):
        return asdict(self)


_TRUE = {"true", "yes", "on", "1"}
_FALSE = {"false", "no", "off", "0"}


def _atoi(text: str) -> int:
    """C ``atoi`` semantics (libs/inireader.cpp GetInteger uses strtol): leading integer part, else 0."""
    m = re.match(r"\s*([+-]?\d+)", text)
    return int(m.group(1)) if m else 0


def parse_ini(path: str, ndim: int = 3) -> SimConfig:
    """Parse a reference ``config.ini`` (src/params.cpp:8-260).

    inih behaviours the goldens rely on are kept: case-insensitive sections/keys, inline ``;`` comments,
    ``nbeads = 8.0`` -> 8, ``sfreq = 1e3`` -> 1, unknown keys ignored (libs/inireader.cpp:49-110).
    """
    cp = configparser.ConfigParser(inline_comment_prefixes=(";",), comment_prefixes=(";", "#"),
                                   strict=False, interpolation=None)
    with open(path, "r", encoding="utf-8-sig") as fh:
        cp.read_file(fh)
    sec = {s.lower(): {k.lower(): v.strip() for k, v in cp.items(s)} for s in cp.sections()}

    def get(section, key, default):
        return sec.get(section, {}).get(key, default)

    def get_bool(section, key, default):
        v = get(section, key, None)
        if v is None:
            return default
        v = v.lower()
        if v in _TRUE:
            return True
        if v in _FALSE:
            return False
        return default

    def get_real(section, key, default):
        v = get(section, key, None)
        if v is None:
            return default
        m = re.match(r"\s*([+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?))", v)
        return float(m.group(1)) if m else default

    c = SimConfig.__new__(SimConfig)
    d = SimConfig()  # defaults
    c.__dict__.update(d.__dict__)
    c.ndim = ndim
    S = "simulation"
    c.dt = get_quantity("time", get(S, "dt", "1.0 femtosecond"))
    c.threshold = get_real(S, "threshold", 0.1)
    c.gamma = get_real(S, "gamma", -1.0)
    if c.gamma < 0:
        c.gamma = 1.0 / (100.0 * c.dt)
    raw_nchains = get(S, "nchains", None)
    c.nchains = _atoi(raw_nchains) if raw_nchains is not None else 4
    c.steps = int(float(get(S, "steps", "1e5")))
    c.sfreq = _atoi(get(S, "sfreq", "1000"))
    c.nbeads = _atoi(get(S, "nbeads", "4"))
    c.seed = int(float(get(S, "seed", "1234")))
    c.bosonic = get_bool(S, "bosonic", False)
    c.fixcom = get_bool(S, "fixcom", True)
    c.pbc = get_bool(S, "pbc", False)
    c.nmthermostat = get_bool(S, "nmthermostat", False)
    c.initial_position = get(S, "initial_position", "random")
    c.initial_velocity = get(S, "initial_velocity", "random")
    c.propagator = get(S, "propagator", "cartesian")
    c.thermostat = get(S, "thermostat", "error")
    if c.thermostat == "error":
        raise ValueError("Thermostat must be specified!")
    if raw_nchains is not None and c.thermostat in ("none", "langevin"):
        raise ValueError("nchains can only be used with Nose-Hoover thermostats!")
    Y = "system"
    c.temperature = get_quantity("temperature", get(Y, "temperature", "1.0 kelvin"))
    c.natoms = _atoi(get(Y, "natoms", "1"))
    c.mass = get_quantity("mass", get(Y, "mass", "1.0 dalton"))
    c.size = get_quantity("length", get(Y, "size", "1.0 picometer"))
    I = "interaction_potential"
    c.interaction = get(I, "name", "free")
    c.cutoff = get_quantity("length", get(I, "cutoff", "-1.0 angstrom"))
    if c.interaction == "free":
        c.cutoff = 0.0
    elif c.interaction == "harmonic":
        c.int_omega = get_quantity("energy", get(I, "omega", "1.0 millielectronvolt"))
    elif c.interaction == "dipole":
        c.int_strength = get_real(I, "strength", 1.0)
    E = "external_potential"
    c.external = get(E, "name", "free")
    if c.external == "harmonic":
        c.ext_omega = get_quantity("energy", get(E, "omega", "1.0 millielectronvolt"))
    elif c.external == "double_well":
        c.ext_strength = get_quantity("energy", get(E, "strength", "1.0 millielectronvolt"))
        c.ext_location = get_quantity("length", get(E, "location", "1.0 angstrom"))
    elif c.external == "cosine":
        c.ext_amplitude = get_quantity("energy", get(E, "amplitude", "1.0 millielectronvolt"))
        c.ext_phase = get_real(E, "phase", 1.0)
    c.out_positions = get("output", "positions", "off")
    c.out_velocities = get("output", "velocities", "off")
    c.out_forces = get("output", "forces", "off")
    c.obs_energy = get("observables", "energy", "kelvin")
    c.obs_classical = get("observables", "classical", "off")
    c.obs_bosonic = get("observables", "bosonic", "off")
    c.obs_gsf = get("observables", "gsf", "off")
    c.validate()
    return c
