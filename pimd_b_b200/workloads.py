"""The BASELINE.json configurations as SimConfig objects plus seeded synthetic initial states.

All states come from ``numpy.random.default_rng(seed)``, are FP64, and are fed to both implementations (the GPU
path directly, the reference through ``xyz(...)`` / ``manual(...)`` files written by pimd_b_b200.io).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .config import SimConfig

KELVIN = 3.1668152e-06
ANGSTROM = 1.8897261
DALTON = 1822.8885
FEMTOSECOND = 1.0e-15 * 4.1341373e16
MEV = 1.0e-3 * 0.036749326
HE4_DENSITY = 0.02186        # atoms / angstrom^3 (L = 28.61 A at N = 512)


def helium_box(natoms: int) -> float:
    return (natoms / HE4_DENSITY) ** (1.0 / 3.0) * ANGSTROM


def config(name: str) -> SimConfig:
    name = name.lower()
    if name == "c1":   # 3-D harmonic trap, N=16 bosons, P=32, free interaction, cartesian + Langevin
        return SimConfig(nbeads=32, natoms=16, ndim=3, bosonic=True, fixcom=False, pbc=False,
                         temperature=5.802 * KELVIN, mass=1.0, size=300.0, interaction="free",
                         external="harmonic", ext_omega=3 * MEV, thermostat="langevin", propagator="cartesian",
                         seed=4242, dt=FEMTOSECOND, obs_classical="kelvin", obs_bosonic="true")
    if name == "c2":   # 2-D dipolar particles in a trap, N=64, P=64, NM propagator + NM thermostat
        # the stock reference rejects bosonic + normal_modes (src/params.cpp:146-148): distinguishable, as in SURVEY 8d
        return SimConfig(nbeads=64, natoms=64, ndim=2, bosonic=False, fixcom=False, pbc=False,
                         temperature=5 * KELVIN, mass=1.0, size=200.0, interaction="dipole", int_strength=1.0,
                         external="harmonic", ext_omega=3 * MEV, thermostat="langevin", nmthermostat=True,
                         propagator="normal_modes", seed=777, dt=FEMTOSECOND, obs_classical="kelvin")
    if name in ("c3", "c4"):   # liquid He-4, Aziz, PBC, bosons (c3 = headline)
        n, p = (512, 64) if name == "c3" else (2048, 128)
        return SimConfig(nbeads=p, natoms=n, ndim=3, bosonic=True, fixcom=True, pbc=True,
                         temperature=2 * KELVIN, mass=4.0026 * DALTON, size=helium_box(n), interaction="aziz",
                         cutoff=-1.0 * ANGSTROM, external="free", thermostat="langevin", propagator="cartesian",
                         seed=12345, dt=FEMTOSECOND, obs_classical="kelvin", obs_bosonic="true")
    if name == "c5":   # exchange stress test: free bosons in a trap, N=8192, P=256
        return SimConfig(nbeads=256, natoms=8192, ndim=3, bosonic=True, fixcom=True, pbc=False,
                         temperature=5 * KELVIN, mass=4.0026 * DALTON, size=100 * ANGSTROM, interaction="free",
                         external="harmonic", ext_omega=3 * MEV, thermostat="langevin", propagator="cartesian",
                         seed=12345, dt=FEMTOSECOND, obs_classical="kelvin", obs_bosonic="true")
    raise ValueError(f"unknown workload {name}")


DESCRIPTION = {
    "c1": "3D harmonic trap, N=16 bosons, P=32, free interaction, cartesian+Langevin",
    "c2": "2D dipolar particles in harmonic trap, N=64, P=64, normal_modes propagator + nmthermostat",
    "c3": "Liquid He-4 Aziz, PBC, N=512 bosons, P=64 beads, 3D, cartesian+Langevin, fixcom",
    "c4": "Liquid He-4 Aziz, PBC, N=2048 bosons, P=128 beads, 3D",
    "c5": "Exchange stress: free bosons in harmonic trap, N=8192, P=256",
}


def lattice(cfg: SimConfig) -> np.ndarray:
    """Cubic lattice sites, cell centres, [-L/2, L/2)^D (the reference's `grid` placement, src/simulation.cpp:133-193)."""
    n, d, L = cfg.natoms, cfg.ndim, cfg.size
    m = int(np.ceil(n ** (1.0 / d) - 1e-9))
    grid = np.stack(np.meshgrid(*[np.arange(m)] * d, indexing="ij"), axis=-1).reshape(-1, d)[:n]
    return (grid + 0.5) * (L / m) - 0.5 * L


def initial_state(cfg: SimConfig, name: str, seed: int = None) -> Tuple[np.ndarray, np.ndarray]:
    """(x, p) as [P][N][D] float64 in atomic units."""
    rng = np.random.default_rng(cfg.seed if seed is None else seed)
    P, N, D = cfg.nbeads, cfg.natoms, cfg.ndim
    if cfg.interaction == "aziz":
        # never random-uniform with Aziz (hard-core overlaps): lattice + Gaussian bead spread
        x = np.repeat(lattice(cfg)[None], P, axis=0) + rng.normal(0.0, 0.15 * ANGSTROM, size=(P, N, D))
    elif cfg.interaction == "dipole":
        # 1/r^3 repulsion: start from a lattice (spacing L/m) with a small bead spread, not from overlapping particles
        x = np.repeat(lattice(cfg)[None], P, axis=0) * 0.9 + rng.normal(0.0, 0.005 * cfg.size, size=(P, N, D))
    elif name.lower() == "c5":
        # thermal cloud of the trap, ring polymers collapsed on their centroid + small spread
        sigma = np.sqrt(cfg.temperature / (cfg.mass * cfg.ext_omega ** 2))
        x = np.repeat(rng.normal(0.0, sigma, size=(1, N, D)), P, axis=0) + rng.normal(0.0, 0.05 * sigma, size=(P, N, D))
    else:
        x = np.repeat(rng.uniform(-0.25 * cfg.size, 0.25 * cfg.size, size=(1, N, D)), P, axis=0)
        x = x + rng.normal(0.0, 0.01 * cfg.size, size=(P, N, D))
    p = rng.normal(0.0, np.sqrt(cfg.mass / cfg.thermo_beta), size=(P, N, D))
    return np.ascontiguousarray(x), np.ascontiguousarray(p)


def pair_flops_per_step(cfg: SimConfig) -> float:
    """Algorithmic FP64 flops of the pair-force kernel per MD step (68 per unique Aziz pair in 3-D
    with PBC, 17 per dipole pair in 2-D, 12 per harmonic pair in 3-D)."""
    per_pair = {"aziz": 68.0, "dipole": 17.0, "harmonic": 12.0}.get(cfg.interaction, 0.0)
    return per_pair * cfg.nbeads * cfg.natoms * (cfg.natoms - 1) / 2.0


def integrator_bytes_per_step(cfg: SimConfig) -> float:
    """80 B per degree of freedom per step."""
    return 80.0 * cfg.nbeads * cfg.natoms * cfg.ndim
