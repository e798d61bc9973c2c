"""Small systems shared by the peer-memory sharding tests (in-process and multi-process)."""
import numpy as np

from pimd_b_b200 import workloads as wl
from pimd_b_b200.config import SimConfig


def make_case(name: str):
    """-> (cfg, x, p) with x, p as [P][N][D]."""
    if name == "he_langevin":      # bosonic He-4, Aziz, PBC, fixcom, Langevin (Philox stream: independent of the sharding)
        cfg = SimConfig(nbeads=8, natoms=64, ndim=3, bosonic=True, fixcom=True, pbc=True, temperature=2 * wl.KELVIN,
                        mass=4.0026 * wl.DALTON, size=wl.helium_box(64), interaction="aziz", cutoff=-1.0 * wl.ANGSTROM,
                        external="free", thermostat="langevin", seed=321, dt=wl.FEMTOSECOND)
        kind = "c3"
    elif name == "he_nve_odd":     # odd particle count (scalar integrator path), NVE, fixcom, uneven bead split
        cfg = SimConfig(nbeads=7, natoms=27, ndim=3, bosonic=True, fixcom=True, pbc=True, temperature=2 * wl.KELVIN,
                        mass=4.0026 * wl.DALTON, size=wl.helium_box(27), interaction="aziz", cutoff=-1.0 * wl.ANGSTROM,
                        external="free", thermostat="none", seed=5, dt=wl.FEMTOSECOND)
        kind = "c3"
    elif name == "trap_nofixcom":  # free bosons in a trap, no COM exchange at all: only the halo hand-shake couples the ranks
        cfg = SimConfig(nbeads=6, natoms=16, ndim=3, bosonic=True, fixcom=False, pbc=False, temperature=5.802 * wl.KELVIN,
                        mass=1.0, size=300.0, interaction="free", external="harmonic", ext_omega=3 * wl.MEV,
                        thermostat="langevin", seed=4242, dt=wl.FEMTOSECOND)
        kind = "c1"
    elif name == "dist_nh":        # distinguishable particles, Nose-Hoover chains + fixcom: both zeroMomentum exchanges kept
        cfg = SimConfig(nbeads=8, natoms=12, ndim=2, bosonic=False, fixcom=True, pbc=False, temperature=5 * wl.KELVIN,
                        mass=1.0, size=200.0, interaction="dipole", int_strength=1.0, external="harmonic",
                        ext_omega=3 * wl.MEV, thermostat="nose_hoover", nchains=3, seed=9, dt=wl.FEMTOSECOND)
        kind = "c2"
    elif name == "nm_langevin":    # C2 in small: 2-D dipoles, normal-mode propagator + normal-mode Langevin thermostat (Philox)
        cfg = SimConfig(nbeads=8, natoms=12, ndim=2, bosonic=False, fixcom=False, pbc=False, temperature=5 * wl.KELVIN,
                        mass=1.0, size=200.0, interaction="dipole", int_strength=1.0, external="harmonic",
                        ext_omega=3 * wl.MEV, thermostat="langevin", nmthermostat=True, propagator="normal_modes", seed=777,
                        dt=wl.FEMTOSECOND)
        kind = "c2"
    elif name == "nm_nve_fixcom":  # normal-mode propagator alone, with the centre-of-mass exchange, uneven bead split
        cfg = SimConfig(nbeads=6, natoms=9, ndim=3, bosonic=False, fixcom=True, pbc=False, temperature=5 * wl.KELVIN,
                        mass=1.0, size=200.0, interaction="harmonic", int_omega=1 * wl.MEV, external="harmonic",
                        ext_omega=3 * wl.MEV, thermostat="none", propagator="normal_modes", seed=3, dt=wl.FEMTOSECOND)
        kind = "c1"
    else:
        raise ValueError(name)
    x, p = wl.initial_state(cfg, kind, seed=cfg.seed)
    return cfg, x, p
