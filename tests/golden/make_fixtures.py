#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (run in the build container, where /root/reference and
oracle/_ref exist; the fixtures are committed, this script documents how they were made).

    python tests/golden/make_fixtures.py

1. refcases.npz   -- frames of the reference's OWN golden regression cases (/root/reference/tests/cases/<case>):
                     positions (angstrom), velocities (angstrom/ps), forces (eV/angstrom) of frames 1, 50 and the
                     last one for every bead, the first rows of simulation.out, and the case's INI text. A frame's
                     positions and forces are a stand-alone force known-answer test.
2. refprobe.npz   -- raw-double outputs of the UNMODIFIED reference (oracle/_ref/ref_probe_ndim*, built by
                     oracle/Makefile) for the paths no reference test covers: Aziz / dipole / harmonic-pair forces,
                     minimum image, cutoff, NDIM = 2, fixcom, exchange tables, observables, and short Langevin /
                     NVE / normal-mode trajectories (these pin the RANMAR stream and the loop order bit for bit).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from pimd_b_b200 import io as pio  # noqa: E402
from pimd_b_b200.config import SimConfig  # noqa: E402
from tests.helpers import (ANGSTROM, DALTON, FEMTOSECOND, KELVIN, MEV, lattice_positions, maxwell_momenta,  # noqa: E402
                           run_ref_probe)

REF_CASES = Path("/root/reference/tests/cases")
OUT = Path(__file__).resolve().parent

CASES = ["bosonic_quadratic_harmonic_dynamics", "dist_harmonic_dynamics",
         "bosonic_quadratic_harmonic_nmthermostat_dynamics", "dist_harmonic_nm_propagation_dynamics",
         "bosonic_quadratic_harmonic", "dist_harmonic", "bosonic_quadratic_harmonic_gsf",
         # (the two bosonic_factorial_* cases belong to the reference's factorial build: the sum over all N! permutations is
         # a pointwise different potential from the Feldman-Hirshberg one -- same partition function, other forces)
         # deterministic after initialisation (Nose-Hoover chains): whole simulation.out is a known-answer test
         "bosonic_quadratic_harmonic_nh_dynamics", "bosonic_quadratic_harmonic_nh_np_dynamics",
         "bosonic_quadratic_harmonic_nh_np_dim_dynamics"]


def make_refcases():
    out = {}
    for case in CASES:
        d = REF_CASES / case
        ini = next(d.glob("*.ini")).read_text()
        out[f"{case}/ini"] = np.array(ini)
        so = pio.read_simulation_out(str(d / "simulation.out"))
        out[f"{case}/simout_columns"] = np.array(list(so.keys()))
        out[f"{case}/simout_head"] = np.stack([v[:6] for v in so.values()], axis=1)
        # the whole simulation.out is a known-answer test: deterministic after initialisation for the Nose-Hoover cases,
        # and for the Langevin cases once the GPU draws from the reference's own generator (rng = ranmars)
        out[f"{case}/simout"] = np.stack(list(so.values()), axis=1)
        if not (d / "position_0.xyz").exists():
            continue
        nb = len(list(d.glob("position_*.xyz")))
        xs, vs, fs = [], [], []
        for b in range(nb):
            xf = pio.read_dump_frames(str(d / f"position_{b}.xyz"), 3)
            vf = pio.read_dump_frames(str(d / f"velocity_{b}.dat"), 3)
            ff = pio.read_dump_frames(str(d / f"force_{b}.dat"), 3)
            sel = [1, min(50, len(xf) - 2), len(xf) - 1]
            xs.append([xf[i] for i in sel])
            vs.append([vf[i] for i in sel])
            fs.append([ff[i] for i in sel])
        # -> [frame][bead][atom][3]
        out[f"{case}/x"] = np.transpose(np.asarray(xs), (1, 0, 2, 3))
        out[f"{case}/v"] = np.transpose(np.asarray(vs), (1, 0, 2, 3))
        out[f"{case}/f"] = np.transpose(np.asarray(fs), (1, 0, 2, 3))
    np.savez_compressed(OUT / "refcases.npz", **out)
    print("wrote refcases.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


def probe_configs():
    """name -> (SimConfig, x, p) for the reference paths that have no golden case."""
    rng = np.random.default_rng(20261017)
    items = {}
    N, P = 27, 6
    L = (N / 0.02186) ** (1 / 3) * ANGSTROM
    he = dict(nbeads=P, natoms=N, ndim=3, bosonic=True, fixcom=True, pbc=True, temperature=2 * KELVIN,
              mass=4.0026 * DALTON, size=L, interaction="aziz", cutoff=-1 * ANGSTROM, external="free",
              thermostat="langevin", seed=12345, dt=FEMTOSECOND, obs_classical="kelvin", obs_bosonic="true")
    c = SimConfig(**he)
    x, p = lattice_positions(c, rng, 0.15 * ANGSTROM), maxwell_momenta(c, rng)
    items["aziz_pbc_bosonic"] = (c, x, p)
    items["aziz_pbc_bosonic_cutoff"] = (SimConfig(**{**he, "cutoff": 5 * ANGSTROM}), x, p)
    items["aziz_pbc_dist_nve"] = (SimConfig(**{**he, "bosonic": False, "thermostat": "none"}), x, p)
    c = SimConfig(**{**he, "natoms": 8, "nbeads": 4, "size": (8 / 0.02186) ** (1 / 3) * ANGSTROM})
    items["aziz_pbc_exact_lattice"] = (c, lattice_positions(c, rng, 0.0), maxwell_momenta(c, rng))
    c = SimConfig(nbeads=8, natoms=12, ndim=2, bosonic=False, fixcom=False, pbc=False, temperature=5 * KELVIN,
                  mass=1.0, size=200.0, interaction="dipole", int_strength=1.0, external="harmonic",
                  ext_omega=3 * MEV, thermostat="langevin", propagator="normal_modes", nmthermostat=True,
                  seed=777, dt=FEMTOSECOND, obs_classical="kelvin")
    items["dipole_2d_nm"] = (c, rng.uniform(-100, 100, size=(8, 12, 2)), maxwell_momenta(c, rng))
    c = SimConfig(nbeads=4, natoms=10, ndim=3, bosonic=True, fixcom=True, pbc=False, temperature=5.802 * KELVIN,
                  mass=1.0, size=300.0, interaction="harmonic", int_omega=1 * MEV, external="harmonic",
                  ext_omega=3 * MEV, thermostat="langevin", nmthermostat=True, seed=5, dt=FEMTOSECOND,
                  obs_classical="kelvin", obs_bosonic="true")
    items["harmonic_pair_bosonic_nmthermo"] = (c, rng.uniform(-30, 30, size=(4, 10, 3)), maxwell_momenta(c, rng))
    c = SimConfig(nbeads=5, natoms=7, ndim=1, bosonic=True, fixcom=False, pbc=True, temperature=3 * KELVIN,
                  mass=2.0, size=50.0, interaction="harmonic", int_omega=2 * MEV, cutoff=20.0, external="free",
                  thermostat="none", seed=9, dt=FEMTOSECOND, obs_classical="kelvin", obs_bosonic="true")
    items["harmonic_pair_1d_pbc_cutoff"] = (c, rng.uniform(-25, 25, size=(5, 7, 1)), maxwell_momenta(c, rng))
    # Nose-Hoover chains coupled to the normal modes (NMCoupling, src/thermostats/thermostat_coupling.cpp:29-47): no
    # golden case of the reference covers them (appended last: the inputs of the cases above do not change)
    for th in ("nose_hoover", "nose_hoover_np", "nose_hoover_np_dim"):
        c = SimConfig(nbeads=6, natoms=9, ndim=3, bosonic=(th == "nose_hoover_np"), fixcom=True, pbc=False,
                      temperature=5.802 * KELVIN, mass=1.0, size=300.0, interaction="harmonic", int_omega=1 * MEV,
                      external="harmonic", ext_omega=3 * MEV, thermostat=th, nmthermostat=True, nchains=4, seed=11,
                      dt=FEMTOSECOND, obs_classical="kelvin", obs_bosonic="true" if th == "nose_hoover_np" else "false")
        items[f"{th}_nmcoupled"] = (c, rng.uniform(-30, 30, size=(6, 9, 3)), maxwell_momenta(c, rng))
    # external double-well and cosine potentials (src/potentials/double_well.cpp, cosine.cpp): no golden case of the
    # reference uses them (appended last again)
    base = dict(nbeads=4, natoms=9, ndim=3, bosonic=False, fixcom=False, pbc=False, temperature=5.802 * KELVIN,
                mass=1.0, size=300.0, interaction="free", thermostat="none", seed=3, dt=FEMTOSECOND,
                obs_classical="kelvin")
    c = SimConfig(**base, external="double_well", ext_strength=1e-6, ext_location=3.0)
    items["double_well_ext"] = (c, rng.uniform(-8, 8, size=(4, 9, 3)), maxwell_momenta(c, rng))
    c = SimConfig(**{**base, "bosonic": True, "obs_bosonic": "true", "interaction": "harmonic", "int_omega": 1 * MEV},
                  external="cosine", ext_amplitude=1e-3, ext_phase=0.3)
    items["cosine_ext_bosonic_pair"] = (c, rng.uniform(-100, 100, size=(4, 9, 3)), maxwell_momenta(c, rng))
    # GSF action observable (src/observables/gsf_action.cpp), free interaction: odd bead count so that odd and even
    # time slices differ in number; harmonic trap (bosonic) and double well
    gs = dict(nbeads=5, natoms=6, ndim=3, fixcom=False, pbc=False, temperature=5.802 * KELVIN, mass=1.0, size=300.0,
              interaction="free", thermostat="none", seed=4, dt=FEMTOSECOND, obs_classical="kelvin", obs_gsf="atomic_unit")
    c = SimConfig(**gs, bosonic=True, obs_bosonic="true", external="harmonic", ext_omega=3 * MEV)
    items["gsf_trap_bosonic"] = (c, rng.uniform(-30, 30, size=(5, 6, 3)), maxwell_momenta(c, rng))
    c = SimConfig(**gs, bosonic=False, external="double_well", ext_strength=1e-6, ext_location=3.0)
    items["gsf_double_well"] = (c, rng.uniform(-8, 8, size=(5, 6, 3)), maxwell_momenta(c, rng))
    return items


def make_refprobe():
    out = {}
    for name, (cfg, x, p) in probe_configs().items():
        out[f"{name}/cfg"] = np.array(repr(cfg.as_dict()))
        out[f"{name}/x"] = x
        out[f"{name}/p"] = p
        r = run_ref_probe(cfg, x, p, "forces")
        for k in ("f", "f_spring", "f_phys"):
            out[f"{name}/{k}"] = r[k]
        out[f"{name}/obs_names"] = np.array(list(r["obs"].keys()))
        out[f"{name}/obs_values"] = np.array(list(r["obs"].values()))
        if cfg.bosonic:
            for k in ("exch_V", "exch_Vb", "exch_E", "exch_prob"):
                out[f"{name}/{k}"] = r[k]
            out[f"{name}/exch_scalar_names"] = np.array(list(r["exch_scalars"].keys()))
            out[f"{name}/exch_scalar_values"] = np.array(list(r["exch_scalars"].values()))
        K = 12
        t = run_ref_probe(cfg, x, p, "traj", k=K, every=K)
        for k in ("x", "p", "f"):
            out[f"{name}/traj{K}_{k}"] = t[f"{k}_{K}"]
    np.savez_compressed(OUT / "refprobe.npz", **out)
    print("wrote refprobe.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


# ---------------------------------------------------------------------------------------------------------------------
# The reference's compile-time alternative exchange class (-DFACTORIAL_BOSONIC_ALGORITHM, the sum over all N! permutations):
# its two golden regression cases, and raw outputs of the unmodified reference BUILT WITH THAT FLAG
# (oracle/_ref/ref_probe_ndim3_factorial, oracle/Makefile) on seeded inputs.
FACTORIAL_CASES = ["bosonic_factorial_harmonic", "bosonic_factorial_harmonic_dynamics"]


def factorial_probe_configs():
    rng = np.random.default_rng(20261018)
    items = {}
    base = dict(ndim=3, bosonic=True, fixcom=False, temperature=5.802 * KELVIN, mass=1.0, interaction="free",
                external="harmonic", ext_omega=3 * MEV, thermostat="langevin", dt=FEMTOSECOND, exchange_alg="factorial",
                obs_classical="kelvin", obs_bosonic="true")
    c = SimConfig(**base, nbeads=4, natoms=3, pbc=False, size=300.0, seed=7)
    items["factorial_trap_n3"] = (c, rng.uniform(-20, 20, size=(4, 3, 3)), maxwell_momenta(c, rng))
    c = SimConfig(**base, nbeads=3, natoms=5, pbc=True, size=60.0, seed=8)
    items["factorial_trap_n5_pbc"] = (c, rng.uniform(-40, 40, size=(3, 5, 3)), maxwell_momenta(c, rng))
    c = SimConfig(**{**base, "fixcom": True, "thermostat": "none"}, nbeads=8, natoms=7, pbc=False, size=300.0, seed=9)
    items["factorial_trap_n7_nve_fixcom"] = (c, rng.uniform(-20, 20, size=(8, 7, 3)), maxwell_momenta(c, rng))
    c = SimConfig(**{**base, "interaction": "harmonic", "int_omega": 1 * MEV}, nbeads=2, natoms=4, pbc=False, size=300.0, seed=10)
    items["factorial_pair_n4_two_beads"] = (c, rng.uniform(-20, 20, size=(2, 4, 3)), maxwell_momenta(c, rng))
    return items


def make_reffactorial():
    out = {}
    for case in FACTORIAL_CASES:
        d = REF_CASES / case
        out[f"{case}/ini"] = np.array(next(d.glob("*.ini")).read_text())
        so = pio.read_simulation_out(str(d / "simulation.out"))
        out[f"{case}/simout_columns"] = np.array(list(so.keys()))
        out[f"{case}/simout"] = np.stack(list(so.values()), axis=1)
        if (d / "position_0.xyz").exists():
            nb = len(list(d.glob("position_*.xyz")))
            xs, fs = [], []
            for b in range(nb):
                xf = pio.read_dump_frames(str(d / f"position_{b}.xyz"), 3)
                ff = pio.read_dump_frames(str(d / f"force_{b}.dat"), 3)
                sel = [1, min(50, len(xf) - 2), len(xf) - 1]
                xs.append([xf[i] for i in sel])
                fs.append([ff[i] for i in sel])
            out[f"{case}/x"] = np.transpose(np.asarray(xs), (1, 0, 2, 3))
            out[f"{case}/f"] = np.transpose(np.asarray(fs), (1, 0, 2, 3))
    for name, (cfg, x, p) in factorial_probe_configs().items():
        out[f"{name}/cfg"] = np.array(repr(cfg.as_dict()))
        out[f"{name}/x"] = x
        out[f"{name}/p"] = p
        r = run_ref_probe(cfg, x, p, "forces")
        for k in ("f", "f_spring", "f_phys"):
            out[f"{name}/{k}"] = r[k]
        out[f"{name}/obs_names"] = np.array(list(r["obs"].keys()))
        out[f"{name}/obs_values"] = np.array(list(r["obs"].values()))
        t = run_ref_probe(cfg, x, p, "traj", k=12, every=12)
        for k in ("x", "p", "f"):
            out[f"{name}/traj12_{k}"] = t[f"{k}_12"]
    np.savez_compressed(OUT / "reffactorial.npz", **out)
    print("wrote reffactorial.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


E2E_INIS = {
    # the reference's own initial conditions (mt19937(seed+bead): uniform positions, Maxwell-Boltzmann momenta) and
    # no noise -> the whole run is deterministic and comparable file by file
    "nve_trap_bosonic": """[simulation]
dt = 1.0 femtosecond
steps = 200
sfreq = 50
threshold = 0.0
nbeads = 4
fixcom = true
bosonic = true
seed = 90846
pbc = false
initial_position = random
initial_velocity = random
thermostat = none
[system]
temperature = 5.802 kelvin
natoms = 8
size = 300.0 atomic_unit
mass = 1.0 atomic_unit
[interaction_potential]
name = free
[external_potential]
name = harmonic
omega = 3.0 millielectronvolt
[output]
positions = angstrom
velocities = angstrom/ps
forces = ev/ang
[observables]
energy = kelvin
classical = kelvin
bosonic = true
""",
    "nve_aziz_grid_pbc": """[simulation]
dt = 0.5 femtosecond
steps = 120
sfreq = 40
threshold = 0.0
nbeads = 4
fixcom = false
bosonic = true
seed = 4242
pbc = true
initial_position = grid
initial_velocity = random
thermostat = none
[system]
temperature = 2.0 kelvin
natoms = 27
size = 10.73 angstrom
mass = 4.0026 dalton
[interaction_potential]
name = aziz
cutoff = 5.0 angstrom
[external_potential]
name = free
[output]
positions = angstrom
velocities = off
forces = ev/ang
[observables]
energy = kelvin
classical = kelvin
bosonic = true
""",
    "nve_dipole_nm_propagator": """[simulation]
dt = 1.0 femtosecond
steps = 150
sfreq = 50
threshold = 0.0
nbeads = 6
fixcom = true
bosonic = false
seed = 777
pbc = false
initial_position = random
initial_velocity = random
propagator = normal_modes
thermostat = none
[system]
temperature = 5.0 kelvin
natoms = 10
size = 200.0 atomic_unit
mass = 1.0 atomic_unit
[interaction_potential]
name = dipole
strength = 1.0
[external_potential]
name = harmonic
omega = 3.0 millielectronvolt
[output]
positions = angstrom
velocities = angstrom/ps
forces = off
[observables]
energy = kelvin
classical = kelvin
""",
}


def make_e2e():
    """Run the UNMODIFIED reference program (oracle/_ref/pimdb_ndim3) end to end and keep its output files."""
    import os, shutil, subprocess, tempfile
    out = {}
    for name, ini in E2E_INIS.items():
        ndim = 2 if "dipole" in name else 3
        tmp = Path(tempfile.mkdtemp(prefix="e2e_"))
        (tmp / "config.ini").write_text(ini)
        nb = int([l for l in ini.splitlines() if l.startswith("nbeads")][0].split("=")[1])
        subprocess.run([str(ROOT / "oracle" / "_ref" / f"pimdb_ndim{ndim}"), "-in", "config.ini"], cwd=tmp, check=True,
                       env=dict(os.environ, PIMDB_NP=str(nb)), stdout=subprocess.DEVNULL)
        out[f"{name}/ini"] = np.array(ini)
        out[f"{name}/ndim"] = np.array(ndim)
        so = pio.read_simulation_out(str(tmp / "output" / "simulation.out"))
        out[f"{name}/simout_columns"] = np.array(list(so.keys()))
        out[f"{name}/simout"] = np.stack(list(so.values()), axis=1)
        out[f"{name}/simout_text"] = np.array((tmp / "output" / "simulation.out").read_text())
        for kind, pat in (("x", "position_{}.xyz"), ("v", "velocity_{}.dat"), ("f", "force_{}.dat")):
            if (tmp / "output" / pat.format(0)).exists():
                out[f"{name}/{kind}"] = np.asarray(
                    [pio.read_dump_frames(str(tmp / "output" / pat.format(b)), ndim) for b in range(nb)])
        shutil.rmtree(tmp, ignore_errors=True)
    np.savez_compressed(OUT / "refe2e.npz", **out)
    print("wrote refe2e.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    which = sys.argv[1:] or ["refcases", "refprobe", "e2e", "factorial"]
    if "refcases" in which:
        make_refcases()
    if "refprobe" in which:
        make_refprobe()
    if "e2e" in which:
        make_e2e()
    if "factorial" in which:
        make_reffactorial()
