"""GPU parity tests proper: the CUDA path, called through the C ABI (libpimdb200.so), against the CPU oracle
(oracle/pimd_oracle.c, itself pinned bit-for-bit to the unmodified reference; see test_oracle_golden.py) on
the same seeded inputs.

Tolerances (north_star: FP64 forces / energies within 1e-10 relative on identical positions). The kernels sum in
another order than the reference's i<j loop, so errors are measured as max|a-b| / max|b| over an array.
"""
import numpy as np
import pytest

from pimd_b_b200.config import SimConfig
from pimd_b_b200.engine import DeviceSim
from tests.helpers import (ANGSTROM, DALTON, FEMTOSECOND, KELVIN, MEV, Oracle, lattice_positions,
                           maxwell_momenta, relerr)

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-10      # north_star
ENERGY_TOL = 1e-10


def helium(N, P, **kw):
    L = (N / 0.02186) ** (1.0 / 3.0) * ANGSTROM
    base = dict(nbeads=P, natoms=N, ndim=3, bosonic=True, fixcom=True, pbc=True, temperature=2 * KELVIN,
                mass=4.0026 * DALTON, size=L, interaction="aziz", cutoff=-1 * ANGSTROM, external="free",
                thermostat="langevin", seed=12345, dt=FEMTOSECOND)
    base.update(kw)
    return SimConfig(**base)


def trap(N, P, D=3, **kw):
    base = dict(nbeads=P, natoms=N, ndim=D, bosonic=True, fixcom=False, pbc=False, temperature=5.802 * KELVIN,
                mass=1.0, size=300.0, interaction="free", external="harmonic", ext_omega=3 * MEV,
                thermostat="langevin", seed=90846, dt=FEMTOSECOND)
    base.update(kw)
    return SimConfig(**base)


def make_inputs(cfg, seed, kind):
    rng = np.random.default_rng(seed)
    if kind == "lattice":
        x = lattice_positions(cfg, rng, 0.15 * ANGSTROM)
    elif kind == "pure_lattice":
        x = lattice_positions(cfg, rng, 0.0)
    else:
        x = rng.uniform(-0.5 * cfg.size, 0.5 * cfg.size, size=(cfg.nbeads, cfg.natoms, cfg.ndim)) * kind
    p = maxwell_momenta(cfg, rng)
    return x, p


FORCE_CASES = {
    "aziz_pbc_bosonic": (helium(64, 8), "lattice"),
    "aziz_pbc_bosonic_cutoff": (helium(64, 8, cutoff=5 * ANGSTROM), "lattice"),
    "aziz_pbc_exact_lattice_half_box": (helium(64, 4), "pure_lattice"),
    "aziz_pbc_dist_ragged_N50": (helium(50, 4, bosonic=False), "lattice"),
    "aziz_open_N33": (helium(33, 3, pbc=False, bosonic=False), "lattice"),
    "aziz_N200_P4": (helium(200, 4), "lattice"),
    "dipole_2d_trap": (trap(64, 8, D=2, bosonic=False, interaction="dipole", int_strength=1.0, size=200.0,
                            temperature=5 * KELVIN), 0.5),
    "dipole_2d_trap_bosonic": (trap(40, 6, D=2, interaction="dipole", int_strength=1.0, size=200.0), 0.5),
    "harmonic_pair_bosonic": (trap(10, 4, interaction="harmonic", int_omega=1 * MEV), 0.1),
    "harmonic_pair_cutoff_1d": (trap(37, 5, D=1, bosonic=False, interaction="harmonic", int_omega=1 * MEV,
                                     cutoff=40.0), 0.3),
    "free_trap_bosonic_golden_like": (trap(8, 8), 1.0),
    # stiff springs + uncorrelated bead positions: beta*E ~ 1e3-1e4 per link, W drops by e^-thousands per step, so
    # every 32-step owner phase of the recurrence leaves its 2^+-400 window and takes the exact fallback path
    "stiff_random_bosonic_fallback_N80": (trap(80, 4, mass=4.0026 * DALTON, temperature=2 * KELVIN, size=40.0), 1.0),
    "stiff_random_bosonic_fallback_N200_pbc": (trap(200, 3, mass=4.0026 * DALTON, temperature=2 * KELVIN, size=60.0,
                                                    pbc=True, external="free"), 1.0),
    "free_trap_bosonic_N16_P32": (trap(16, 32), 1.0),
    "free_trap_bosonic_P2": (trap(12, 2), 0.2),
    "free_trap_bosonic_N1": (trap(1, 4), 0.2),
    "free_free_bosonic_pbc": (trap(20, 4, external="free", pbc=True, size=40.0), 2.0),
    "dist_free_trap_P1": (trap(5, 1, bosonic=False), 0.2),
    "double_well_ext": (trap(9, 4, bosonic=False, external="double_well", ext_strength=1e-6, ext_location=3.0), 0.05),
    "cosine_ext": (trap(9, 4, bosonic=False, external="cosine", ext_amplitude=1e-3, ext_phase=0.3), 0.5),
}


@pytest.mark.parametrize("name", list(FORCE_CASES))
def test_forces_match_oracle(gpu_required, name):
    cfg, kind = FORCE_CASES[name]
    x, p = make_inputs(cfg, 7, kind)
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.set("p", p)
    orc.update_forces()
    sim = DeviceSim(cfg)
    sim.set("x", x)
    sim.set("p", p)
    sim.update_forces()
    f, fs, fp = sim.get("f"), sim.get("f_spring"), sim.get("f_phys")
    assert relerr(fs, orc.get("s")) < FORCE_TOL
    assert relerr(fp, orc.get("e")) < FORCE_TOL
    assert relerr(f, orc.get("f")) < FORCE_TOL
    # state round trip through the AoS<->SoA boundary is exact
    assert np.array_equal(sim.get("x"), x)
    assert np.array_equal(sim.get("p"), p)
    sim.close()


@pytest.mark.parametrize("name", ["aziz_pbc_bosonic", "free_trap_bosonic_golden_like", "free_trap_bosonic_N16_P32",
                                  "free_free_bosonic_pbc", "dipole_2d_trap_bosonic", "free_trap_bosonic_P2",
                                  "stiff_random_bosonic_fallback_N80", "stiff_random_bosonic_fallback_N200_pbc"])
def test_exchange_tables_match_oracle(gpu_required, name):
    cfg, kind = FORCE_CASES[name]
    x, p = make_inputs(cfg, 11, kind)
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.update_forces()
    sim = DeviceSim(cfg)
    sim.set("x", x)
    sim.exchange_prepare()
    V, Vb = sim.exchange("V"), sim.exchange("Vb")
    assert relerr(V, orc.exchange("V")) < ENERGY_TOL
    assert relerr(Vb, orc.exchange("B")) < ENERGY_TOL
    assert Vb[0] == V[-1]                                  # V_backwards[0] = V[N]
    assert relerr(sim.exchange("E"), orc.exchange("E")) < ENERGY_TOL
    prob = sim.exchange("prob").reshape(cfg.natoms, cfg.natoms)
    ref = orc.exchange("P").reshape(cfg.natoms, cfg.natoms)
    assert np.max(np.abs(prob - ref)) < 1e-10
    # every particle's last bead connects somewhere: rows sum to one (SURVEY.md 8c invariant)
    assert np.allclose(prob.sum(axis=1), 1.0, atol=1e-10)
    sim.close()


@pytest.mark.parametrize("name", list(FORCE_CASES))
def test_observables_match_oracle(gpu_required, name):
    cfg, kind = FORCE_CASES[name]
    x, p = make_inputs(cfg, 13, kind)
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.set("p", p)
    ref = orc.observables()
    sim = DeviceSim(cfg)
    sim.set("x", x)
    sim.set("p", p)
    got = sim.observables()
    # energy columns are differences of large terms (kinetic = const - spring/P, virial = signed sum): measure
    # their error against the largest energy in the row; momentum columns against themselves
    scale = max(abs(ref["kinetic"]), abs(ref["potential"]), abs(ref["cl_spring"]) / cfg.nbeads, abs(ref["virial"]))
    for key in ("kinetic", "potential", "ext_pot", "int_pot", "virial", "cl_spring"):
        tol = ENERGY_TOL * max(scale, abs(ref[key]))
        assert abs(got[key] - ref[key]) <= tol, (key, got[key], ref[key])
    for key in ("cl_kinetic", "temperature"):
        assert abs(got[key] - ref[key]) <= ENERGY_TOL * abs(ref[key]), (key, got[key], ref[key])
    for key in ("prob_dist", "prob_all"):
        assert abs(got[key] - ref[key]) <= 1e-9 * max(abs(ref[key]), 1e-300) + 1e-300, (key, got[key], ref[key])
    for key in ("w_gsf", "pot_gsf"):   # GSF action observable: defined for a free interaction only
        if cfg.interaction == "free":
            assert abs(got[key] - ref[key]) <= ENERGY_TOL * max(abs(ref[key]), 1e-300), (key, got[key], ref[key])
        else:
            assert np.isnan(got[key]) and np.isnan(ref[key])
    sim.close()


TRAJ_CASES = {
    "aziz_nve_bosonic": (helium(64, 8, thermostat="none"), "lattice"),
    "aziz_nve_dist_nofixcom": (helium(50, 4, thermostat="none", bosonic=False, fixcom=False), "lattice"),
    "trap_nve_bosonic": (trap(8, 8, thermostat="none", fixcom=True), 0.2),
    "dipole_nm_propagator_nve": (trap(32, 8, D=2, bosonic=False, interaction="dipole", size=200.0,
                                      thermostat="none", propagator="normal_modes"), 0.5),
    "trap_nm_propagator_oddP": (trap(6, 7, bosonic=False, thermostat="none", propagator="normal_modes",
                                     fixcom=True), 0.2),
    # Nose-Hoover chains are deterministic: trajectory-level parity
    "trap_nose_hoover_bosonic": (trap(8, 4, thermostat="nose_hoover", dt=0.1 * FEMTOSECOND), 0.2),
    "trap_nose_hoover_np_bosonic": (trap(8, 4, thermostat="nose_hoover_np", dt=0.1 * FEMTOSECOND, fixcom=True), 0.2),
    "trap_nose_hoover_np_dim": (trap(9, 5, thermostat="nose_hoover_np_dim", dt=0.1 * FEMTOSECOND, nchains=3,
                                     bosonic=False), 0.2),
    "aziz_nose_hoover_2chains": (helium(33, 4, thermostat="nose_hoover", nchains=2), "lattice"),
}


@pytest.mark.parametrize("name", list(TRAJ_CASES))
def test_nve_trajectory_matches_oracle(gpu_required, name):
    """Deterministic (noise-free) trajectories: 40 iterations of the run-loop body, fused + graph-replayed on the
    GPU vs the oracle. Tolerance 1e-9 relative on x, p, f (rounding differences amplify slowly over 40 steps)."""
    cfg, kind = TRAJ_CASES[name]
    x, p = make_inputs(cfg, 17, kind)
    K = 40
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.set("p", p)
    for _ in range(K):
        orc.run_iteration()
    sim = DeviceSim(cfg)
    sim.set("x", x)
    sim.set("p", p)
    sim.step(K)
    sim.synchronize()
    assert relerr(sim.get("x"), orc.get("x")) < 1e-9
    assert relerr(sim.get("p"), orc.get("p")) < 1e-9
    assert relerr(sim.get("f"), orc.get("f")) < 1e-9
    sim.close()


def test_piecewise_calls_equal_fused_step(gpu_required):
    """Driving the reference's loop order call by call gives the same state as the fused, graph-captured step."""
    cfg, kind = TRAJ_CASES["aziz_nve_bosonic"]
    x, p = make_inputs(cfg, 19, kind)
    a, b = DeviceSim(cfg), DeviceSim(cfg)
    for s in (a, b):
        s.set("x", x)
        s.set("p", p)
    for _ in range(5):
        a.thermostat_step()
        a.zero_momentum()
        a.moment_step()
        a.coords_step()
        a.update_neighboring_coordinates()
        a.update_forces()
        a.moment_step()
        a.thermostat_step()
        a.zero_momentum()
    b.step(5)
    for w in ("x", "p", "f"):
        assert np.array_equal(a.get(w), b.get(w)), w
    a.close()
    b.close()


def test_zero_momentum(gpu_required):
    cfg, kind = FORCE_CASES["aziz_pbc_dist_ragged_N50"]
    x, p = make_inputs(cfg, 23, kind)
    p += 0.37
    orc = Oracle(cfg)
    orc.set("p", p)
    orc.zero_momentum()
    sim = DeviceSim(cfg)
    sim.set("x", x)
    sim.set("p", p)
    sim.zero_momentum()
    got = sim.get("p")
    assert relerr(got, orc.get("p")) < 1e-13
    assert np.max(np.abs(got.sum(axis=(0, 1)))) < 1e-9 * np.abs(p).sum()
    sim.close()


@pytest.mark.parametrize("scalar", [False, True], ids=["blocked", "scalar"])
@pytest.mark.parametrize("natoms", [129, 300, 512, 700, 1000, 1100, 2100, 4200])
def test_exchange_large_n_paths(gpu_required, natoms, scalar, monkeypatch):
    """Every recurrence kernel against the oracle on V, V_backwards and the exterior forces. Default path: the blocked
    recurrence on thread-block clusters (one row block per warp up to N = 2048, several beyond; N not a multiple of 32
    or 4). Cross-check path (PIMDB_EXCH_NOBLOCKED=1, also what N > 8192 runs): the scalar extended-range kernel, one
    row per thread up to N = 1024, then 2 rows per thread with cp.async staging, 4+ rows with direct loads."""
    if scalar:
        monkeypatch.setenv("PIMDB_EXCH_NOBLOCKED", "1")
    cfg = trap(natoms, 3, temperature=1.0 * KELVIN, size=2000.0)
    rng = np.random.default_rng(natoms)
    centroid = rng.normal(0.0, 60.0, size=(1, natoms, 3))
    x = np.repeat(centroid, 3, axis=0) + rng.normal(0.0, 6.0, size=(3, natoms, 3))
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.update_forces()
    sim = DeviceSim(cfg)
    sim.set("x", x)
    sim.update_forces()
    assert relerr(sim.exchange("V"), orc.exchange("V")) < ENERGY_TOL
    assert relerr(sim.exchange("Vb"), orc.exchange("B")) < ENERGY_TOL
    assert relerr(sim.get("f_spring"), orc.get("s")) < FORCE_TOL
    if natoms <= 2100:
        ref = orc.observables()
        got = sim.observables()
        for key in ("kinetic", "cl_spring", "prob_dist", "prob_all"):
            assert abs(got[key] - ref[key]) <= 1e-9 * max(abs(ref[key]), abs(ref["cl_spring"]) / cfg.nbeads) + 1e-300, key
    sim.close()
