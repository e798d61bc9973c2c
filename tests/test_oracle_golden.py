"""Pin the CPU oracle (oracle/pimd_oracle.c) to the reference:

* against frames of the reference's OWN golden regression cases (tests/golden/refcases.npz, extracted from
  /root/reference/tests/cases by tests/golden/make_fixtures.py): positions -> forces known-answer tests;
* against raw-double outputs of the unmodified reference compiled in the build container
  (tests/golden/refprobe.npz): forces, exchange tables, observables and 12-step trajectories, BIT-identical;
* live against oracle/_ref/ref_probe when it is present (build container only).
"""
import ast
import os

import numpy as np
import pytest

from pimd_b_b200.config import SimConfig, parse_ini, convert_to_internal
from tests.helpers import GOLDEN_DIR, Oracle, ref_probe_available, relerr, run_ref_probe

REFCASES = np.load(GOLDEN_DIR / "refcases.npz")
REFPROBE = np.load(GOLDEN_DIR / "refprobe.npz")
DYN_CASES = ["bosonic_quadratic_harmonic_dynamics", "dist_harmonic_dynamics",
             "bosonic_quadratic_harmonic_nmthermostat_dynamics", "dist_harmonic_nm_propagation_dynamics"]
PROBE_CASES = sorted({k.split("/")[0] for k in REFPROBE.files})


def _cfg_from_ini_text(text, tmp_path):
    p = tmp_path / "case.ini"
    p.write_text(text)
    return parse_ini(str(p), ndim=3)


@pytest.mark.parametrize("case", DYN_CASES)
def test_oracle_reproduces_reference_golden_force_frames(case, tmp_path):
    cfg = _cfg_from_ini_text(str(REFCASES[f"{case}/ini"]), tmp_path)
    ang = convert_to_internal("length", "angstrom", 1.0)
    evang = convert_to_internal("force", "ev/ang", 1.0)
    X, F = REFCASES[f"{case}/x"], REFCASES[f"{case}/f"]
    assert X.shape[1:] == (cfg.nbeads, cfg.natoms, 3)
    orc = Oracle(cfg)
    for frame in range(X.shape[0]):
        orc.set("x", X[frame] * ang)
        orc.update_forces()
        got = orc.get("f")
        # the dumps carry 13 significant digits (src/states/force.cpp:44): print-precision limited
        assert relerr(got, F[frame] * evang) < 5e-12, (case, frame)
    orc.close()


@pytest.mark.parametrize("case", PROBE_CASES)
def test_oracle_bit_identical_to_reference_raw_outputs(case):
    cfg = SimConfig(**ast.literal_eval(str(REFPROBE[f"{case}/cfg"])))
    x, p = REFPROBE[f"{case}/x"], REFPROBE[f"{case}/p"]
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.set("p", p)
    orc.update_forces()
    assert np.array_equal(orc.get("f"), REFPROBE[f"{case}/f"])
    assert np.array_equal(orc.get("s"), REFPROBE[f"{case}/f_spring"])
    assert np.array_equal(orc.get("e"), REFPROBE[f"{case}/f_phys"])
    if cfg.bosonic:
        assert np.array_equal(orc.exchange("V"), REFPROBE[f"{case}/exch_V"])
        assert np.array_equal(orc.exchange("B"), REFPROBE[f"{case}/exch_Vb"])
        assert np.array_equal(orc.exchange("E"), REFPROBE[f"{case}/exch_E"])
        assert np.array_equal(orc.exchange("P"), REFPROBE[f"{case}/exch_prob"])
    obs = orc.observables()
    kelvin = convert_to_internal("temperature", "kelvin", 1.0)
    for name, val in zip(REFPROBE[f"{case}/obs_names"], REFPROBE[f"{case}/obs_values"]):
        mine = obs[str(name)] / kelvin if name == "temperature" else obs[str(name)]   # reference prints kelvin
        assert abs(mine - val) <= 4e-16 * max(abs(val), abs(obs["cl_spring"]), 1e-300), (name, mine, val)
    # 12 iterations of the run-loop body, incl. RANMAR Langevin noise and the normal-mode paths
    orc.set("x", x)
    orc.set("p", p)
    orc.set("f", np.zeros_like(x))
    orc2 = Oracle(cfg)   # fresh RNG streams and zero forces, like the reference at start-up
    orc2.set("x", x)
    orc2.set("p", p)
    for _ in range(12):
        orc2.run_iteration()
    for w in ("x", "p", "f"):
        assert np.array_equal(orc2.get(w), REFPROBE[f"{case}/traj12_{w}"]), (case, w)
    orc.close()
    orc2.close()


def test_connection_probability_rows_sum_to_one():
    case = "aziz_pbc_bosonic"
    n = int(np.sqrt(REFPROBE[f"{case}/exch_prob"].size))
    prob = REFPROBE[f"{case}/exch_prob"].reshape(n, n)
    assert np.allclose(prob.sum(axis=1), 1.0, atol=1e-12)
    assert REFPROBE[f"{case}/exch_Vb"][0] == REFPROBE[f"{case}/exch_V"][-1]


@pytest.mark.skipif(not (ref_probe_available(3) and os.path.isdir("/root/reference")),
                    reason="unmodified reference binary only exists in the build container")
def test_oracle_live_against_reference_binary():
    from tests.helpers import ANGSTROM, DALTON, FEMTOSECOND, KELVIN, lattice_positions, maxwell_momenta
    rng = np.random.default_rng(99)
    N, P = 20, 5
    cfg = SimConfig(nbeads=P, natoms=N, ndim=3, bosonic=True, fixcom=True, pbc=True, temperature=2 * KELVIN,
                    mass=4.0026 * DALTON, size=(N / 0.02186) ** (1 / 3) * ANGSTROM, interaction="aziz",
                    cutoff=6 * ANGSTROM, external="free", thermostat="langevin", seed=31, dt=FEMTOSECOND)
    x, p = lattice_positions(cfg, rng, 0.2 * ANGSTROM), maxwell_momenta(cfg, rng)
    ref = run_ref_probe(cfg, x, p, "traj", k=8, every=8)
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.set("p", p)
    for _ in range(8):
        orc.run_iteration()
    for w in ("x", "p", "f"):
        assert np.array_equal(orc.get(w), ref[f"{w}_8"])


# ----------------------------------------------------------------------------- factorial exchange (the reference's other class)
REFFACT = np.load(GOLDEN_DIR / "reffactorial.npz")
FACT_PROBES = sorted({k.split("/")[0] for k in REFFACT.files if k.startswith("factorial_")})


@pytest.mark.parametrize("case", FACT_PROBES)
def test_oracle_factorial_exchange_bit_identical_to_reference_factorial_build(case):
    """oracle (cfg.exchange_alg = "factorial", restating src/bosonic_exchange/factorial_bosonic_exchange.cpp) against raw
    outputs of the unmodified reference built with -DFACTORIAL_BOSONIC_ALGORITHM (tests/golden/make_fixtures.py factorial)."""
    cfg = SimConfig(**ast.literal_eval(str(REFFACT[f"{case}/cfg"])))
    assert cfg.exchange_alg == "factorial"
    x, p = REFFACT[f"{case}/x"], REFFACT[f"{case}/p"]
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.set("p", p)
    orc.update_forces()
    for w, k in (("f", "f"), ("s", "f_spring"), ("e", "f_phys")):
        assert np.array_equal(orc.get(w), REFFACT[f"{case}/{k}"]), k
    obs = orc.observables()
    kelvin = convert_to_internal("temperature", "kelvin", 1.0)
    for name, val in zip(REFFACT[f"{case}/obs_names"], REFFACT[f"{case}/obs_values"]):
        mine = obs[str(name)] / kelvin if name == "temperature" else obs[str(name)]
        assert abs(mine - val) <= 4e-16 * max(abs(val), abs(obs["cl_spring"]), 1e-300), (name, mine, val)
    orc2 = Oracle(cfg)
    orc2.set("x", x)
    orc2.set("p", p)
    for _ in range(12):
        orc2.run_iteration()
    for w in ("x", "p", "f"):
        assert np.array_equal(orc2.get(w), REFFACT[f"{case}/traj12_{w}"]), (case, w)
    orc.close(); orc2.close()


def test_oracle_factorial_exchange_reproduces_reference_golden_force_frames(tmp_path):
    """The reference's own golden case of its factorial build (tests/cases/bosonic_factorial_harmonic_dynamics): every
    dumped position frame -> the dumped force frame."""
    import dataclasses
    case = "bosonic_factorial_harmonic_dynamics"
    cfg = dataclasses.replace(_cfg_from_ini_text(str(REFFACT[f"{case}/ini"]), tmp_path), exchange_alg="factorial")
    ang = convert_to_internal("length", "angstrom", 1.0)
    evang = convert_to_internal("force", "ev/ang", 1.0)
    X, F = REFFACT[f"{case}/x"], REFFACT[f"{case}/f"]
    orc = Oracle(cfg)
    for frame in range(X.shape[0]):
        orc.set("x", X[frame] * ang)
        orc.update_forces()
        assert relerr(orc.get("f"), F[frame] * evang) < 5e-12, frame
    orc.close()
