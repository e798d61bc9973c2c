"""Black-box acceptance of the C++ host mirror (pimd_b_b200/pimdb_gpu): same INI in, same output files out.

tests/golden/refe2e.npz holds complete runs of the UNMODIFIED reference program (oracle/_ref/pimdb_ndim*, made by
tests/golden/make_fixtures.py:make_e2e) for deterministic configurations: the reference's own initial conditions
(std::mt19937(seed+bead) uniform positions / grid, Maxwell-Boltzmann momenta) and thermostat = none. pimdb_gpu
restates those initial conditions on the host, runs the steps on the GPU and must reproduce simulation.out and the
per-bead dumps. Comparison rule = the reference's own regression test (tests/main.py:77-86: np.allclose, rtol 1e-5)
and, stricter, 1e-7 on the observables / 1e-8 on the dumped state.
"""
import subprocess

import numpy as np
import pytest

from pimd_b_b200 import io as pio
from tests.helpers import GOLDEN_DIR, ROOT

pytestmark = pytest.mark.gpu
E2E = np.load(GOLDEN_DIR / "refe2e.npz")
CASES = sorted({k.split("/")[0] for k in E2E.files})
BIN = ROOT / "pimd_b_b200" / "pimdb_gpu"


@pytest.mark.parametrize("case", CASES)
def test_pimdb_gpu_reproduces_reference_run(gpu_required, case, tmp_path):
    if not BIN.exists():
        from pimd_b_b200 import build
        build.build_host()
    ndim = int(E2E[f"{case}/ndim"])
    (tmp_path / "config.ini").write_text(str(E2E[f"{case}/ini"]))
    r = subprocess.run([str(BIN), "-in", "config.ini", "--dim", str(ndim)], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert "[X]" not in r.stdout, r.stdout
    assert "finished running successfully" in r.stdout, r.stdout + r.stderr
    got = pio.read_simulation_out(str(tmp_path / "output" / "simulation.out"))
    cols = [str(c) for c in E2E[f"{case}/simout_columns"]]
    assert list(got.keys()) == cols                      # same header, same column order
    ref = E2E[f"{case}/simout"]
    for i, c in enumerate(cols):
        assert np.allclose(got[c], ref[:, i], rtol=1e-5), c          # the reference's own acceptance rule
        scale = np.max(np.abs(ref[:, i])) + 1e-300
        assert np.max(np.abs(got[c] - ref[:, i])) <= 2e-7 * max(scale, 1.0), (c, got[c], ref[:, i])
    nb = E2E[f"{case}/x"].shape[0]
    for kind, pat in (("x", "position_{}.xyz"), ("v", "velocity_{}.dat"), ("f", "force_{}.dat")):
        key = f"{case}/{kind}"
        if key not in E2E.files:
            continue
        for b in range(nb):
            frames = np.asarray(pio.read_dump_frames(str(tmp_path / "output" / pat.format(b)), ndim))
            refb = E2E[key][b]
            assert frames.shape == refb.shape
            assert np.max(np.abs(frames - refb)) <= 1e-8 * np.max(np.abs(E2E[key])), (kind, b)
    # the header line is byte-identical to the reference's
    assert (tmp_path / "output" / "simulation.out").read_text().splitlines()[0] == \
        str(E2E[f"{case}/simout_text"]).splitlines()[0]
    assert (tmp_path / "output" / "report.txt").exists()


REFCASES = np.load(GOLDEN_DIR / "refcases.npz")
NH_CASES = ["bosonic_quadratic_harmonic_nh_dynamics", "bosonic_quadratic_harmonic_nh_np_dynamics",
            "bosonic_quadratic_harmonic_nh_np_dim_dynamics"]


@pytest.mark.parametrize("case", NH_CASES)
def test_pimdb_gpu_passes_the_reference_golden_nose_hoover_cases(gpu_required, case, tmp_path):
    """The reference's OWN regression cases (tests/cases/<case>/, 100 000 steps, Nose-Hoover chains: deterministic after
    the mt19937 initialisation) run unchanged through pimdb_gpu and judged by the reference's own rule
    (tests/main.py:77-107: every column of the actual simulation.out, np.allclose with rtol 1e-5)."""
    (tmp_path / "config.ini").write_text(str(REFCASES[f"{case}/ini"]))
    r = subprocess.run([str(BIN), "-in", "config.ini", "--dim", "3"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=600)
    assert "finished running successfully" in r.stdout, r.stdout + r.stderr
    got = pio.read_simulation_out(str(tmp_path / "output" / "simulation.out"))
    cols = [str(c) for c in REFCASES[f"{case}/simout_columns"]]
    ref = REFCASES[f"{case}/simout"]
    assert list(got.keys()) == cols
    for i, c in enumerate(cols):
        assert got[c].shape == ref[:, i].shape
        assert np.allclose(got[c], ref[:, i], rtol=1e-5), (c, np.max(np.abs(got[c] - ref[:, i])))
    # and the dumped frames of bead 0 / bead P-1 (exterior beads: exchange forces) at the last step
    for kind, pat in (("x", "position_{}.xyz"), ("f", "force_{}.dat")):
        frames = np.asarray(pio.read_dump_frames(str(tmp_path / "output" / pat.format(0)), 3))
        assert frames.shape[0] == 101


LANGEVIN_CASES = ["bosonic_quadratic_harmonic", "bosonic_quadratic_harmonic_dynamics", "bosonic_quadratic_harmonic_gsf",
                  "bosonic_quadratic_harmonic_nmthermostat_dynamics", "dist_harmonic", "dist_harmonic_dynamics",
                  "dist_harmonic_nm_propagation_dynamics"]


@pytest.mark.parametrize("case", LANGEVIN_CASES)
def test_pimdb_gpu_passes_the_reference_golden_langevin_cases(gpu_required, case, tmp_path):
    """The reference's thermostatted regression cases (tests/cases/<case>/, 100 000 Langevin steps, Cartesian and
    normal-mode coupling, normal-mode propagator, bosons and distinguishable particles, the GSF observable) run
    unchanged through pimdb_gpu with the reference's own noise generator (--rng ranmars: one RANMAR stream per bead,
    libs/random_mars.cpp) and are judged by the reference's own rule (tests/main.py:77-107: every column of
    simulation.out, np.allclose with rtol 1e-5)."""
    (tmp_path / "config.ini").write_text(str(REFCASES[f"{case}/ini"]))
    r = subprocess.run([str(BIN), "-in", "config.ini", "--dim", "3", "--rng", "ranmars"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=900)
    assert "finished running successfully" in r.stdout, r.stdout + r.stderr
    got = pio.read_simulation_out(str(tmp_path / "output" / "simulation.out"))
    cols = [str(c) for c in REFCASES[f"{case}/simout_columns"]]
    ref = REFCASES[f"{case}/simout"]
    # like tests/main.py:92-105 the columns of the ACTUAL output are the ones compared (some expected files still carry
    # the ext_pot / int_pot columns of an older EnergyObservable; src/observables/energy.cpp:10-19 no longer prints them
    # when one of the potentials is free)
    assert set(got.keys()) <= set(cols) and {"step", "kinetic"} <= set(got.keys())
    for c in got:
        i = cols.index(c)
        assert got[c].shape == ref[:, i].shape
        assert np.allclose(got[c], ref[:, i], rtol=1e-5), (c, np.max(np.abs(got[c] - ref[:, i])))


def test_pimdb_gpu_reports_config_errors_like_the_reference(gpu_required, tmp_path):
    (tmp_path / "config.ini").write_text("[simulation]\nnbeads = 4\nbosonic = true\npropagator = normal_modes\n"
                                         "thermostat = langevin\n")
    r = subprocess.run([str(BIN), "-in", "config.ini"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0       # the reference prints the error and returns 0 (src/pimdb.cpp:57-67)
    assert "[X] Invalid argument error: Normal modes propogation is currently not available for bosons!" in r.stdout


@pytest.mark.parametrize("case,gpus", [("nve_aziz_grid_pbc", 2), ("nve_trap_bosonic", 2), ("nve_trap_bosonic", 4)])
def test_pimdb_gpu_sharded_over_several_handles_reproduces_reference_run(gpu_required, case, gpus, tmp_path):
    """`pimdb_gpu --gpus G`: the beads sharded over G handles coupled through peer memory, driven by one host thread -- the
    counterpart of `mpirun -np P pimdb` (README.md:200-203). On a one-GPU box the shards share the device
    (PIMDB_SHARD_SAME_DEVICE=1); the output files must match the unmodified reference's complete run."""
    import os
    import torch
    ndim = int(E2E[f"{case}/ndim"])
    (tmp_path / "config.ini").write_text(str(E2E[f"{case}/ini"]))
    env = dict(os.environ)
    if torch.cuda.device_count() < gpus:
        env["PIMDB_SHARD_SAME_DEVICE"] = "1"
    r = subprocess.run([str(BIN), "-in", "config.ini", "--dim", str(ndim), "--gpus", str(gpus)], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300, env=env)
    assert "[X]" not in r.stdout, r.stdout
    assert "finished running successfully" in r.stdout, r.stdout + r.stderr
    got = pio.read_simulation_out(str(tmp_path / "output" / "simulation.out"))
    cols = [str(c) for c in E2E[f"{case}/simout_columns"]]
    ref = E2E[f"{case}/simout"]
    assert list(got.keys()) == cols
    for i, c in enumerate(cols):
        assert np.allclose(got[c], ref[:, i], rtol=1e-5), c
        scale = np.max(np.abs(ref[:, i])) + 1e-300
        assert np.max(np.abs(got[c] - ref[:, i])) <= 2e-7 * max(scale, 1.0), (c, got[c], ref[:, i])
    nb = E2E[f"{case}/x"].shape[0]
    for kind, pat in (("x", "position_{}.xyz"), ("v", "velocity_{}.dat"), ("f", "force_{}.dat")):
        key = f"{case}/{kind}"
        if key not in E2E.files:
            continue
        for b in range(nb):
            frames = np.asarray(pio.read_dump_frames(str(tmp_path / "output" / pat.format(b)), ndim))
            assert np.max(np.abs(frames - E2E[key][b])) <= 1e-8 * np.max(np.abs(E2E[key])), (kind, b)


REFFACT = np.load(GOLDEN_DIR / "reffactorial.npz")


@pytest.mark.parametrize("case", ["bosonic_factorial_harmonic", "bosonic_factorial_harmonic_dynamics"])
def test_pimdb_gpu_passes_the_reference_golden_factorial_cases(gpu_required, case, tmp_path):
    """The last two of the reference's twelve golden regression cases belong to its factorial build
    (-DFACTORIAL_BOSONIC_ALGORITHM); `pimdb_gpu --factorial --rng ranmars` runs them unchanged and is judged by the
    reference's own rule (tests/main.py:77-107: np.allclose, rtol 1e-5, every column of the actual simulation.out)."""
    (tmp_path / "config.ini").write_text(str(REFFACT[f"{case}/ini"]))
    r = subprocess.run([str(BIN), "-in", "config.ini", "--dim", "3", "--rng", "ranmars", "--factorial"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=900)
    assert "finished running successfully" in r.stdout, r.stdout + r.stderr
    got = pio.read_simulation_out(str(tmp_path / "output" / "simulation.out"))
    cols = [str(c) for c in REFFACT[f"{case}/simout_columns"]]
    ref = REFFACT[f"{case}/simout"]
    assert set(got.keys()) <= set(cols) and {"step", "kinetic"} <= set(got.keys())
    for c in got:
        i = cols.index(c)
        assert got[c].shape == ref[:, i].shape
        assert np.allclose(got[c], ref[:, i], rtol=1e-5), (c, np.max(np.abs(got[c] - ref[:, i])))
