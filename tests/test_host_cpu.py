"""CPU-side checks (no GPU): the INI/units mirror of the reference's Params, the text formats either side of the
hot path, the C ABI surface (every symbol include/pimdb200.h declares is exported and typed), loud failure without
a device, and the documented Philox4x32-10 noise stream (known-answer vectors)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from pimd_b_b200 import _cabi, io as pio
from pimd_b_b200.config import (SimConfig, convert_to_internal, convert_to_user, get_quantity, parse_ini,
                                separate_prefix_unit)
from tests.helpers import GOLDEN_DIR, ROOT

REFCASES = np.load(GOLDEN_DIR / "refcases.npz")


# ----------------------------------------------------------------------------- config / units
def test_units_match_reference_tables():
    assert convert_to_internal("length", "angstrom", 1.0) == 1.8897261
    assert convert_to_internal("time", "femtosecond", 1.0) == pytest.approx(1e-15 * 4.1341373e16)
    assert convert_to_internal("energy", "millielectronvolt", 3.0) == pytest.approx(3e-3 * 0.036749326)
    assert convert_to_internal("mass", "dalton", 1.0) == 1822.8885
    assert convert_to_user("temperature", "kelvin", 3.1668152e-06) == pytest.approx(1.0)
    assert separate_prefix_unit("picometer") == ("pico", "meter")
    assert separate_prefix_unit("atomic_unit") == ("", "atomic_unit")
    assert get_quantity("length", "300.0 atomic_unit") == 300.0
    with pytest.raises(ValueError, match="undefined unit for kind length"):
        convert_to_internal("length", "parsec", 1.0)


@pytest.mark.parametrize("case", ["bosonic_quadratic_harmonic_dynamics", "dist_harmonic_nm_propagation_dynamics",
                                  "bosonic_quadratic_harmonic_nmthermostat_dynamics", "dist_harmonic"])
def test_parse_reference_golden_inis(case, tmp_path):
    p = tmp_path / "c.ini"
    p.write_text(str(REFCASES[f"{case}/ini"]))
    cfg = parse_ini(str(p))
    assert cfg.nbeads == 8 and cfg.natoms in (8, 12)          # "nbeads = 8.0" -> 8 (atoi semantics)
    assert cfg.steps == 100000 and cfg.threshold == 0.0
    assert cfg.fixcom is False and cfg.pbc is False
    assert cfg.external == "harmonic" and cfg.interaction == "free" and cfg.cutoff == 0.0
    assert cfg.ext_omega == pytest.approx(3e-3 * 0.036749326)
    assert cfg.gamma == pytest.approx(1.0 / (100.0 * cfg.dt))  # default friction
    assert cfg.bosonic == case.startswith("bosonic")
    assert cfg.propagator == ("normal_modes" if "nm_propagation" in case else "cartesian")
    assert cfg.nmthermostat == ("nmthermostat" in case)


def test_ini_round_trip_and_validation(tmp_path):
    cfg = SimConfig(nbeads=6, natoms=5, ndim=2, bosonic=True, pbc=True, interaction="dipole", int_strength=2.5,
                    external="harmonic", ext_omega=1e-4, thermostat="langevin", cutoff=7.5, size=33.0,
                    temperature=1e-5, mass=7000.0, dt=40.0, seed=77)
    p = tmp_path / "rt.ini"
    p.write_text(cfg.to_ini())
    back = parse_ini(str(p), ndim=2)
    for k in ("nbeads", "natoms", "bosonic", "pbc", "interaction", "int_strength", "external", "ext_omega", "cutoff",
              "size", "temperature", "mass", "dt", "seed", "gamma", "fixcom"):
        assert getattr(back, k) == getattr(cfg, k), k
    bad = SimConfig(nbeads=4, bosonic=True, propagator="normal_modes", thermostat="langevin")
    with pytest.raises(ValueError, match="Normal modes propogation is currently not available for bosons!"):
        bad.validate()
    with pytest.raises(ValueError, match="nmthermostat cannot be used in nve ensemble!"):
        SimConfig(nmthermostat=True, thermostat="none").validate()
    p.write_text("[simulation]\nnbeads = 4\n")
    with pytest.raises(ValueError, match="Thermostat must be specified!"):
        parse_ini(str(p))
    p.write_text("[simulation]\nthermostat = langevin\nnchains = 3\n")
    with pytest.raises(ValueError, match="nchains can only be used with Nose-Hoover thermostats!"):
        parse_ini(str(p))
    # cutoff clamp and "free" rule of the Simulation constructor
    assert SimConfig(interaction="aziz", pbc=True, size=10.0, cutoff=8.0).cutoff_effective == 5.0
    assert SimConfig(interaction="aziz", pbc=True, size=10.0, cutoff=-1.0).cutoff_effective == -1.0
    assert SimConfig(interaction="free", cutoff=3.0).cutoff_effective == 0.0


# ----------------------------------------------------------------------------- text formats
def test_dump_and_logger_formats(tmp_path):
    arr = np.array([[2.130326814812, 38.41428738278, -39.00125211552]])
    frame = pio.format_frame("position", 0, arr, 3)
    assert frame.splitlines()[2] == "1  2.130326814812e+00   3.841428738278e+01  -3.900125211552e+01 "
    frame = pio.format_frame("force", 1000, np.array([[1.086635523979e-04, -1.148003115387e-04]]), 2)
    assert frame.splitlines()[1] == "Step 1000"
    assert frame.splitlines()[2] == "1 1  1.086635523979e-04  -1.148003115387e-04  0.0"
    log = pio.ObservablesLogger(["kinetic", "potential"], folder=str(tmp_path))
    log.log(1000, {"kinetic": 189.234521, "potential": 164.123278})
    log.close()
    lines = (tmp_path / "simulation.out").read_text().splitlines()
    assert lines[0] == "      step           kinetic         potential    "
    assert lines[1] == " 1.00000000e+03   1.89234521e+02   1.64123278e+02 "
    back = pio.read_simulation_out(str(tmp_path / "simulation.out"))
    assert back["kinetic"][0] == pytest.approx(189.234521)


def test_initial_state_files_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.normal(size=(5, 3)) * 10
    p = rng.normal(size=(5, 3))
    pio.write_xyz_positions(str(tmp_path / "pos_0.xyz"), x)
    pio.write_manual_velocities(str(tmp_path / "vel_0.dat"), p, 7296.0)
    assert np.allclose(pio.load_xyz_positions(str(tmp_path / "pos_0.xyz"), 5, 3), x, rtol=1e-15)
    assert np.allclose(pio.load_manual_momenta(str(tmp_path / "vel_0.dat"), 5, 3, 7296.0), p, rtol=1e-15)
    with pytest.raises(RuntimeError, match="does not match the requested number of atoms"):
        pio.load_xyz_positions(str(tmp_path / "pos_0.xyz"), 6, 3)
    with pytest.raises(RuntimeError, match="Cannot open the xyz file"):
        pio.load_xyz_positions(str(tmp_path / "nope.xyz"), 5, 3)


# ----------------------------------------------------------------------------- C ABI surface
def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "pimdb200.h").read_text()
    declared = set(re.findall(r"\b(pimdb_[a-z0-9_]+)\s*\(", header))
    declared -= {"pimdb_config", "pimdb_observables", "pimdb_sim"}
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    lib = _cabi.load()   # raises if a declared symbol is missing from libpimdb200.so
    assert lib.pimdb_abi_version() == _cabi.ABI_VERSION == 2
    assert C.sizeof(_cabi.PimdbConfig) == 12 * 4 + 13 * 8 + 8 + 3 * 4 + 4 * 4 + 4   # matches the C struct layout (+ tail padding)


def test_create_fails_loudly_without_device_or_with_bad_config():
    import ctypes
    lib = _cabi.load()
    from pimd_b_b200.engine import make_c_config
    h = ctypes.c_void_p()
    bad = make_c_config(SimConfig(nbeads=4, natoms=4, bosonic=True, propagator="normal_modes"))
    rc = lib.pimdb_create(ctypes.byref(bad), ctypes.byref(h))
    assert rc == _cabi.ERR_INVALID_ARGUMENT and not h.value
    assert b"Normal modes propogation" in lib.pimdb_last_error(None)
    ok = make_c_config(SimConfig(nbeads=4, natoms=4))
    rc = lib.pimdb_create(ctypes.byref(ok), ctypes.byref(h))
    if rc != _cabi.PIMDB_OK:      # CPU container: must be a CUDA error with a message, never a silent fallback
        assert rc == _cabi.ERR_CUDA and not h.value
        assert b"no CPU fallback" in lib.pimdb_last_error(None) or b"CUDA" in lib.pimdb_last_error(None)
    else:
        lib.pimdb_destroy(h)


# ----------------------------------------------------------------------------- documented noise stream
def philox4x32_10(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def test_philox_known_answer_vectors():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_ctypes_structs_match_the_c_header(tmp_path):
    """pimdb_config / pimdb_observables as declared in include/pimdb200.h (compiled with gcc) and as mirrored in
    pimd_b_b200/_cabi.py must agree in size and in the offset of every field."""
    import ctypes as C
    import shutil
    import subprocess
    from pimd_b_b200 import _cabi
    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        pytest.skip("no C compiler")
    fields_cfg = [n for n, _ in _cabi.PimdbConfig._fields_]
    fields_obs = [n for n, _ in _cabi.PimdbObservables._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "pimdb200.h"', 'int main(void) {',
           'printf("config %zu\\n", sizeof(pimdb_config));', 'printf("observables %zu\\n", sizeof(pimdb_observables));']
    src += [f'printf("config.{n} %zu\\n", offsetof(pimdb_config, {n}));' for n in fields_cfg]
    src += [f'printf("observables.{n} %zu\\n", offsetof(pimdb_observables, {n}));' for n in fields_obs]
    src += ['return 0; }']
    (tmp_path / "t.c").write_text("\n".join(src))
    subprocess.check_call([gcc, "-I", str(ROOT / "include"), "-o", str(tmp_path / "t"), str(tmp_path / "t.c")])
    out = dict(line.split() for line in subprocess.check_output([str(tmp_path / "t")], text=True).splitlines())
    assert int(out["config"]) == C.sizeof(_cabi.PimdbConfig)
    assert int(out["observables"]) == C.sizeof(_cabi.PimdbObservables)
    for n in fields_cfg:
        assert int(out[f"config.{n}"]) == getattr(_cabi.PimdbConfig, n).offset, n
    for n in fields_obs:
        assert int(out[f"observables.{n}"]) == getattr(_cabi.PimdbObservables, n).offset, n
