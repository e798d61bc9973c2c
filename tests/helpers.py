"""Test-side helpers: ctypes wrapper for the C oracle (oracle/liboracle.so), the ref_probe runner
(unmodified reference compiled into oracle/_ref, only in the build container) and input generators.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may touch oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile
from pathlib import Path

import numpy as np

from pimd_b_b200.config import SimConfig

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
REF_DIR = ORACLE_DIR / "_ref"
GOLDEN_DIR = ROOT / "tests" / "golden"

POT_ID = {"free": 0, "aziz": 1, "harmonic": 2, "dipole": 3, "double_well": 4, "cosine": 5}
PROP_ID = {"cartesian": 0, "normal_modes": 1}
THERMO_ID = {"none": 0, "langevin": 1, "nose_hoover": 2, "nose_hoover_np": 3, "nose_hoover_np_dim": 4}


class OrcConfig(C.Structure):
    _fields_ = [
        ("natoms", C.c_int), ("nbeads", C.c_int), ("ndim", C.c_int),
        ("bosonic", C.c_int), ("fixcom", C.c_int), ("pbc", C.c_int),
        ("propagator", C.c_int), ("thermostat", C.c_int), ("nmthermostat", C.c_int),
        ("int_pot", C.c_int), ("ext_pot", C.c_int),
        ("int_omega", C.c_double), ("int_strength", C.c_double), ("ext_omega", C.c_double),
        ("cutoff", C.c_double),
        ("mass", C.c_double), ("temperature", C.c_double), ("dt", C.c_double), ("gamma", C.c_double),
        ("size", C.c_double),
        ("seed", C.c_uint),
        ("nchains", C.c_int),
        ("ext_strength", C.c_double), ("ext_location", C.c_double),
        ("ext_amplitude", C.c_double), ("ext_phase", C.c_double),
        ("factorial", C.c_int),
    ]


class OrcObservables(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "kinetic", "potential", "ext_pot", "int_pot", "virial",
        "temperature", "cl_kinetic", "cl_spring", "prob_dist", "prob_all", "nh_energy", "w_gsf", "pot_gsf")]


_lib = None


def oracle_lib():
    """Load (building if needed) oracle/liboracle.so."""
    global _lib
    if _lib is not None:
        return _lib
    so = ORACLE_DIR / "liboracle.so"
    src = ORACLE_DIR / "pimd_oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(ORACLE_DIR), "oracle"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(str(so))
    lib.orc_create.restype = C.c_void_p
    lib.orc_create.argtypes = [C.POINTER(OrcConfig)]
    lib.orc_destroy.argtypes = [C.c_void_p]
    for name in ("orc_beta", "orc_spring_constant", "orc_cutoff_effective"):
        getattr(lib, name).restype = C.c_double
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.orc_set.argtypes = [C.c_void_p, C.c_char, C.c_void_p]
    lib.orc_get.argtypes = [C.c_void_p, C.c_char, C.c_void_p]
    for name in ("orc_update_forces", "orc_run_iteration", "orc_thermostat_step", "orc_zero_momentum",
                 "orc_propagator_step"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = None
    lib.orc_exchange_get.argtypes = [C.c_void_p, C.c_char, C.c_void_p]
    lib.orc_exchange_get.restype = C.c_int
    lib.orc_observables_calc.argtypes = [C.c_void_p, C.POINTER(OrcObservables)]
    lib.orc_ranmars_new.restype = C.c_void_p
    lib.orc_ranmars_new.argtypes = [C.c_int]
    lib.orc_ranmars_uniform.restype = C.c_double
    lib.orc_ranmars_uniform.argtypes = [C.c_void_p]
    lib.orc_ranmars_gaussian.restype = C.c_double
    lib.orc_ranmars_gaussian.argtypes = [C.c_void_p]
    lib.orc_ranmars_free.argtypes = [C.c_void_p]
    lib.orc_pair_potential.restype = C.c_double
    lib.orc_pair_potential.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]
    lib.orc_exchange_ld.restype = C.c_int
    lib.orc_exchange_ld.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double] + [C.c_void_p] * 9 + [C.c_int]
    _lib = lib
    return lib


def exchange_long_double(cfg: SimConfig, x, l_stride: int = 1, want_prim: bool = True) -> dict:
    """The exchange algorithm in long double on the exterior beads of x [P][N][D] (oracle/pimd_oracle.c orc_exchange_ld):
    V, Vb, the exterior-bead spring forces and the primitive-estimator recursion's e[N] -- the yardstick that tells the
    double-precision algorithm's own rounding noise from a real difference."""
    lib = oracle_lib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    P, N, D = x.shape
    V, Vb = np.empty(N + 1), np.empty(N + 1)
    ff, fl = np.full((N, D), np.nan), np.full((N, D), np.nan)   # l_stride > 1: only every l_stride-th particle is filled
    prim = np.full(1, np.nan)
    sl = [np.ascontiguousarray(x[b]) for b in (0, P - 1, 1 % P, (P - 2) % P)]
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.orc_exchange_ld(N, D, int(cfg.pbc), cfg.size, cfg.spring_constant, cfg.beta / cfg.nbeads,
                             ptr(sl[0]), ptr(sl[1]), ptr(sl[2]), ptr(sl[3]), ptr(V), ptr(Vb), ptr(ff), ptr(fl), ptr(prim) if want_prim else None, int(l_stride))
    assert rc == 0
    return dict(V=V, Vb=Vb, f_first=ff, f_last=fl, prim=float(prim[0]))


def to_orc_config(cfg: SimConfig) -> OrcConfig:
    if cfg.thermostat not in THERMO_ID:
        raise ValueError(f"oracle does not restate thermostat {cfg.thermostat}")
    return OrcConfig(
        natoms=cfg.natoms, nbeads=cfg.nbeads, ndim=cfg.ndim,
        bosonic=int(cfg.bosonic), fixcom=int(cfg.fixcom), pbc=int(cfg.pbc),
        propagator=PROP_ID[cfg.propagator], thermostat=THERMO_ID[cfg.thermostat],
        nmthermostat=int(cfg.nmthermostat),
        int_pot=POT_ID[cfg.interaction], ext_pot=POT_ID[cfg.external],
        int_omega=cfg.int_omega, int_strength=cfg.int_strength, ext_omega=cfg.ext_omega,
        cutoff=cfg.cutoff, mass=cfg.mass, temperature=cfg.temperature, dt=cfg.dt, gamma=cfg.gamma,
        size=cfg.size, seed=cfg.seed, nchains=cfg.nchains,
        ext_strength=cfg.ext_strength, ext_location=cfg.ext_location,
        ext_amplitude=cfg.ext_amplitude, ext_phase=cfg.ext_phase,
        factorial=int(getattr(cfg, "exchange_alg", "quadratic") == "factorial"))


class Oracle:
    """CPU restatement of the reference hot path; state arrays are [P][N][NDIM] float64."""

    def __init__(self, cfg: SimConfig):
        self.cfg = cfg
        self.lib = oracle_lib()
        self._c = to_orc_config(cfg)
        self.h = self.lib.orc_create(C.byref(self._c))
        self.shape = (cfg.nbeads, cfg.natoms, cfg.ndim)

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, which: str, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(self.shape)
        self.lib.orc_set(self.h, which.encode(), a.ctypes.data_as(C.c_void_p))

    def get(self, which: str):
        out = np.empty(self.shape, dtype=np.float64)
        self.lib.orc_get(self.h, which.encode(), out.ctypes.data_as(C.c_void_p))
        return out

    def update_forces(self):
        self.lib.orc_update_forces(self.h)

    def run_iteration(self):
        self.lib.orc_run_iteration(self.h)

    def thermostat_step(self):
        self.lib.orc_thermostat_step(self.h)

    def zero_momentum(self):
        self.lib.orc_zero_momentum(self.h)

    def propagator_step(self):
        self.lib.orc_propagator_step(self.h)

    def exchange(self, which: str):
        n = self.lib.orc_exchange_get(self.h, which.encode(), None)
        out = np.empty(n, dtype=np.float64)
        self.lib.orc_exchange_get(self.h, which.encode(), out.ctypes.data_as(C.c_void_p))
        return out

    def observables(self) -> dict:
        o = OrcObservables()
        self.lib.orc_observables_calc(self.h, C.byref(o))
        return {n: getattr(o, n) for n, _ in OrcObservables._fields_}


# ----------------------------------------------------------------------------- ref_probe (container only)
def ref_probe_available(ndim: int = 3) -> bool:
    return (REF_DIR / f"ref_probe_ndim{ndim}").exists()


def run_ref_probe(cfg: SimConfig, x, p=None, mode="forces", k=1, every=None, timeout=600) -> dict:
    """Run the unmodified reference (oracle/_ref/ref_probe_ndim<d>) on [P][N][NDIM] inputs.

    Returns {name: ndarray} for every *.bin written plus 'obs' / 'exch_scalars' dicts.
    """
    exe = REF_DIR / f"ref_probe_ndim{cfg.ndim}"
    if getattr(cfg, "exchange_alg", "quadratic") == "factorial":      # the reference's compile-time alternative (NDIM = 3 build)
        exe = REF_DIR / "ref_probe_ndim3_factorial"
    P, N, D = cfg.nbeads, cfg.natoms, cfg.ndim
    tmp = Path(tempfile.mkdtemp(prefix="refprobe_"))
    try:
        (tmp / "in").mkdir()
        np.ascontiguousarray(x, dtype=np.float64).reshape(P, N, D).tofile(tmp / "in" / "x.bin")
        if p is not None:
            np.ascontiguousarray(p, dtype=np.float64).reshape(P, N, D).tofile(tmp / "in" / "p.bin")
        (tmp / "cfg.ini").write_text(cfg.to_ini())
        cmd = [str(exe), str(tmp / "cfg.ini"), str(tmp / "in"), str(tmp / "out"), mode]
        if mode == "traj":
            cmd += [str(k), str(every or k)]
        env = dict(os.environ, PIMDB_NP=str(P))
        subprocess.run(cmd, check=True, env=env, cwd=tmp, timeout=timeout,
                       stdout=subprocess.DEVNULL)
        out = {}
        for f in sorted((tmp / "out").iterdir()):
            if f.suffix == ".bin":
                arr = np.fromfile(f, dtype=np.float64)
                if arr.size == P * N * D:
                    arr = arr.reshape(P, N, D)
                out[f.stem] = arr
            elif f.suffix == ".txt":
                d = {}
                for line in f.read_text().splitlines():
                    kname, v = line.split()
                    d[kname] = float(v)
                out[f.stem] = d
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ----------------------------------------------------------------------------- synthetic inputs
KELVIN = 3.1668152e-06
ANGSTROM = 1.8897261
DALTON = 1822.8885
FEMTOSECOND = 1.0e-15 * 4.1341373e16
MEV = 1.0e-3 * 0.036749326


def lattice_positions(cfg: SimConfig, rng: np.random.Generator, spread: float) -> np.ndarray:
    """Cubic-lattice sites replicated over beads plus a Gaussian bead spread."""
    N, P, D, L = cfg.natoms, cfg.nbeads, cfg.ndim, cfg.size
    m = int(np.ceil(N ** (1.0 / D) - 1e-9))
    grid = np.stack(np.meshgrid(*[np.arange(m)] * D, indexing="ij"), axis=-1).reshape(-1, D)[:N]
    sites = (grid + 0.5) * (L / m) - 0.5 * L
    x = np.repeat(sites[None, :, :], P, axis=0) + rng.normal(0.0, spread, size=(P, N, D))
    return np.ascontiguousarray(x)


def maxwell_momenta(cfg: SimConfig, rng: np.random.Generator) -> np.ndarray:
    sigma = np.sqrt(cfg.mass / cfg.thermo_beta)
    return rng.normal(0.0, sigma, size=(cfg.nbeads, cfg.natoms, cfg.ndim))


def relerr(a, b) -> float:
    """max |a-b| / max |b| — the mixed criterion used for FP64 force parity (summation order differs)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b))
    if scale == 0.0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b)) / scale)
