"""ctypes binding of the C ABI in include/pimdb200.h (libpimdb200.so, built in-tree by pimd_b_b200.build).

There is deliberately no fallback: if the shared object is missing or the device is unusable the import /
constructor raises.  Status codes are mapped back to the exception types the reference throws
(src/pimdb.cpp:57-63): std::invalid_argument -> ValueError, std::overflow_error -> OverflowError,
std::runtime_error / CUDA failures -> RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# PIMDB200_LIB points at an alternative build of the same ABI (kernel experiments); the default is the in-tree library
LIB_PATH = Path(os.environ.get("PIMDB200_LIB") or Path(__file__).resolve().parent / "libpimdb200.so")

PIMDB_OK, ERR_INVALID_ARGUMENT, ERR_OVERFLOW, ERR_RUNTIME, ERR_CUDA = range(5)
ABI_VERSION = 2
PEER_BLOB_BYTES = 256

POTENTIAL = {"free": 0, "aziz": 1, "harmonic": 2, "dipole": 3, "double_well": 4, "cosine": 5}
PROPAGATOR = {"cartesian": 0, "normal_modes": 1}
RNG = {"philox": 0, "ranmars": 1}
EXCHANGE_ALG = {"quadratic": 0, "factorial": 1}
THERMOSTAT = {"none": 0, "langevin": 1, "nose_hoover": 2, "nose_hoover_np": 3, "nose_hoover_np_dim": 4}
ARRAY = {"x": 0, "p": 1, "f": 2, "f_spring": 3, "f_phys": 4}
EXCH_TABLE = {"V": 0, "Vb": 1, "E": 2, "prob": 3}


class PimdbConfig(C.Structure):
    _fields_ = [
        ("natoms", C.c_int), ("nbeads", C.c_int), ("ndim", C.c_int),
        ("bosonic", C.c_int), ("fixcom", C.c_int), ("pbc", C.c_int),
        ("propagator", C.c_int), ("thermostat", C.c_int), ("nmthermostat", C.c_int), ("nchains", C.c_int),
        ("int_potential", C.c_int), ("ext_potential", C.c_int),
        ("int_omega", C.c_double), ("int_strength", C.c_double),
        ("ext_omega", C.c_double), ("ext_strength", C.c_double), ("ext_location", C.c_double),
        ("ext_amplitude", C.c_double), ("ext_phase", C.c_double),
        ("cutoff", C.c_double),
        ("mass", C.c_double), ("temperature", C.c_double), ("dt", C.c_double), ("gamma", C.c_double),
        ("size", C.c_double),
        ("seed", C.c_ulonglong),
        ("bead_begin", C.c_int), ("bead_end", C.c_int),
        ("device", C.c_int),
        ("rng", C.c_int),
        ("exchange_alg", C.c_int),
        ("reserved", C.c_int * 2),
    ]


class PimdbObservables(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "kinetic", "potential", "ext_pot", "int_pot", "virial",
        "temperature", "cl_kinetic", "cl_spring", "prob_dist", "prob_all", "nh_energy", "w_gsf", "pot_gsf")] + [
            ("reserved", C.c_double * 3)]


OBS_FIELDS = tuple(n for n, _ in PimdbObservables._fields_ if n != "reserved")

# every symbol include/pimdb200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "pimdb_abi_version": (C.c_int, []),
    "pimdb_create": (C.c_int, [C.POINTER(PimdbConfig), C.POINTER(_VP)]),
    "pimdb_destroy": (None, [_VP]),
    "pimdb_last_error": (C.c_char_p, [_VP]),
    "pimdb_set_state": (C.c_int, [_VP, C.c_int, _VP]),
    "pimdb_get_state": (C.c_int, [_VP, C.c_int, _VP]),
    "pimdb_upload_state": (C.c_int, [_VP, _VP, _VP]),
    "pimdb_download_state": (C.c_int, [_VP, _VP, _VP, _VP]),
    "pimdb_update_neighbors": (C.c_int, [_VP]),
    "pimdb_update_forces": (C.c_int, [_VP]),
    "pimdb_moment_step": (C.c_int, [_VP]),
    "pimdb_coords_step": (C.c_int, [_VP]),
    "pimdb_propagator_step": (C.c_int, [_VP]),
    "pimdb_thermostat_step": (C.c_int, [_VP]),
    "pimdb_zero_momentum": (C.c_int, [_VP]),
    "pimdb_step": (C.c_int, [_VP, C.c_int]),
    "pimdb_step_download": (C.c_int, [_VP, C.c_int, _VP, _VP, _VP]),
    "pimdb_synchronize": (C.c_int, [_VP]),
    "pimdb_exchange_prepare": (C.c_int, [_VP]),
    "pimdb_exchange_get": (C.c_int, [_VP, C.c_int, _VP, C.c_size_t]),
    "pimdb_observables_calc": (C.c_int, [_VP, C.POINTER(PimdbObservables)]),
    "pimdb_get_stream": (_VP, [_VP]),
    "pimdb_set_stream": (C.c_int, [_VP, _VP]),
    "pimdb_halo_ptr": (_VP, [_VP, C.c_int, C.POINTER(C.c_size_t)]),
    "pimdb_com_ptr": (_VP, [_VP]),
    "pimdb_step_phase": (C.c_int, [_VP, C.c_int]),
    "pimdb_peer_export": (C.c_int, [_VP, _VP]),
    "pimdb_peer_attach": (C.c_int, [_VP, C.c_int, C.c_int, _VP]),
    "pimdb_peer_attached": (C.c_int, [_VP]),
    "pimdb_settle": (C.c_int, [_VP]),
    "pimdb_launch_count": (C.c_ulonglong, [_VP]),
    "pimdb_timing_enable": (C.c_int, [_VP, C.c_int]),
    "pimdb_timing_get": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]),
    "pimdb_timing_integrator_bytes": (C.c_int, [_VP, C.POINTER(C.c_double)]),
    "pimdb_bench_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """dlopen libpimdb200.so and type every entry point. Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m pimd_b_b200.build` "
            "(there is no CPU fallback for the PIMD hot path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def raise_for_status(lib, handle, rc: int):
    if rc == PIMDB_OK:
        return
    msg = lib.pimdb_last_error(handle)
    text = msg.decode() if msg else f"pimdb error {rc}"
    if rc == ERR_INVALID_ARGUMENT:
        raise ValueError(text)
    if rc == ERR_OVERFLOW:
        raise OverflowError(text)
    raise RuntimeError(text)
