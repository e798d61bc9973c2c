"""DeviceSim: one handle of the C ABI (= the beads one GPU owns), with the reference's method names.

Host arrays use the reference's per-rank ``dVec`` layout concatenated over the owned beads:
``[nbeads_local][natoms][ndim]`` float64 (include/common.h:79-239, include/simulation.h:59-60).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _cabi
from .config import SimConfig


def make_c_config(cfg: SimConfig, bead_begin: int = 0, bead_end: Optional[int] = None, device: int = 0) -> _cabi.PimdbConfig:
    if cfg.interaction not in _cabi.POTENTIAL or cfg.external not in _cabi.POTENTIAL:
        raise ValueError("unknown potential name")
    return _cabi.PimdbConfig(
        natoms=cfg.natoms, nbeads=cfg.nbeads, ndim=cfg.ndim,
        bosonic=int(cfg.bosonic), fixcom=int(cfg.fixcom), pbc=int(cfg.pbc),
        propagator=_cabi.PROPAGATOR[cfg.propagator], thermostat=_cabi.THERMOSTAT[cfg.thermostat],
        nmthermostat=int(cfg.nmthermostat), nchains=cfg.nchains,
        int_potential=_cabi.POTENTIAL[cfg.interaction], ext_potential=_cabi.POTENTIAL[cfg.external],
        int_omega=cfg.int_omega, int_strength=cfg.int_strength,
        ext_omega=cfg.ext_omega, ext_strength=cfg.ext_strength, ext_location=cfg.ext_location,
        ext_amplitude=cfg.ext_amplitude, ext_phase=cfg.ext_phase,
        cutoff=cfg.cutoff, mass=cfg.mass, temperature=cfg.temperature, dt=cfg.dt, gamma=cfg.gamma,
        size=cfg.size, seed=cfg.seed,
        bead_begin=bead_begin, bead_end=cfg.nbeads if bead_end is None else bead_end, device=device,
        rng=_cabi.RNG[getattr(cfg, "rng", "philox")],
        exchange_alg=_cabi.EXCHANGE_ALG[getattr(cfg, "exchange_alg", "quadratic")])


class DeviceSim:
    """GPU-resident bead state + the reference's per-step operations (Simulation / Propagator / Thermostat /
    BosonicExchange / Observable call surface) forwarded to libpimdb200.so."""

    def __init__(self, cfg: SimConfig, bead_begin: int = 0, bead_end: Optional[int] = None, device: int = 0):
        self.cfg = cfg
        self.lib = _cabi.load()
        self._c = make_c_config(cfg, bead_begin, bead_end, device)
        self.bead_begin = self._c.bead_begin
        self.bead_end = self._c.bead_end
        self.nlocal = self.bead_end - self.bead_begin
        self.shape = (self.nlocal, cfg.natoms, cfg.ndim)
        h = C.c_void_p()
        rc = self.lib.pimdb_create(C.byref(self._c), C.byref(h))
        if rc != _cabi.PIMDB_OK:
            _cabi.raise_for_status(self.lib, None, rc)
        self.h = h

    # -- lifecycle
    def close(self):
        if getattr(self, "h", None):
            self.lib.pimdb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _cabi.raise_for_status(self.lib, self.h, rc)

    # -- state (Simulation::coord / momenta / forces)
    def set(self, which: str, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        if a.shape != self.shape:
            raise ValueError(f"expected shape {self.shape}, got {a.shape}")
        self._ck(self.lib.pimdb_set_state(self.h, _cabi.ARRAY[which], a.ctypes.data_as(C.c_void_p)))

    def get(self, which: str, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        self._ck(self.lib.pimdb_get_state(self.h, _cabi.ARRAY[which], out.ctypes.data_as(C.c_void_p)))
        return out

    def _hostptr(self, a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def _checked(self, a):
        if a is None:
            return None
        if a.dtype != np.float64 or not a.flags.c_contiguous or a.shape != self.shape:
            raise ValueError(f"expected a C-contiguous float64 array of shape {self.shape}")
        return a

    def upload(self, x=None, p=None):
        """Several arrays per call: one PCIe copy each, one transpose kernel (pimdb_upload_state). With page-locked arrays the
        copy is only enqueued: keep them untouched until the next synchronising call (download / step_download / get /
        synchronize) has returned. ``set`` waits for the copy."""
        x = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
        p = None if p is None else np.ascontiguousarray(p, dtype=np.float64)
        self._ck(self.lib.pimdb_upload_state(self.h, self._hostptr(self._checked(x)), self._hostptr(self._checked(p))))

    def download(self, x=None, p=None, f=None):
        """Fill the given arrays (coordinates / momenta / forces) with one synchronisation (pimdb_download_state)."""
        self._ck(self.lib.pimdb_download_state(self.h, self._hostptr(self._checked(x)), self._hostptr(self._checked(p)),
                                               self._hostptr(self._checked(f))))

    # -- bead sharding over peer memory (include/pimdb200.h)
    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(_cabi.PEER_BLOB_BYTES)
        self._ck(self.lib.pimdb_peer_export(self.h, buf))
        return buf.raw

    def peer_attach(self, world: int, rank: int, blobs: bytes):
        if len(blobs) != world * _cabi.PEER_BLOB_BYTES:
            raise ValueError("blobs must hold one PEER_BLOB_BYTES record per rank, in rank order")
        self._ck(self.lib.pimdb_peer_attach(self.h, int(world), int(rank), C.c_char_p(blobs)))

    def settle(self):
        """Enqueue the deferred (collective) momentum work without waiting; see pimdb_settle."""
        self._ck(self.lib.pimdb_settle(self.h))

    @property
    def peer_attached(self) -> bool:
        return bool(self.lib.pimdb_peer_attached(self.h))

    # -- the reference's calls
    def update_neighboring_coordinates(self):
        self._ck(self.lib.pimdb_update_neighbors(self.h))

    def update_forces(self):
        self._ck(self.lib.pimdb_update_forces(self.h))

    def moment_step(self):
        self._ck(self.lib.pimdb_moment_step(self.h))

    def coords_step(self):
        self._ck(self.lib.pimdb_coords_step(self.h))

    def propagator_step(self):
        self._ck(self.lib.pimdb_propagator_step(self.h))

    def thermostat_step(self):
        self._ck(self.lib.pimdb_thermostat_step(self.h))

    def zero_momentum(self):
        self._ck(self.lib.pimdb_zero_momentum(self.h))

    def step(self, nsteps: int = 1):
        self._ck(self.lib.pimdb_step(self.h, int(nsteps)))

    def step_download(self, nsteps: int, x=None, p=None, f=None):
        """``step(nsteps)`` + ``download(x, p, f)`` with the copy of x overlapping the last force evaluation."""
        self._ck(self.lib.pimdb_step_download(self.h, int(nsteps), self._hostptr(self._checked(x)),
                                              self._hostptr(self._checked(p)), self._hostptr(self._checked(f))))

    def step_phase(self, phase: int):
        self._ck(self.lib.pimdb_step_phase(self.h, int(phase)))

    def synchronize(self):
        self._ck(self.lib.pimdb_synchronize(self.h))

    # -- bosonic exchange
    def exchange_prepare(self):
        self._ck(self.lib.pimdb_exchange_prepare(self.h))

    def exchange(self, table: str) -> np.ndarray:
        n = self.cfg.natoms
        size = {"V": n + 1, "Vb": n + 1, "E": n * (n + 1) // 2, "prob": n * n}[table]
        out = np.empty(size, dtype=np.float64)
        self._ck(self.lib.pimdb_exchange_get(self.h, _cabi.EXCH_TABLE[table], out.ctypes.data_as(C.c_void_p), size))
        return out

    # -- observables (atomic units, summed over owned beads)
    def observables(self) -> Dict[str, float]:
        o = _cabi.PimdbObservables()
        self._ck(self.lib.pimdb_observables_calc(self.h, C.byref(o)))
        return {n: getattr(o, n) for n in _cabi.OBS_FIELDS}

    # -- plumbing
    @property
    def stream(self) -> int:
        return int(self.lib.pimdb_get_stream(self.h) or 0)

    def set_stream(self, cuda_stream: int):
        self._ck(self.lib.pimdb_set_stream(self.h, C.c_void_p(cuda_stream)))

    def halo_ptr(self, which: int):
        cnt = C.c_size_t()
        p = self.lib.pimdb_halo_ptr(self.h, which, C.byref(cnt))
        return int(p), int(cnt.value)

    def com_ptr(self) -> int:
        return int(self.lib.pimdb_com_ptr(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.pimdb_launch_count(self.h))

    def timing_enable(self, on: bool):
        self._ck(self.lib.pimdb_timing_enable(self.h, int(on)))

    def timing_integrator_bytes(self) -> float:
        b = C.c_double()
        self._ck(self.lib.pimdb_timing_integrator_bytes(self.h, C.byref(b)))
        return b.value

    def timing_get(self, what: int):
        ms = C.c_double()
        cnt = C.c_ulonglong()
        self._ck(self.lib.pimdb_timing_get(self.h, what, C.byref(ms), C.byref(cnt)))
        return ms.value, int(cnt.value)
