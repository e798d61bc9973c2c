"""The blocked exchange recurrence (csrc/exchange.cu, k_exch_recur_cluster / k_exch_recur_cluster_multi: 32 unknowns per chain step through the
precomputed inverses of the diagonal blocks) against the CPU oracle, through the C ABI: V, V_backwards, exterior
spring forces and connection probabilities at block-boundary sizes, plus which path (matrix-vector product / exact
sequential steps) every block actually took.

Reference: src/bosonic_exchange/quadratic_bosonic_exchange.cpp:73-128 (the two recursions), :142-215.
"""
import ctypes as C

import numpy as np
import pytest

from pimd_b_b200.engine import DeviceSim
from tests.helpers import DALTON, KELVIN, Oracle, relerr
from tests.test_gpu_parity import ENERGY_TOL, FORCE_TOL, helium, make_inputs, trap

pytestmark = pytest.mark.gpu


def block_status(sim):
    """(forward, backward) lists of per-block status in step order: 1 = G*rho, 2 = exact sequential, 0 = no steps."""
    fn = sim.lib.pimdb_debug_exchange_blocks
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    buf = (C.c_int * 1024)()
    nb = fn(sim.h, buf)
    return list(buf[:nb]), list(buf[nb:2 * nb])


def check(cfg, x, expect=None, force_tol=FORCE_TOL):
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.update_forces()
    sim = DeviceSim(cfg)
    sim.set("x", x)
    sim.update_forces()
    assert relerr(sim.exchange("V"), orc.exchange("V")) < ENERGY_TOL
    assert relerr(sim.exchange("Vb"), orc.exchange("B")) < ENERGY_TOL
    assert relerr(sim.get("f_spring"), orc.get("s")) < force_tol
    n = cfg.natoms
    prob = sim.exchange("prob").reshape(n, n)
    assert np.max(np.abs(prob - orc.exchange("P").reshape(n, n))) < 1e-10 * (force_tol / FORCE_TOL)
    fwd, bwd = block_status(sim)
    sim.close()
    nb = (n + 31) // 32
    assert len(fwd) == nb and all(s in (1, 2) for s in fwd)
    assert all(s in (1, 2) for s in bwd) if n > 1 else all(s == 0 for s in bwd)
    if expect is not None:
        assert all(s == expect for s in fwd + (bwd if n > 1 else [])), (fwd, bwd)
    return fwd, bwd


@pytest.mark.parametrize("natoms", [1, 2, 3, 5, 31, 32, 33, 34, 63, 64, 65, 95, 97, 128, 129, 255, 257, 300, 449, 481, 511, 512,
                                    513, 700, 1023, 1024, 1025, 1500, 2047, 2048, 2049, 2100, 3000, 4127])
def test_blocked_recurrence_sizes(gpu_required, natoms):
    """Correlated ring polymers in a trap: every block takes the matrix-vector path. Sizes straddle the 32-row block
    boundaries (ragged first / last blocks in either direction), the 512-particle limit of the tiles' own prefix sums and
    the warps-per-cluster-block steps (1 .. 8 warps, N <= 2048)."""
    cfg = trap(natoms, 3, temperature=1.0 * KELVIN, size=2000.0)
    rng = np.random.default_rng(natoms)
    centroid = rng.normal(0.0, 60.0, size=(1, natoms, 3))
    x = np.repeat(centroid, 3, axis=0) + rng.normal(0.0, 6.0, size=(3, natoms, 3))
    check(cfg, x, expect=1)


@pytest.mark.parametrize("natoms,pbc", [(80, False), (200, True), (333, False), (512, True), (1100, False), (2500, False)])
def test_blocked_recurrence_exact_blocks(gpu_required, natoms, pbc):
    """Stiff springs + uncorrelated beads (beta*E ~ 1e3-1e4 per link): the block inverses or the new values leave the
    plain-double window and the blocks are redone exactly; consumers then apply them column by column."""
    cfg = trap(natoms, 3, mass=4.0026 * DALTON, temperature=2 * KELVIN, size=60.0, pbc=pbc,
               external="free" if pbc else "harmonic")
    x, _ = make_inputs(cfg, 11 + natoms, 1.0)
    # (N = 2500: the reference's own probabilities exp(-beta (V[u] + E + Vb - V[N])) difference energies of order 1e5
    # there and carry ~1e-10 themselves; V and V_backwards still agree to 1e-10)
    fwd, bwd = check(cfg, x, force_tol=FORCE_TOL if natoms < 2000 else 1e-9)
    assert 2 in fwd + bwd, (fwd, bwd)


def test_blocked_recurrence_mixed_blocks(gpu_required):
    """Half of the particles correlated, half scattered: fast and exact blocks alternate inside one recurrence, so
    both consumer paths and both hand-off formats are exercised in one run."""
    natoms = 256
    cfg = trap(natoms, 3, mass=4.0026 * DALTON, temperature=2 * KELVIN, size=60.0)
    rng = np.random.default_rng(5)
    centroid = rng.normal(0.0, 10.0, size=(1, natoms, 3))
    x = np.repeat(centroid, 3, axis=0) + rng.normal(0.0, 0.02, size=(3, natoms, 3))
    scattered = (np.arange(natoms) // 32) % 2 == 1
    x[:, scattered, :] = rng.uniform(-30.0, 30.0, size=(3, int(scattered.sum()), 3))
    fwd, bwd = check(cfg, x)
    assert 1 in fwd + bwd and 2 in fwd + bwd, (fwd, bwd)


def test_blocked_recurrence_helium_c3_slice(gpu_required):
    """The headline system's exchange problem (N = 512 He-4 atoms, PBC, lattice + bead spread): all blocks fast."""
    cfg = helium(512, 4)
    x, _ = make_inputs(cfg, 3, "lattice")
    check(cfg, x, expect=1)


def test_blocked_equals_scalar_recurrence(gpu_required, monkeypatch):
    """Same positions through the blocked kernel and through the scalar one-unknown-per-step kernel (PIMDB_EXCH_NOBLOCKED):
    V and the forces agree to rounding."""
    cfg = trap(300, 4, temperature=1.0 * KELVIN, size=2000.0)
    rng = np.random.default_rng(8)
    x = np.repeat(rng.normal(0.0, 60.0, size=(1, 300, 3)), 4, axis=0) + rng.normal(0.0, 6.0, size=(4, 300, 3))
    out = {}
    for mode in ("blocked", "scalar"):
        if mode == "scalar":
            monkeypatch.setenv("PIMDB_EXCH_NOBLOCKED", "1")
        sim = DeviceSim(cfg)
        sim.set("x", x)
        sim.update_forces()
        out[mode] = (sim.exchange("V"), sim.exchange("Vb"), sim.get("f_spring"))
        sim.close()
    for a, b in zip(out["blocked"], out["scalar"]):
        assert relerr(a, b) < 1e-12
