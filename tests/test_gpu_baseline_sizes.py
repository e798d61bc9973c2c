"""Parity with the oracle at the full size of every BASELINE.json configuration (C2 .. C5), on hardware, through the C ABI.

C3: all 64 beads (interior and exterior) + a short NVE trajectory, with a per-component criterion next to the max-norm
one; C2: N=64, P=64, 2-D dipoles, normal-mode propagator + normal-mode Langevin thermostat with the reference's own
noise; C4: pair tiles and the cluster recurrence at N=2048 in the helium geometry on a 4-bead ring with C4's spring
constant; C5: the four-blocks-per-warp cluster recurrence and the exterior forces at N=8192.
Reference: BASELINE.json `configs`; src/simulation.cpp:353-455, src/bosonic_exchange/quadratic_bosonic_exchange.cpp:34-215.
"""
import dataclasses

import numpy as np
import pytest

from pimd_b_b200 import workloads as wl
from pimd_b_b200.config import SimConfig
from pimd_b_b200.engine import DeviceSim
from tests.helpers import Oracle, exchange_long_double, relerr

pytestmark = pytest.mark.gpu


def per_component_ok(got, ref, rel=1e-10, floor=1e-11):
    """|got - ref| <= rel |ref| + floor max|ref| for every component: a small force component may not hide behind the
    largest one (the floor covers the summation-order noise of components that are sums of cancelling pair terms)."""
    got, ref = np.asarray(got), np.asarray(ref)
    bad = np.abs(got - ref) > rel * np.abs(ref) + floor * np.max(np.abs(ref))
    return not bad.any(), int(bad.sum())


def check_exterior_forces(sim, orc, cfg, x, l_stride=1, want_prim=True):
    """Exterior-bead spring (exchange) forces. The reference's double-precision log-sum-exp loses digits where beta*V is
    large, so at N >= 512 its own connection probabilities carry 1e-11 .. 1e-9 of rounding noise (the oracle, bit-identical
    to the reference, inherits it). The yardstick is the same algorithm in long double: the GPU path must match IT to
    1e-10, and may differ from the double-precision oracle by no more than the oracle itself differs from the yardstick."""
    ld = exchange_long_double(cfg, x, l_stride=l_stride, want_prim=want_prim)
    fs, fo = sim.get("f_spring"), orc.get("s")
    sel = slice(None, None, l_stride)
    scale = max(np.max(np.abs(ld["f_first"][sel])), np.max(np.abs(ld["f_last"][sel])))
    noise = 0.0
    for b, key in ((0, "f_first"), (cfg.nbeads - 1, "f_last")):
        err_gpu = np.max(np.abs(fs[b][sel] - ld[key][sel])) / scale
        err_orc = np.max(np.abs(fo[b][sel] - ld[key][sel])) / scale
        assert err_gpu < 1e-10, (key, err_gpu, err_orc)
        ok, nbad = per_component_ok(fs[b][sel], ld[key][sel])
        assert ok, (key, nbad)
        noise = max(noise, err_orc)
    assert relerr(sim.exchange("V"), ld["V"]) < 1e-12 and relerr(sim.exchange("Vb"), ld["Vb"]) < 1e-12
    return noise, ld


def shrink_beads(cfg, nbeads):
    """The same system on a shorter ring with the spring constant m (P/beta)^2 and exchange beta/P of the full one."""
    sub = SimConfig(**{**cfg.as_dict(), "nbeads": nbeads})
    sub.temperature = cfg.temperature * cfg.nbeads / nbeads
    return sub


def test_c3_full_size_forces_all_beads_and_nve_steps(gpu_required):
    cfg = dataclasses.replace(wl.config("c3"), thermostat="none")
    x, p = wl.initial_state(cfg, "c3")
    sim, orc = DeviceSim(cfg), Oracle(cfg)
    sim.upload(x, p)
    orc.set("x", x); orc.set("p", p)
    sim.update_forces(); orc.update_forces()
    f, fr = sim.get("f"), orc.get("f")
    assert relerr(f, fr) < 1e-10
    ok, nbad = per_component_ok(f[1:63], fr[1:63])                # interior beads: every component
    assert ok, nbad
    ok, nbad = per_component_ok(sim.get("f_phys"), orc.get("e"))  # pair forces of all 64 beads: every component
    assert ok, nbad
    noise, _ = check_exterior_forces(sim, orc, cfg, x)            # exterior beads: against the long-double yardstick
    assert relerr(sim.get("f_spring"), orc.get("s")) < 1e-10 + 2 * noise
    for w, k in (("V", "V"), ("Vb", "B")):
        assert relerr(sim.exchange(w), orc.exchange(k)) < 1e-10
    sim.step(5)
    for _ in range(5):
        orc.run_iteration()
    assert relerr(sim.get("x"), orc.get("x")) < 1e-10
    assert relerr(sim.get("p"), orc.get("p")) < 1e-9
    ok, nbad = per_component_ok(sim.get("f"), orc.get("f"), rel=1e-9, floor=1e-10)
    assert ok, nbad
    o, r = sim.observables(), orc.observables()
    for k in ("kinetic", "potential", "virial", "cl_kinetic", "cl_spring", "temperature"):
        assert abs(o[k] - r[k]) <= 1e-9 * abs(r[k]), (k, o[k], r[k])
    sim.close(); orc.close()


def test_c2_full_size_normal_modes_with_reference_noise(gpu_required):
    cfg = dataclasses.replace(wl.config("c2"), rng="ranmars")
    assert (cfg.natoms, cfg.nbeads, cfg.ndim) == (64, 64, 2) and cfg.propagator == "normal_modes" and cfg.nmthermostat
    x, p = wl.initial_state(cfg, "c2")
    sim, orc = DeviceSim(cfg), Oracle(cfg)
    sim.upload(x, p)
    orc.set("x", x); orc.set("p", p)
    # (no force evaluation ahead of the first step: the reference's NormalModesPropagator kicks with its own copy of the
    # physical forces, which is zero until its first step has run -- normal_modes_propagator.cpp:19-71)
    sim.step(10)
    for _ in range(10):
        orc.run_iteration()
    assert relerr(sim.get("x"), orc.get("x")) < 1e-10
    assert relerr(sim.get("p"), orc.get("p")) < 1e-9
    f, fr = sim.get("f"), orc.get("f")
    assert relerr(f, fr) < 1e-9
    o, r = sim.observables(), orc.observables()
    for k in ("kinetic", "potential", "ext_pot", "int_pot", "virial", "cl_kinetic", "cl_spring", "temperature"):
        assert abs(o[k] - r[k]) <= 1e-9 * abs(r[k]) + 1e-18, (k, o[k], r[k])
    # forces and estimators on identical positions: the 1e-10 criterion, every component
    sim.upload(orc.get("x"), orc.get("p"))
    sim.update_forces()
    f = sim.get("f")
    assert relerr(f, fr) < 1e-10
    ok, nbad = per_component_ok(f, fr)
    assert ok, nbad
    o = sim.observables()
    for k in ("kinetic", "potential", "ext_pot", "int_pot", "virial", "cl_kinetic", "cl_spring", "temperature"):
        assert abs(o[k] - r[k]) <= 1e-10 * abs(r[k]) + 1e-18, (k, o[k], r[k])
    sim.close(); orc.close()


def test_c2_full_size_nve_normal_mode_propagator(gpu_required):
    """The propagator alone (thermostat off): 10 exact free-ring rotations + kicks at P = 64."""
    cfg = dataclasses.replace(wl.config("c2"), thermostat="none", nmthermostat=False)
    x, p = wl.initial_state(cfg, "c2")
    sim, orc = DeviceSim(cfg), Oracle(cfg)
    sim.upload(x, p)
    orc.set("x", x); orc.set("p", p)
    sim.step(10)
    for _ in range(10):
        orc.run_iteration()
    assert relerr(sim.get("x"), orc.get("x")) < 1e-10
    assert relerr(sim.get("p"), orc.get("p")) < 1e-9
    sim.close(); orc.close()


def test_c4_particle_count_pair_tiles_and_exchange(gpu_required):
    """N = 2048 in the helium geometry of C4 on a 4-bead ring: 64 x 64 pair tiles per bead, prefix kernel + factor tiles,
    cluster recurrence with 8 warps per block, exterior forces -- every bead against the oracle."""
    full = wl.config("c4")
    cfg = dataclasses.replace(shrink_beads(full, 4), thermostat="none")
    x, p = wl.initial_state(full, "c4")
    x, p = np.ascontiguousarray(x[:4]), np.ascontiguousarray(p[:4])
    sim, orc = DeviceSim(cfg), Oracle(cfg)
    sim.upload(x, p)
    orc.set("x", x); orc.set("p", p)
    sim.update_forces(); orc.update_forces()
    f, fr = sim.get("f"), orc.get("f")
    ok, nbad = per_component_ok(f[1:3], fr[1:3])
    assert ok, nbad
    ok, nbad = per_component_ok(sim.get("f_phys"), orc.get("e"))
    assert ok, nbad
    noise, _ = check_exterior_forces(sim, orc, cfg, x, want_prim=False)
    assert relerr(f, fr) < 1e-10 + 2 * noise
    for w, k in (("V", "V"), ("Vb", "B"), ("E", "E")):
        assert relerr(sim.exchange(w), orc.exchange(k)) < 1e-10, w
    prob = sim.exchange("prob").reshape(cfg.natoms, cfg.natoms)
    assert np.max(np.abs(prob - orc.exchange("P").reshape(cfg.natoms, cfg.natoms))) < 1e-10 + 2 * noise
    assert np.max(np.abs(prob.sum(axis=1) - 1.0)) < 1e-12        # (the reference's own rows: 3e-12)
    sim.step(3)
    for _ in range(3):
        orc.run_iteration()
    assert relerr(sim.get("x"), orc.get("x")) < 1e-10
    assert relerr(sim.get("p"), orc.get("p")) < 1e-9 + 10 * noise
    sim.close(); orc.close()


@pytest.mark.slow
def test_c5_particle_count_exchange(gpu_required):
    """N = 8192 free bosons in the trap (C5) on a 2-bead ring: k_exch_recur_cluster_multi with FOUR row blocks per warp
    (N > 6144), weights-only staging of the exterior forces. V, V_backwards and the forces of both (exterior) beads."""
    full = wl.config("c5")
    cfg = dataclasses.replace(shrink_beads(full, 2), thermostat="none")
    x, p = wl.initial_state(full, "c5")
    x, p = np.ascontiguousarray(x[:2]), np.ascontiguousarray(p[:2])
    sim, orc = DeviceSim(cfg), Oracle(cfg)
    sim.upload(x, p)
    orc.set("x", x); orc.set("p", p)
    sim.update_forces(); orc.update_forces()
    for w, k in (("V", "V"), ("Vb", "B")):
        assert relerr(sim.exchange(w), orc.exchange(k)) < 1e-10, w
    # every 16th particle of both beads against the long-double yardstick (its O(N^2) expl() calls are what costs here)
    noise, ld = check_exterior_forces(sim, orc, cfg, x, l_stride=16)
    f, fr = sim.get("f"), orc.get("f")
    assert relerr(f, fr) < 1e-10 + 2 * noise
    assert relerr(sim.get("f_phys"), orc.get("e")) < 1e-12          # the trap
    o, r = sim.observables(), orc.observables()
    # the primitive estimator's e[N] (energy.cpp:30-48) shares the recursion's rounding noise: GPU and oracle differ by the
    # same amount in e[N] as in the kinetic column, and the GPU value is the one next to the yardstick's
    assert abs(o["kinetic"] - r["kinetic"]) <= (1e-10 + 10 * noise) * abs(r["kinetic"]), (o["kinetic"], r["kinetic"])
    for k in ("cl_spring", "prob_dist", "prob_all"):
        assert abs(o[k] - r[k]) <= 1e-9 * abs(r[k]) + 1e-300, (k, o[k], r[k])
    sim.close(); orc.close()


def test_step_zero_before_the_first_step_is_a_no_op(gpu_required):
    """pimdb_step(0) as the first step call must not capture anything: with Nose-Hoover chains + fixcom, which
    is nonlinear in p, step(0) followed by step(n) has to follow the oracle like step(n) alone."""
    cfg = SimConfig(nbeads=6, natoms=10, ndim=3, bosonic=True, fixcom=True, pbc=False, temperature=5.802 * wl.KELVIN,
                    mass=1.0, size=300.0, interaction="free", external="harmonic", ext_omega=3 * wl.MEV,
                    thermostat="nose_hoover", nchains=3, seed=1, dt=wl.FEMTOSECOND)
    x, p = wl.initial_state(cfg, "c1", seed=2)
    p = p + 0.05                                # a centre-of-mass momentum for zeroMomentum to remove
    sim, orc = DeviceSim(cfg), Oracle(cfg)
    sim.upload(x, p)
    orc.set("x", x); orc.set("p", p)
    sim.zero_momentum(); orc.zero_momentum()    # leaves momentum sums behind that a wrongly captured graph would reuse
    sim.step(0)
    sim.step(8)
    for _ in range(8):
        orc.run_iteration()
    assert relerr(sim.get("x"), orc.get("x")) < 1e-10
    assert relerr(sim.get("p"), orc.get("p")) < 1e-9
    sim.close(); orc.close()
