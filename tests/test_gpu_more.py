"""More GPU tests through the C ABI: the reference's golden frames, the documented noise stream, thermostat
statistics, bead sharding (two handles), size-independent properties at the full BASELINE size, error paths."""
import ast
import math

import numpy as np
import pytest

from pimd_b_b200 import workloads as wl
from pimd_b_b200.config import SimConfig, convert_to_internal, parse_ini
from pimd_b_b200.engine import DeviceSim
from tests.helpers import GOLDEN_DIR, Oracle, relerr
from tests.test_host_cpu import philox4x32_10

pytestmark = pytest.mark.gpu

REFCASES = np.load(GOLDEN_DIR / "refcases.npz")
REFPROBE = np.load(GOLDEN_DIR / "refprobe.npz")
DYN_CASES = ["bosonic_quadratic_harmonic_dynamics", "dist_harmonic_dynamics",
             "bosonic_quadratic_harmonic_nmthermostat_dynamics", "dist_harmonic_nm_propagation_dynamics"]
PROBE_CASES = sorted({k.split("/")[0] for k in REFPROBE.files})


# ----------------------------------------------------------------------------- golden vectors of the reference
@pytest.mark.parametrize("case", DYN_CASES)
def test_reference_golden_force_frames(gpu_required, case, tmp_path):
    """positions of a dumped frame -> forces of the same frame (tests/cases/<case>/position_b.xyz, force_b.dat)."""
    p = tmp_path / "case.ini"
    p.write_text(str(REFCASES[f"{case}/ini"]))
    cfg = parse_ini(str(p), ndim=3)
    ang = convert_to_internal("length", "angstrom", 1.0)
    evang = convert_to_internal("force", "ev/ang", 1.0)
    X, F = REFCASES[f"{case}/x"], REFCASES[f"{case}/f"]
    sim = DeviceSim(cfg)
    for frame in range(X.shape[0]):
        sim.set("x", X[frame] * ang)
        sim.update_forces()
        assert relerr(sim.get("f"), F[frame] * evang) < 5e-12, (case, frame)   # 13 printed digits
    sim.close()


@pytest.mark.parametrize("case", PROBE_CASES)
def test_reference_raw_outputs(gpu_required, case):
    """forces / exchange tables / observables / 12-step NVE trajectories of the unmodified reference (raw doubles)."""
    cfg = SimConfig(**ast.literal_eval(str(REFPROBE[f"{case}/cfg"])))
    x, p = REFPROBE[f"{case}/x"], REFPROBE[f"{case}/p"]
    sim = DeviceSim(cfg)
    sim.set("x", x)
    sim.set("p", p)
    sim.update_forces()
    assert relerr(sim.get("f"), REFPROBE[f"{case}/f"]) < 1e-10
    assert relerr(sim.get("f_spring"), REFPROBE[f"{case}/f_spring"]) < 1e-10
    assert relerr(sim.get("f_phys"), REFPROBE[f"{case}/f_phys"]) < 1e-10
    if cfg.bosonic:
        assert relerr(sim.exchange("V"), REFPROBE[f"{case}/exch_V"]) < 1e-10
        assert relerr(sim.exchange("Vb"), REFPROBE[f"{case}/exch_Vb"]) < 1e-10
        assert relerr(sim.exchange("E"), REFPROBE[f"{case}/exch_E"]) < 1e-10
        assert np.max(np.abs(sim.exchange("prob") - REFPROBE[f"{case}/exch_prob"])) < 1e-10
    obs = sim.observables()
    kelvin = convert_to_internal("temperature", "kelvin", 1.0)
    ref = dict(zip(map(str, REFPROBE[f"{case}/obs_names"]), REFPROBE[f"{case}/obs_values"]))
    scale = max(abs(v) for k, v in ref.items() if k not in ("temperature", "prob_dist", "prob_all", "w_gsf"))
    for name, val in ref.items():
        mine = obs[name] / kelvin if name == "temperature" else obs[name]
        tol = 1e-10 * (abs(val) if name in ("temperature", "cl_kinetic", "w_gsf") else max(scale / cfg.nbeads, abs(val)))
        if name in ("prob_dist", "prob_all"):
            tol = 1e-9 * abs(val) + 1e-300
        assert abs(mine - val) <= tol, (name, mine, val)
    if cfg.thermostat == "langevin":   # the reference's own noise stream (one RANMAR generator per bead): its trajectory
        import dataclasses
        sim2 = DeviceSim(dataclasses.replace(cfg, rng="ranmars"))
        sim2.set("x", x)
        sim2.set("p", p)
        sim2.step(12)
        for w in ("x", "p", "f"):
            assert relerr(sim2.get(w), REFPROBE[f"{case}/traj12_{w}"]) < 1e-9, (case, w)
        sim2.close()
    if cfg.thermostat == "none":   # deterministic: trajectory-level parity with the reference itself
        sim2 = DeviceSim(cfg)
        sim2.set("x", x)
        sim2.set("p", p)
        sim2.step(12)
        for w in ("x", "p", "f"):
            assert relerr(sim2.get(w), REFPROBE[f"{case}/traj12_{w}"]) < 1e-10, (case, w)
        sim2.close()
    sim.close()


# ----------------------------------------------------------------------------- the documented noise stream
def ref_gaussian(q, row, draw, seed):
    r = philox4x32_10((q, row, draw & 0xFFFFFFFF, draw >> 32), (seed & 0xFFFFFFFF, seed >> 32))
    a = (r[1] << 32) | r[0]
    b = (r[3] << 32) | r[2]
    u1 = ((a >> 11) + 1) * 2.0 ** -53
    u2 = (b >> 11) * 2.0 ** -53
    rad = math.sqrt(-2.0 * math.log(u1))
    return rad * math.cos(2 * math.pi * u2), rad * math.sin(2 * math.pi * u2)


def expected_noise(cfg, draw, bead_offset=0):
    out = np.empty((cfg.nbeads, cfg.natoms, cfg.ndim))
    for b in range(cfg.nbeads):
        for a in range(cfg.ndim):
            for n in range(cfg.natoms):
                z = ref_gaussian(n >> 1, (b + bead_offset) * cfg.ndim + a, draw, cfg.seed)
                out[b, n, a] = z[n & 1]
    return out


def test_langevin_noise_stream_is_the_documented_philox_stream(gpu_required):
    """xi(bead b, axis a, particle n, half-step k) = BoxMuller(Philox4x32-10(ctr=(n>>1, b*D+a, k, 0), key=seed))[n&1]."""
    cfg = SimConfig(nbeads=3, natoms=7, ndim=3, bosonic=False, fixcom=False, thermostat="langevin", seed=987654321012,
                    temperature=1e-5, mass=3.0, dt=40.0, external="harmonic", ext_omega=1e-4)
    sim = DeviceSim(cfg)
    zero = np.zeros((3, 7, 3))
    sim.set("x", zero)
    sim.set("p", zero)
    c1 = math.exp(-0.5 * cfg.gamma * cfg.dt)
    c2 = math.sqrt((1 - c1 * c1) * cfg.mass / cfg.thermo_beta)
    sim.thermostat_step()
    p0 = sim.get("p")
    assert np.max(np.abs(p0 / c2 - expected_noise(cfg, 0))) < 1e-12
    sim.thermostat_step()
    p1 = sim.get("p")
    assert np.max(np.abs((p1 - c1 * p0) / c2 - expected_noise(cfg, 1))) < 1e-12
    # the stream does not depend on how beads are sharded: a handle owning beads [1,3) draws the same numbers
    part = DeviceSim(cfg, 1, 3)
    part.set("x", zero[1:])
    part.set("p", zero[1:])
    part.step_phase(0)
    assert np.array_equal(part.get("p"), p0[1:])
    sim.close()
    part.close()


def test_langevin_noise_statistics_and_equilibrium(gpu_required):
    cfg = SimConfig(nbeads=16, natoms=512, ndim=3, bosonic=False, fixcom=False, thermostat="langevin", seed=42,
                    temperature=2 * wl.KELVIN, mass=4.0026 * wl.DALTON, dt=wl.FEMTOSECOND, interaction="free",
                    external="free", gamma=1.0 / wl.FEMTOSECOND)   # strong friction: c1 = e^-0.5
    sim = DeviceSim(cfg)
    z = np.zeros((16, 512, 3))
    sim.set("x", z)
    sim.set("p", z)
    c1 = math.exp(-0.5 * cfg.gamma * cfg.dt)
    c2 = math.sqrt((1 - c1 * c1) * cfg.mass / cfg.thermo_beta)
    sim.thermostat_step()
    xi = sim.get("p") / c2
    n = xi.size
    assert abs(xi.mean()) < 5 / math.sqrt(n)
    assert abs(xi.var() - 1.0) < 5 * math.sqrt(2.0 / n)
    assert abs((xi ** 4).mean() - 3.0) < 0.1                       # Gaussian kurtosis
    assert abs(np.corrcoef(xi[:, 0::2, :].ravel(), xi[:, 1::2, :].ravel())[0, 1]) < 5 / math.sqrt(n / 2)
    for _ in range(60):                                             # O steps alone equilibrate <p^2> = m / beta_P
        sim.thermostat_step()
    p = sim.get("p")
    assert abs((p ** 2).mean() * cfg.thermo_beta / cfg.mass - 1.0) < 5 * math.sqrt(2.0 / n)
    sim.close()


def test_nm_thermostat_draws_noise_per_mode(gpu_required):
    cfg = SimConfig(nbeads=6, natoms=5, ndim=2, bosonic=False, fixcom=False, thermostat="langevin", nmthermostat=True,
                    seed=31337, temperature=1e-5, mass=2.0, dt=30.0, external="harmonic", ext_omega=1e-4)
    sim = DeviceSim(cfg)
    z = np.zeros((6, 5, 2))
    sim.set("x", z)
    sim.set("p", z)
    sim.thermostat_step()
    c1 = math.exp(-0.5 * cfg.gamma * cfg.dt)
    c2 = math.sqrt((1 - c1 * c1) * cfg.mass / cfg.thermo_beta)
    pn = c2 * expected_noise(cfg, 0)                                # momenta in normal-mode space, [mode][atom][axis]
    P = cfg.nbeads
    Cm = np.zeros((P, P))
    for k in range(P):
        for j in range(P):
            if k == 0:
                Cm[k, j] = 1 / math.sqrt(P)
            elif k < P / 2:
                Cm[k, j] = math.sqrt(2 / P) * math.cos(2 * math.pi * k * j / P)
            elif k == P / 2:
                Cm[k, j] = (-1) ** j / math.sqrt(P)
            else:
                Cm[k, j] = -math.sqrt(2 / P) * math.sin(2 * math.pi * k * j / P)
    expect = np.einsum("kj,kna->jna", Cm, pn)                       # back to Cartesian beads
    assert np.max(np.abs(sim.get("p") - expect)) < 1e-12 * np.max(np.abs(expect))
    sim.close()


# ----------------------------------------------------------------------------- bead sharding on one GPU
@pytest.mark.parametrize("fixcom", [False, True])
def test_two_sharded_handles_equal_one_handle(gpu_required, fixcom):
    """Two handles owning beads [0,3) and [3,8) driven through the phases, halos and momentum sums moved by hand,
    against one handle owning all beads (fused, graph-replayed). Langevin noise included: the stream is shard-independent."""
    import torch
    from pimd_b_b200.distributed import _DevicePtrView
    N, P = 40, 8
    cfg = SimConfig(nbeads=P, natoms=N, ndim=3, bosonic=True, fixcom=fixcom, pbc=True, temperature=2 * wl.KELVIN,
                    mass=4.0026 * wl.DALTON, size=wl.helium_box(N), interaction="aziz", cutoff=-1.0 * wl.ANGSTROM,
                    external="free", thermostat="langevin", seed=2024, dt=wl.FEMTOSECOND)
    x, p = wl.initial_state(cfg, "c3", seed=3)
    whole = DeviceSim(cfg)
    whole.set("x", x)
    whole.set("p", p)
    shards = [DeviceSim(cfg, 0, 3), DeviceSim(cfg, 3, 8)]
    dev = torch.device("cuda", 0)

    def view(sim, which):
        ptr, cnt = sim.halo_ptr(which)
        return torch.as_tensor(_DevicePtrView(ptr, cnt), device=dev)

    for s, (lo, hi) in zip(shards, [(0, 3), (3, 8)]):
        s.set("x", x[lo:hi])
        s.set("p", p[lo:hi])
    coms = [torch.as_tensor(_DevicePtrView(s.com_ptr(), 4), device=dev) for s in shards]

    def sync():
        for s in shards:
            s.synchronize()

    def halos():
        sync()
        a, b = shards
        view(a, 3).copy_(view(b, 0)); view(a, 2).copy_(view(b, 1))
        view(b, 3).copy_(view(a, 0)); view(b, 2).copy_(view(a, 1))
        torch.cuda.synchronize()

    def allreduce():
        if not fixcom:
            return
        sync()
        tot = coms[0] + coms[1]
        coms[0].copy_(tot); coms[1].copy_(tot)
        torch.cuda.synchronize()

    halos()
    K = 6
    for _ in range(K):
        for s in shards:
            s.step_phase(0)
        allreduce()
        for s in shards:
            s.step_phase(1)
        halos()
        for s in shards:
            s.step_phase(2)
        allreduce()
        for s in shards:
            s.step_phase(3)
    whole.step(K)
    for w in ("x", "p", "f"):
        got = np.concatenate([shards[0].get(w), shards[1].get(w)])
        ref = whole.get(w)
        if fixcom:
            assert relerr(got, ref) < 1e-12, w        # momentum sums are added in another order
        else:
            assert np.array_equal(got, ref), w
    ow = whole.observables()
    o0, o1 = shards[0].observables(), shards[1].observables()
    for k in ow:
        if np.isnan(ow[k]):   # the GSF columns are undefined with an interaction potential (include/pimdb200.h)
            assert k in ("w_gsf", "pot_gsf") and np.isnan(o0[k]) and np.isnan(o1[k])
            continue
        assert abs(o0[k] + o1[k] - ow[k]) <= 1e-11 * max(abs(ow[k]), abs(ow["cl_spring"])), k
    for s in shards + [whole]:
        s.close()


# ----------------------------------------------------------------------------- full BASELINE size: properties
def test_c3_full_size_properties(gpu_required):
    cfg = wl.config("c3")
    cfg.thermostat = "none"
    x, p = wl.initial_state(cfg, "c3")
    sim = DeviceSim(cfg)
    sim.set("x", x)
    sim.set("p", p)
    sim.update_forces()
    fp = sim.get("f_phys")
    f = sim.get("f")
    big = np.abs(fp).sum(axis=1).max()
    assert np.max(np.abs(fp.sum(axis=1))) < 1e-11 * big               # Newton's third law per bead
    assert np.max(np.abs(f.sum(axis=(0, 1)))) < 1e-10 * np.abs(f).sum()  # ring springs + exchange forces are internal too
    # periodic images: shifting every particle by a lattice vector leaves minimum-image forces unchanged
    shift = np.zeros_like(x)
    shift[:, ::3, 0] = cfg.size
    shift[:, 1::3, 2] = -2 * cfg.size
    sim.set("x", x + shift)
    sim.update_forces()
    assert relerr(sim.get("f"), f) < 1e-9
    # relabelling the particles permutes the pair forces (different tiles, different summation order)
    perm = np.random.default_rng(0).permutation(cfg.natoms)
    cfgd = SimConfig(**{**cfg.as_dict(), "bosonic": False})
    d = DeviceSim(cfgd)
    d.set("x", x)
    d.update_forces()
    f0 = d.get("f")
    d.set("x", x[:, perm])
    d.update_forces()
    assert relerr(d.get("f"), f0[:, perm]) < 1e-11
    d.close()
    # NVE: the ring-polymer Hamiltonian  sum p^2/2m + V_B + springs + sum_beads V  is conserved
    sim.set("x", x)
    sim.set("p", p)
    sim.update_forces()

    def hamiltonian():
        o = sim.observables()
        return o["cl_kinetic"] + o["cl_spring"] + o["potential"] * cfg.nbeads

    h0 = hamiltonian()
    sim.step(100)
    h1 = hamiltonian()
    ke = sim.observables()["cl_kinetic"]
    assert abs(h1 - h0) < 2e-3 * ke, (h0, h1, ke)
    sim.close()


def test_c3_against_oracle_on_a_bead_subset(gpu_required):
    """Full N=512: forces of the interior beads 5..6 need only beads 4..7 -> compare with the oracle run on 4 beads."""
    cfg = wl.config("c3")
    x, p = wl.initial_state(cfg, "c3")
    sim = DeviceSim(cfg)
    sim.set("x", x)
    sim.update_forces()
    f = sim.get("f")
    sub = SimConfig(**{**cfg.as_dict(), "nbeads": 4, "bosonic": False})
    # the spring constant m (P/beta)^2 must stay that of P = 64: scale the temperature by 64/4
    sub.temperature = cfg.temperature * cfg.nbeads / 4
    orc = Oracle(sub)
    orc.set("x", x[4:8])
    orc.update_forces()
    assert relerr(f[5:7], orc.get("f")[1:3]) < 1e-10
    sim.close()


# ----------------------------------------------------------------------------- errors
def test_error_paths(gpu_required):
    with pytest.raises(ValueError, match="Normal modes propogation is currently not available for bosons!"):
        DeviceSim(SimConfig(nbeads=4, natoms=4, bosonic=True, propagator="normal_modes"))
    with pytest.raises(ValueError, match="nmthermostat cannot be used in nve ensemble!"):
        DeviceSim(SimConfig(nbeads=4, natoms=4, nmthermostat=True, thermostat="none"))
    cfg = SimConfig(nbeads=4, natoms=6, ndim=3, bosonic=True, fixcom=False, thermostat="none", temperature=1e-5,
                    mass=1.0, external="harmonic", ext_omega=1e-4)
    sim = DeviceSim(cfg)
    with pytest.raises(ValueError):
        sim.set("x", np.zeros((4, 6, 2)))
    x = np.random.default_rng(1).normal(size=(4, 6, 3))
    x[0, 2, 1] = np.nan
    sim.set("x", x)
    sim.update_forces()
    with pytest.raises(OverflowError, match="bosonic exchange potential"):   # std::overflow_error in the reference
        sim.synchronize()
    part = DeviceSim(cfg, 0, 2)
    with pytest.raises(ValueError, match="needs all beads"):
        part.step(1)
    sim.close()
    part.close()


def test_page_locked_host_buffers_take_the_direct_copy_path(gpu_required):
    """pimdb_set_state / pimdb_get_state copy straight from / into page-locked caller buffers and stage pageable
    ones; both must move the same bytes."""
    import torch
    cfg = wl.config("c1")
    x, p = wl.initial_state(cfg, "c1")
    sim = DeviceSim(cfg)
    pinned = [torch.empty(x.shape, dtype=torch.float64, pin_memory=True) for _ in range(3)]
    hx, hp, hout = (t.numpy() for t in pinned)
    hx[...] = x
    hp[...] = p
    sim.set("x", hx)
    sim.set("p", hp)
    assert np.array_equal(sim.get("x"), x) and np.array_equal(sim.get("p"), p)      # pinned in, pageable out
    sim.set("x", x)
    sim.set("p", p)
    assert np.array_equal(sim.get("x", hout), x)                                      # pageable in, pinned out
    sim.step(3)
    f_pageable = sim.get("f")
    assert np.array_equal(sim.get("f", hout), f_pageable)
    sim.close()


@pytest.mark.parametrize("rng", ["philox", "ranmars"])
def test_fixcom_with_odd_particle_count_in_the_fused_step(gpu_required, rng):
    """Regression: the fused integrator handles particles in pairs; with an odd particle count the empty second slot of
    the last pair must not enter the centre-of-mass sum after the COM shift / Langevin step of the same kernel have
    acted on it. pimdb_step (fused, deferred COM shift) must equal the call-by-call order of Simulation::run
    (src/simulation.cpp:246-259) and leave no centre-of-mass momentum (Simulation::zeroMomentum :581-603)."""
    import dataclasses
    case = "aziz_pbc_bosonic"          # N = 27, fixcom, Langevin
    cfg = dataclasses.replace(SimConfig(**ast.literal_eval(str(REFPROBE[f"{case}/cfg"]))), rng=rng)
    assert cfg.natoms % 2 == 1 and cfg.fixcom and cfg.thermostat == "langevin"
    x, p = REFPROBE[f"{case}/x"], REFPROBE[f"{case}/p"]
    a = DeviceSim(cfg); a.set("x", x); a.set("p", p)
    a.step(7)
    b = DeviceSim(cfg); b.set("x", x); b.set("p", p)
    for _ in range(7):
        b.thermostat_step(); b.zero_momentum(); b.propagator_step(); b.thermostat_step(); b.zero_momentum()
    pa = a.get("p")
    assert np.max(np.abs(pa.sum(axis=(0, 1)))) < 1e-12 * np.abs(pa).sum()
    for w in ("x", "p", "f"):
        assert relerr(a.get(w), b.get(w)) < 1e-12, w
    a.close(); b.close()


# ----------------------------------------------------------------------------- factorial exchange (the reference's other class)
REFFACT = np.load(GOLDEN_DIR / "reffactorial.npz")
FACT_PROBES = sorted({k.split("/")[0] for k in REFFACT.files if k.startswith("factorial_")})


@pytest.mark.parametrize("case", FACT_PROBES)
def test_factorial_exchange_matches_reference_factorial_build(gpu_required, case):
    """exchange_alg = "factorial" (csrc/factorial.cu) against raw outputs of the unmodified reference built with
    -DFACTORIAL_BOSONIC_ALGORITHM (tests/golden/reffactorial.npz): forces, estimators, NVE trajectory."""
    cfg = SimConfig(**ast.literal_eval(str(REFFACT[f"{case}/cfg"])))
    x, p = REFFACT[f"{case}/x"], REFFACT[f"{case}/p"]
    sim = DeviceSim(cfg)
    sim.upload(x, p)
    sim.update_forces()
    for w, k in (("f", "f"), ("f_spring", "f_spring"), ("f_phys", "f_phys")):
        assert relerr(sim.get(w), REFFACT[f"{case}/{k}"]) < 1e-10, k
    obs = sim.observables()
    kelvin = wl.KELVIN
    for name, val in zip(REFFACT[f"{case}/obs_names"], REFFACT[f"{case}/obs_values"]):
        mine = obs[str(name)] / kelvin if name == "temperature" else obs[str(name)]
        assert abs(mine - val) <= 1e-10 * max(abs(val), abs(obs["cl_spring"]), 1e-300), (name, mine, val)
    if cfg.thermostat == "none":
        fresh = DeviceSim(cfg)
        fresh.upload(x, p)
        fresh.step(12)
        for w in ("x", "p", "f"):
            assert relerr(fresh.get(w), REFFACT[f"{case}/traj12_{w}"]) < 1e-9, w
        fresh.close()
    else:   # the reference's own noise: the Langevin trajectory itself
        import dataclasses
        fresh = DeviceSim(dataclasses.replace(cfg, rng="ranmars"))
        fresh.upload(x, p)
        fresh.step(12)
        for w in ("x", "p", "f"):
            assert relerr(fresh.get(w), REFFACT[f"{case}/traj12_{w}"]) < 1e-9, w
        fresh.close()
    sim.close()


def test_factorial_exchange_matches_oracle_on_random_inputs_and_rejects_large_n(gpu_required):
    from tests.helpers import Oracle
    rng = np.random.default_rng(42)
    for N, P, pbc in ((1, 3, False), (2, 2, False), (4, 5, True), (8, 4, False)):
        cfg = SimConfig(nbeads=P, natoms=N, ndim=3, bosonic=True, fixcom=False, pbc=pbc, temperature=5.802 * wl.KELVIN, mass=1.0,
                        size=60.0 if pbc else 300.0, interaction="free", external="harmonic", ext_omega=3 * wl.MEV,
                        thermostat="none", seed=1, dt=wl.FEMTOSECOND, exchange_alg="factorial")
        x = rng.uniform(-25, 25, size=(P, N, 3))
        sim, orc = DeviceSim(cfg), Oracle(cfg)
        sim.upload(x, np.zeros_like(x))
        orc.set("x", x)
        sim.update_forces(); orc.update_forces()
        assert relerr(sim.get("f"), orc.get("f")) < 1e-10, N
        o, r = sim.observables(), orc.observables()
        for k in ("kinetic", "cl_spring"):
            assert abs(o[k] - r[k]) <= 1e-10 * abs(r[k]), (N, k)
        sim.close(); orc.close()
    with pytest.raises(ValueError, match="natoms <= 10"):
        DeviceSim(SimConfig(nbeads=4, natoms=11, bosonic=True, exchange_alg="factorial"))


def test_step_download_equals_step_then_download(gpu_required):
    """pimdb_step_download (coordinates copied to the host while the forces of the last iteration are still being computed)
    returns exactly what pimdb_step + pimdb_download_state return, for page-locked and pageable destinations, and leaves the
    handle in the same state."""
    import torch
    cfg = SimConfig(nbeads=8, natoms=64, ndim=3, bosonic=True, fixcom=True, pbc=True, temperature=2 * wl.KELVIN,
                    mass=4.0026 * wl.DALTON, size=wl.helium_box(64), interaction="aziz", cutoff=-1.0 * wl.ANGSTROM,
                    external="free", thermostat="langevin", seed=12345, dt=wl.FEMTOSECOND)
    x, p = wl.initial_state(cfg, "c3", seed=1)
    a, b = DeviceSim(cfg), DeviceSim(cfg)
    a.upload(x, p); b.upload(x, p)
    pin = torch.empty((3,) + x.shape, dtype=torch.float64, pin_memory=True)
    hx, hp, hf = (pin[i].numpy() for i in range(3))
    for nsteps in (1, 3, 1):
        a.step(nsteps)
        ref = (a.get("x"), a.get("p"), a.get("f"))
        b.step_download(nsteps, hx, hp, hf)
        for got, want in zip((hx, hp, hf), ref):
            assert np.array_equal(got, want)
    px, pp, pf = np.empty_like(x), np.empty_like(x), np.empty_like(x)      # pageable: the plain sequence
    a.step(2); b.step_download(2, px, pp, pf)
    assert np.array_equal(px, a.get("x")) and np.array_equal(pp, a.get("p")) and np.array_equal(pf, a.get("f"))
    b.step_download(1, hx, None, None)
    a.step(1)
    assert np.array_equal(hx, a.get("x")) and np.array_equal(b.get("p"), a.get("p"))
    a.close(); b.close()


@pytest.mark.parametrize("natoms", [64, 33])
def test_prefetched_langevin_noise_is_the_in_place_noise(gpu_required, natoms, monkeypatch):
    """The Gaussians of the counter-based Langevin stream are drawn ahead, beside the force kernels (k_noise_prefetch), and the
    thermostat half steps load them; a handle created with PIMDB_NO_NOISE_PREFETCH=1 draws them in place. Same numbers, bit
    for bit, through pimdb_step (captured and repeated), through the call-by-call entry points in between, and after the
    momenta were replaced from the host (even and odd particle counts: vector and scalar kernels)."""
    cfg = SimConfig(nbeads=8, natoms=natoms, ndim=3, bosonic=True, fixcom=True, pbc=True, temperature=2 * wl.KELVIN,
                    mass=4.0026 * wl.DALTON, size=wl.helium_box(64), interaction="aziz", cutoff=-1.0 * wl.ANGSTROM,
                    external="free", thermostat="langevin", seed=777, dt=wl.FEMTOSECOND)
    import dataclasses
    x, p = wl.initial_state(dataclasses.replace(cfg, natoms=64), "c3", seed=3)
    x, p = np.ascontiguousarray(x[:, :natoms]), np.ascontiguousarray(p[:, :natoms])
    a = DeviceSim(cfg)
    monkeypatch.setenv("PIMDB_NO_NOISE_PREFETCH", "1")
    b = DeviceSim(cfg)
    monkeypatch.delenv("PIMDB_NO_NOISE_PREFETCH")
    # (the third handle also takes the general launch sequence, where the last block of every thermostat launch advances the
    # draw counter; a and b let the kick-and-drift launch between the two thermostat launches advance it)
    monkeypatch.setenv("PIMDB_NO_TICKETLESS", "1")
    c = DeviceSim(cfg)
    for s in (a, b, c):
        s.upload(x, p)
        s.step(3)
        s.thermostat_step(); s.zero_momentum()      # a half step outside pimdb_step consumes the slot drawn for "the next opening"
        s.step(2)
        s.set("p", p)
        s.step(1); s.step(4)
    for k in ("x", "p", "f"):
        assert np.array_equal(a.get(k), b.get(k)), k
        assert np.array_equal(a.get(k), c.get(k)), k
    a.close(); b.close(); c.close()
