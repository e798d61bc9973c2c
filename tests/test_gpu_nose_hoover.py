"""Nose-Hoover chain thermostats on the GPU against the oracle, which is bit-identical to the
reference's three variants (src/thermostats/nose_hoover.cpp): chain state is carried across steps, so the
`nh_energy` column (Thermostat::getAdditionToH summed over beads) and the conserved quantity are checked too."""
import numpy as np
import pytest

from pimd_b_b200.config import SimConfig
from pimd_b_b200.engine import DeviceSim
from tests.helpers import FEMTOSECOND, KELVIN, MEV, Oracle, maxwell_momenta, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("thermostat", ["nose_hoover", "nose_hoover_np", "nose_hoover_np_dim"])
@pytest.mark.parametrize("bosonic,nmcoupled", [(True, False), (False, False), (True, True), (False, True)])
def test_nose_hoover_trajectory_and_conserved_quantity(gpu_required, thermostat, bosonic, nmcoupled):
    """Cartesian coupling and coupling to the normal-mode momenta (nmthermostat = true; the oracle's NM-coupled chains
    are bit-identical to the reference on tests/golden/refprobe.npz *_nmcoupled)."""
    cfg = SimConfig(nbeads=4, natoms=8, ndim=3, bosonic=bosonic, fixcom=False, pbc=False,
                    temperature=5.802 * KELVIN, mass=1.0, size=300.0, interaction="harmonic", int_omega=1 * MEV,
                    external="harmonic", ext_omega=3 * MEV, thermostat=thermostat, nmthermostat=nmcoupled, nchains=4,
                    seed=90846, dt=0.1 * FEMTOSECOND)
    rng = np.random.default_rng(8)
    x = rng.uniform(-30, 30, size=(4, 8, 3))
    p = maxwell_momenta(cfg, rng)
    K = 60
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.set("p", p)
    sim = DeviceSim(cfg)
    sim.set("x", x)
    sim.set("p", p)

    def conserved(o):   # ring-polymer Hamiltonian + thermostat contribution
        return o["cl_kinetic"] + o["cl_spring"] + o["potential"] * cfg.nbeads + o["nh_energy"]

    sim.step(1)
    orc.run_iteration()
    h0 = conserved(sim.observables())
    for _ in range(K):
        orc.run_iteration()
    sim.step(K)
    for w in ("x", "p", "f"):
        assert relerr(sim.get(w), orc.get(w)) < 1e-9, w
    got, ref = sim.observables(), orc.observables()
    assert abs(got["nh_energy"] - ref["nh_energy"]) <= 1e-9 * max(abs(ref["nh_energy"]), abs(ref["cl_kinetic"]))
    assert abs(got["nh_energy"]) > 0.0
    assert abs(conserved(got) - h0) < 1e-5 * abs(got["cl_kinetic"])
    sim.close()


@pytest.mark.parametrize("case", ["nose_hoover_nmcoupled", "nose_hoover_np_nmcoupled", "nose_hoover_np_dim_nmcoupled"])
def test_nm_coupled_chains_match_the_reference_trajectory(gpu_required, case):
    """12 iterations from the reference's own raw state dump (ref_probe, tests/golden/refprobe.npz)."""
    import ast
    from tests.helpers import GOLDEN_DIR
    ref = np.load(GOLDEN_DIR / "refprobe.npz")
    cfg = SimConfig(**ast.literal_eval(str(ref[f"{case}/cfg"])))
    sim = DeviceSim(cfg)
    sim.set("x", ref[f"{case}/x"])
    sim.set("p", ref[f"{case}/p"])
    sim.step(12)
    for w in ("x", "p", "f"):
        assert relerr(sim.get(w), ref[f"{case}/traj12_{w}"]) < 1e-9, (case, w)
    sim.close()
