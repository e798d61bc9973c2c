"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): bosonic He-4 Aziz, N = 70 (3 row blocks,
ragged), P = 4, a few fused steps + observables; and a stiff-spring case that takes the exact-fallback path.
    compute-sanitizer --tool memcheck python profiles/sanitize_small.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from pimd_b_b200 import workloads as wl
from pimd_b_b200.config import SimConfig
from pimd_b_b200.engine import DeviceSim

cfg = SimConfig(nbeads=4, natoms=70, ndim=3, bosonic=True, fixcom=True, pbc=True, temperature=2 * wl.KELVIN,
                mass=4.0026 * wl.DALTON, size=wl.helium_box(70), interaction="aziz", cutoff=-1.0 * wl.ANGSTROM,
                external="free", thermostat="langevin", seed=7, dt=wl.FEMTOSECOND)
x, p = wl.initial_state(cfg, "c3", seed=1)
sim = DeviceSim(cfg); sim.set("x", x); sim.set("p", p)
sim.update_forces(); sim.step(3); print(sim.observables()["kinetic"]); sim.close()
cfg2 = SimConfig(nbeads=3, natoms=80, ndim=3, bosonic=True, fixcom=False, pbc=False, temperature=2 * wl.KELVIN,
                 mass=4.0026 * wl.DALTON, size=40.0, interaction="free", external="harmonic", ext_omega=3 * wl.MEV,
                 thermostat="langevin", seed=7, dt=wl.FEMTOSECOND, rng="ranmars")
rng = np.random.default_rng(3)
x2 = rng.uniform(-20, 20, size=(3, 80, 3)); p2 = rng.normal(0, 1, size=(3, 80, 3))
sim = DeviceSim(cfg2); sim.set("x", x2); sim.set("p", p2)
sim.update_forces(); sim.step(2); print(sim.exchange("V")[-1]); sim.close()
