"""Diagnostics: error growth of the C2 trajectory (normal-mode propagator + normal-mode Langevin, reference noise)."""
import dataclasses, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from pimd_b_b200 import workloads as wl
from pimd_b_b200.engine import DeviceSim
from tests.helpers import Oracle, relerr

def mind(x):
    d = x[:, :, None, :] - x[:, None, :, :]
    r = np.sqrt((d ** 2).sum(-1)) + 1e9 * np.eye(x.shape[1])[None]
    return r.min()

for label, kw in (("nm+nmthermo ranmars", dict(rng="ranmars")),
                  ("nm propagator, cartesian langevin ranmars", dict(rng="ranmars", nmthermostat=False)),
                  ("nve", dict(thermostat="none", nmthermostat=False)),
                  ("cartesian propagator + cartesian langevin ranmars", dict(rng="ranmars", nmthermostat=False, propagator="cartesian"))):
    cfg = dataclasses.replace(wl.config("c2"), **kw)
    x, p = wl.initial_state(cfg, "c2")
    sim, orc = DeviceSim(cfg), Oracle(cfg)
    sim.upload(x, p); orc.set("x", x); orc.set("p", p)
    print("==", label, "gamma", cfg.gamma, "dt", cfg.dt, "min dist", mind(x), flush=True)
    for it in range(14):
        sim.step(1); orc.run_iteration()
        xs, xo = sim.get("x"), orc.get("x")
        print(f"  step {it+1}: x {relerr(xs, xo):.1e} p {relerr(sim.get('p'), orc.get('p')):.1e} f {relerr(sim.get('f'), orc.get('f')):.1e}"
              f" max|f| {np.abs(orc.get('f')).max():.2e} max|p| {np.abs(orc.get('p')).max():.2e} min dist {mind(xo):.3f}", flush=True)
    # re-sync the GPU to the oracle state and compare a single force evaluation there
    sim.upload(orc.get("x"), orc.get("p")); sim.update_forces()
    fo = orc.get("f").copy()
    orc2 = Oracle(cfg); orc2.set("x", orc.get("x")); orc2.update_forces()
    print("  forces on the oracle's final positions:", relerr(sim.get("f"), orc2.get("f")))
    sim.close(); orc.close(); orc2.close()
