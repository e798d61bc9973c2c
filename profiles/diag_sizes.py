"""Diagnostics behind tests/test_gpu_baseline_sizes.py: where the GPU path and the oracle differ at full size."""
import dataclasses, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from pimd_b_b200 import workloads as wl
from pimd_b_b200.config import SimConfig
from pimd_b_b200.engine import DeviceSim
from tests.helpers import Oracle, relerr
from tests.test_gpu_baseline_sizes import shrink_beads, per_component_ok

def where(a, b, tag):
    d = np.abs(a - b); i = np.unravel_index(np.argmax(d), d.shape)
    print(f"  {tag}: relerr {relerr(a,b):.3e} worst at {i}: got {a[i]:.15e} ref {b[i]:.15e}; per-bead relerr",
          " ".join(f"{relerr(a[k], b[k]):.1e}" for k in range(min(a.shape[0], 8))), flush=True)

which = sys.argv[1:] or ["c3", "c2", "c4", "c5"]
if "c3" in which:
    print("== c3")
    cfg = dataclasses.replace(wl.config("c3"), thermostat="none")
    x, p = wl.initial_state(cfg, "c3")
    sim, orc = DeviceSim(cfg), Oracle(cfg)
    sim.upload(x, p); orc.set("x", x); orc.set("p", p)
    sim.update_forces(); orc.update_forces()
    f, fr = sim.get("f"), orc.get("f")
    where(f, fr, "f")
    for beads in (slice(0, 1), slice(63, 64), slice(1, 63)):
        print("  per-component", beads, per_component_ok(f[beads], fr[beads]))
    where(sim.get("f_spring"), orc.get("s"), "f_spring")
    where(sim.get("f_phys"), orc.get("e"), "f_phys")
    print("  V", relerr(sim.exchange("V"), orc.exchange("V")), "Vb", relerr(sim.exchange("Vb"), orc.exchange("B")))
    sim.step(5)
    for _ in range(5): orc.run_iteration()
    where(sim.get("x"), orc.get("x"), "x5"); where(sim.get("p"), orc.get("p"), "p5"); where(sim.get("f"), orc.get("f"), "f5")
    print("  per-component f5", per_component_ok(sim.get("f"), orc.get("f"), rel=1e-9, floor=1e-10))
    o, r = sim.observables(), orc.observables()
    for k in o: print(f"  obs {k}: {o[k]:.15e} {r.get(k, float('nan')):.15e}")
    sim.close(); orc.close()
if "c2" in which:
    print("== c2")
    for P, N in ((8, 12), (16, 64), (32, 64), (64, 64), (64, 16)):
        for rng in ("ranmars",):
            cfg = dataclasses.replace(wl.config("c2"), rng=rng, nbeads=P, natoms=N)
            x, p = wl.initial_state(cfg, "c2")
            sim, orc = DeviceSim(cfg), Oracle(cfg)
            sim.upload(x, p); orc.set("x", x); orc.set("p", p)
            errs = []
            for it in range(3):
                sim.step(1); orc.run_iteration()
                errs.append((relerr(sim.get("x"), orc.get("x")), relerr(sim.get("p"), orc.get("p"))))
            print(f"  P={P} N={N} {rng}: (x,p) relerr per step", " ".join(f"({a:.1e},{b:.1e})" for a, b in errs), flush=True)
            # thermostat step alone
            sim.upload(x, p); orc.set("x", x); orc.set("p", p)
            sim.thermostat_step(); orc.thermostat_step()
            print(f"     thermostat step alone: p relerr {relerr(sim.get('p'), orc.get('p')):.2e}")
            sim.propagator_step(); orc.propagator_step()
            print(f"     + propagator step: x {relerr(sim.get('x'), orc.get('x')):.2e} p {relerr(sim.get('p'), orc.get('p')):.2e}")
            sim.close(); orc.close()
if "c4" in which:
    print("== c4")
    full = wl.config("c4")
    cfg = dataclasses.replace(shrink_beads(full, 4), thermostat="none")
    x, p = wl.initial_state(full, "c4"); x, p = np.ascontiguousarray(x[:4]), np.ascontiguousarray(p[:4])
    sim, orc = DeviceSim(cfg), Oracle(cfg)
    sim.upload(x, p); orc.set("x", x); orc.set("p", p)
    sim.update_forces(); orc.update_forces()
    where(sim.get("f"), orc.get("f"), "f"); where(sim.get("f_spring"), orc.get("s"), "f_spring"); where(sim.get("f_phys"), orc.get("e"), "f_phys")
    print("  V", relerr(sim.exchange("V"), orc.exchange("V")), "Vb", relerr(sim.exchange("Vb"), orc.exchange("B")),
          "E", relerr(sim.exchange("E"), orc.exchange("E")))
    pr, pro = sim.exchange("prob").reshape(2048, 2048), orc.exchange("P").reshape(2048, 2048)
    print("  prob max abs diff", np.max(np.abs(pr - pro)), "rows sum-1:", np.max(np.abs(pr.sum(1) - 1)), "oracle rows:", np.max(np.abs(pro.sum(1) - 1)))
    sim.close(); orc.close()
if "c5" in which:
    print("== c5")
    full = wl.config("c5")
    cfg = dataclasses.replace(shrink_beads(full, 2), thermostat="none")
    x, p = wl.initial_state(full, "c5"); x, p = np.ascontiguousarray(x[:2]), np.ascontiguousarray(p[:2])
    sim, orc = DeviceSim(cfg), Oracle(cfg)
    sim.upload(x, p); orc.set("x", x); orc.set("p", p)
    sim.update_forces(); orc.update_forces()
    where(sim.get("f"), orc.get("f"), "f")
    V, Vo = sim.exchange("V"), orc.exchange("V")
    print("  V", relerr(V, Vo), "Vb", relerr(sim.exchange("Vb"), orc.exchange("B")), "beta*max|V|", np.max(np.abs(Vo)) / cfg.temperature / 1 * 1.0 / cfg.nbeads)
    pr = sim.exchange("prob").reshape(8192, 8192); pro = orc.exchange("P").reshape(8192, 8192)
    print("  prob max abs diff", np.max(np.abs(pr - pro)), "gpu rows sum-1:", np.max(np.abs(pr.sum(1) - 1)), "oracle rows sum-1:", np.max(np.abs(pro.sum(1) - 1)))
    o, r = sim.observables(), orc.observables()
    for k in ("kinetic", "cl_spring", "prob_dist", "prob_all"): print(f"  obs {k}: {o[k]:.15e} {r[k]:.15e}")
    sim.close(); orc.close()
