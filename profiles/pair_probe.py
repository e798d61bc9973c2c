"""Pair-tile kernel A/B probe: (1) the pair tiles + assembly alone (distinguishable copy of the workload, 300 eager
update_forces back to back, wall clock / call), (2) the whole captured step back to back, (3) the in-kernel timeline of
one captured step (longest slot = pair tiles).   python profiles/pair_probe.py c3|c4"""
import ctypes as C, dataclasses, json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
os.environ["PIMDB_TIMELINE"] = "1"
from pimd_b_b200 import workloads as wl
from pimd_b_b200.engine import DeviceSim
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = wl.config(name)
x, p = wl.initial_state(cfg, name)
out = {"workload": name, "lib": os.environ.get("PIMDB200_LIB", "main").split("/")[-1], "dual": os.environ.get("PIMDB_PAIR_DUAL"),
       "split": os.environ.get("PIMDB_PAIR_SPLIT")}
d = DeviceSim(dataclasses.replace(cfg, bosonic=False, obs_bosonic="false")); d.set("x", x); d.set("p", p)
for _ in range(20): d.update_forces()
d.synchronize()
n = 300 if cfg.natoms <= 512 else 40
t0 = time.perf_counter()
for _ in range(n): d.update_forces()
d.synchronize()
out["pair_plus_assemble_us"] = (time.perf_counter() - t0) / n * 1e6
d.close()
sim = DeviceSim(cfg); sim.set("x", x); sim.set("p", p)
nst = 1000 if cfg.natoms <= 512 else 60
sim.step(20); sim.synchronize()
t0 = time.perf_counter(); sim.step(nst); sim.synchronize()
out["graph_step_us"] = (time.perf_counter() - t0) / nst * 1e6
fn = sim.lib.pimdb_debug_timeline; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
buf = (C.c_ulonglong * 64)()
fn(sim.h, buf)
best = []
for rep in range(5):
    sim.step(1); sim.synchronize()
    k = fn(sim.h, buf)
    t = np.array(buf[:], dtype=np.uint64).reshape(32, 2)[:k].astype(np.int64)
    dur = (t[:, 1] - t[:, 0]) / 1e3
    best.append((float(dur.max()), float((t[:, 1].max() - t[:, 0].min()) / 1e3)))
out["timeline_longest_kernel_us"] = min(b[0] for b in best)
out["timeline_step_span_us"] = min(b[1] for b in best)
print(json.dumps(out))
