"""Quick timing of the pair-force kernel and the whole step on a BASELINE workload (eager pass with CUDA events per
launch, then graph replay).   python profiles/quick_pair.py c3 [nsteps]"""
import json
import sys
import time

sys.path.insert(0, ".")
from pimd_b_b200 import workloads as wl  # noqa: E402
from pimd_b_b200.engine import DeviceSim  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
cfg = wl.config(name)
x, p = wl.initial_state(cfg, name)
sim = DeviceSim(cfg)
sim.set("x", x)
sim.set("p", p)
sim.step(5)
sim.synchronize()
sim.timing_enable(True)
sim.step(max(5, nsteps // 20))
pair_ms, n = sim.timing_get(0)
step_ms, _ = sim.timing_get(1)
sim.timing_enable(False)
sim.step(5)
sim.synchronize()
t0 = time.perf_counter()
sim.step(nsteps)
sim.synchronize()
t1 = time.perf_counter()
exch_us = None
if cfg.bosonic:   # the exchange chain alone (prefix + factors + recurrences + exterior forces), warm caches
    sim.exchange_prepare()
    sim.synchronize()
    t2 = time.perf_counter()
    for _ in range(200):
        sim.exchange_prepare()
    sim.synchronize()
    exch_us = (time.perf_counter() - t2) / 200 * 1e6
parts = None
if cfg.bosonic:
    import ctypes as C
    arr = (C.c_double * 2)()
    sim.lib.pimdb_debug_exchange_timing.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    sim.lib.pimdb_debug_exchange_timing(sim.h, 100, arr)
    parts = {"prefix_plus_factors_us": arr[0], "recurrences_plus_forces_us": arr[1]}
print(json.dumps({"exchange_chain_us": exch_us, "parts": parts}))
obs = sim.observables()
print(json.dumps({"workload": name, "natoms": cfg.natoms, "nbeads": cfg.nbeads, "pair_us": pair_ms * 1e3,
                  "pair_tflops_alg": wl.pair_flops_per_step(cfg) / (pair_ms * 1e-3) * 1e-12 if pair_ms else None,
                  "eager_step_us": step_ms * 1e3, "graph_step_us": (t1 - t0) / nsteps * 1e6,
                  "steps_per_s": nsteps / (t1 - t0), "kinetic": obs["kinetic"], "potential": obs["potential"],
                  "temperature_K": obs["temperature"] / wl.KELVIN}))
