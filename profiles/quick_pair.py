"""Quick timing of the pair-force kernel and the whole step on C3 (eager pass with CUDA events per launch)."""
import sys, json
sys.path.insert(0, ".")
from pimd_b_b200 import workloads as wl
from pimd_b_b200.engine import DeviceSim
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = wl.config(name)
x, p = wl.initial_state(cfg, name)
sim = DeviceSim(cfg)
sim.set("x", x); sim.set("p", p)
sim.step(20); sim.synchronize()
sim.timing_enable(True)
sim.step(100)
pair_ms, n = sim.timing_get(0)
step_ms, _ = sim.timing_get(1)
sim.timing_enable(False)
import time
sim.step(50); sim.synchronize()
t0 = time.perf_counter(); sim.step(2000); sim.synchronize(); t1 = time.perf_counter()
print(json.dumps({"workload": name, "pair_us": pair_ms * 1e3, "pair_tflops_alg": wl.pair_flops_per_step(cfg) / (pair_ms * 1e-3) * 1e-12 if pair_ms else None,
                  "eager_step_us": step_ms * 1e3, "graph_step_us": (t1 - t0) / 2000 * 1e6, "steps_per_s": 2000 / (t1 - t0)}))
