"""Exchange-chain timing versus N (warm, CUDA events): separates the per-step chain latency from consumer throughput."""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from pimd_b_b200 import workloads as wl  # noqa: E402
from pimd_b_b200.config import SimConfig  # noqa: E402
from pimd_b_b200.engine import DeviceSim  # noqa: E402

for n in (32, 64, 128, 256, 512, 1024):
    cfg = SimConfig(nbeads=4, natoms=n, ndim=3, bosonic=True, fixcom=False, pbc=False, temperature=5 * wl.KELVIN,
                    mass=4.0026 * wl.DALTON, size=100 * wl.ANGSTROM, interaction="free", external="harmonic",
                    ext_omega=3 * wl.MEV, thermostat="none")
    rng = np.random.default_rng(n)
    x = np.repeat(rng.normal(0, 20.0, size=(1, n, 3)), 4, axis=0) + rng.normal(0, 1.0, size=(4, n, 3))
    sim = DeviceSim(cfg)
    sim.set("x", x)
    arr = (C.c_double * 2)()
    sim.lib.pimdb_debug_exchange_timing.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    sim.lib.pimdb_debug_exchange_timing(sim.h, 50, arr)
    print(json.dumps({"N": n, "prefix_factors_us": round(arr[0], 2), "recur_forces_us": round(arr[1], 2),
                      "ns_per_step": round(arr[1] * 1e3 / n, 1)}))
    sim.close()
