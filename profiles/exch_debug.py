"""Per-warp clock64 breakdown of the decoupled exchange recurrence (PIMDB_EXCH_DEBUG=1)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ["PIMDB_EXCH_DEBUG"] = "1"
from pimd_b_b200 import workloads as wl
from pimd_b_b200.config import SimConfig
from pimd_b_b200.engine import DeviceSim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = SimConfig(nbeads=4, natoms=n, ndim=3, bosonic=True, fixcom=False, pbc=False, temperature=5 * wl.KELVIN,
                mass=4.0026 * wl.DALTON, size=100 * wl.ANGSTROM, interaction="free", external="harmonic",
                ext_omega=3 * wl.MEV, thermostat="none")
rng = np.random.default_rng(n)
x = np.repeat(rng.normal(0, 20.0, size=(1, n, 3)), 4, axis=0) + rng.normal(0, 1.0, size=(4, n, 3))
sim = DeviceSim(cfg); sim.set("x", x)
arr = (C.c_double * (2 + 192))()
sim.lib.pimdb_debug_exchange_timing.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
sim.lib.pimdb_debug_exchange_timing(sim.h, 20, arr)
print("recur+forces us", arr[1])
d = np.array(arr[2:]).reshape(2, 32, 3)
for name, blk in zip(("fwd", "bwd"), d):
    nw = (n + 31) // 32
    print(name, "consumer cycles:", blk[:nw, 0].astype(int).tolist())
    print(name, "owner cycles:   ", blk[:nw, 1].astype(int).tolist())
    print(name, "owner prologue: ", blk[:nw, 2].astype(int).tolist())
