"""clock64 stamps of the cluster recurrence on the C3 exchange problem: per owner warp, cycles from seeing the previous
block's flag to [applied, values computed, remote stores issued, fence + flags issued]. Clocks of different SMs are
not synchronised, so only differences within one warp are meaningful."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ["PIMDB_EXCH_DEBUG"] = "1"; os.environ["PIMDB_EXCH_DEBUG_FULL"] = "1"
from pimd_b_b200 import workloads as wl
from pimd_b_b200.engine import DeviceSim
cfg = wl.config("c3"); x, p = wl.initial_state(cfg, "c3")
sim = DeviceSim(cfg); sim.set("x", x); sim.set("p", p); sim.update_forces()
KD = 64 * 3 + 2048 + 32
arr = (C.c_double * (2 + KD))()
sim.lib.pimdb_debug_exchange_timing.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
sim.lib.pimdb_debug_exchange_timing(sim.h, 50, arr)
print("tiles us", arr[0], "recur+forces us", arr[1])
st = np.array(arr[2 + 192:2 + 192 + 2048]).reshape(2, 16, 64)
for di, name in enumerate(("fwd", "bwd")):
    for w in range(16):
        s = st[di, w]
        print(name, "warp", w, "start->flag", int(s[0] - s[5]) if s[0] else None, "| flag->applied", int(s[1] - s[0]) if s[0] else int(s[1] - s[5]),
              "| ->computed", int(s[2] - s[1]), "| ->stores", int(s[3] - s[2]), "| ->flagged", int(s[4] - s[3]))
