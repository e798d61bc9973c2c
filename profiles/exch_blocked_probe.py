"""Warm timing of the exchange chain (coefficients | recurrences + forces) and clock64 breakdown of the blocked
recurrence on the C3 exchange problem (N = 512 He-4 atoms). Usage: python profiles/exch_blocked_probe.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ["PIMDB_EXCH_DEBUG"] = "1"
os.environ["PIMDB_EXCH_DEBUG_FULL"] = "1"
os.environ["PIMDB_EXCH_REASONS"] = "1"
from pimd_b_b200 import workloads as wl
from pimd_b_b200.engine import DeviceSim
cfg = wl.config("c3")
x, p = wl.initial_state(cfg, "c3")
sim = DeviceSim(cfg); sim.set("x", x); sim.set("p", p)
sim.update_forces()
KD = 64 * 3 + 2048 + 32
arr = (C.c_double * (2 + KD))()
sim.lib.pimdb_debug_exchange_timing.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
sim.lib.pimdb_debug_exchange_timing(sim.h, 50, arr)
print("coeff us", arr[0], "recur+forces us", arr[1])
raw = np.array(arr[2:])
d = raw[:192].reshape(2, 32, 3)
st = raw[192:192 + 2048].reshape(2, 16, 16, 4)
own = raw[192 + 2048:].reshape(2, 16)
for di, name in enumerate(("fwd", "bwd")):
    blk = d[di]
    t0 = blk[:16, 2].min()
    print(name, "consumer cycles:", blk[:16, 0].astype(int).tolist())
    print(name, "owner cycles:   ", blk[:16, 1].astype(int).tolist())
    print(name, "start stamp:    ", (blk[:16, 2] - t0).astype(int).tolist())
    order = range(16) if di == 0 else range(15, -1, -1)
    print(name, "owner start/end (rel):", [(int(own[di, w] - t0), int(own[di, w] - t0 + blk[w, 1])) for w in order])
    w = 15 if di == 0 else 0    # the last owner: consumes every earlier block
    print(name, f"warp {w} per block [flag seen, +half0, +half1, +applied]:")
    for pos in range(15):
        s = st[di, w, pos]
        print("   pos", pos, int(s[0] - t0), int(s[1] - s[0]), int(s[2] - s[0]), int(s[3] - s[0]))
    w = 8 if di == 0 else 7
    print(name, f"warp {w} per block [flag seen, +half0, +half1, +applied]:")
    for pos in range(8):
        s = st[di, w, pos]
        print("   pos", pos, int(s[0] - t0), int(s[1] - s[0]), int(s[2] - s[0]), int(s[3] - s[0]))
fn = sim.lib.pimdb_debug_exchange_blocks; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
buf = (C.c_int * 64)(); nb = fn(sim.h, buf); print("block status (+16*reason)", list(buf[:2 * nb]))
