"""In-kernel %globaltimer timeline of ONE captured MD step (graph replay) on the C3 workload: when each kernel's first
block started and its last block ended, relative to the step's first kernel. Usage: python profiles/step_timeline.py [c3]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ["PIMDB_TIMELINE"] = "1"
from pimd_b_b200 import workloads as wl
from pimd_b_b200.engine import DeviceSim
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = wl.config(name)
x, p = wl.initial_state(cfg, name)
sim = DeviceSim(cfg); sim.set("x", x); sim.set("p", p)
sim.update_forces()
sim.step(50); sim.synchronize()
fn = sim.lib.pimdb_debug_timeline; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
buf = (C.c_ulonglong * 64)()
fn(sim.h, buf)                       # reset after warm-up
for rep in range(3):
    sim.step(1); sim.synchronize()
    n = fn(sim.h, buf)
    t = np.array(buf[:], dtype=np.uint64).reshape(32, 2)[:n].astype(np.int64)
    t0 = t[:, 0].min()
    print(f"step {rep}: kernels in launch order [start us, end us, duration us]")
    for i, (a, b) in enumerate(t):
        print(f"   #{i}: {(a - t0) / 1e3:8.2f} {(b - t0) / 1e3:8.2f} {(b - a) / 1e3:8.2f}")
    fs = sim.lib.pimdb_debug_integrate_stamps; fs.restype = C.c_int; fs.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
    sb = (C.c_ulonglong * 64)()
    k = fs(sim.h, sb)
    st = np.array(sb[:], dtype=np.uint64).reshape(8, 8).astype(np.int64)
    for i in range(min(k, 8)):   # block 0: start, counters read, sums in, waits done, main loop done, ticket taken; last block: 6, 7 = end
        print(f"   k_integrate launch {i}: phase stamps", " ".join(f"{(v - t0) / 1e3:7.2f}" if v else "      -" for v in st[i]))
