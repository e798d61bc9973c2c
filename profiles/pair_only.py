"""The pair tiles + assembly alone on a BASELINE workload (distinguishable copy): target of the ncu captures of the pair
kernel.   python profiles/pair_only.py c3 [calls]"""
import dataclasses, sys
sys.path.insert(0, ".")
from pimd_b_b200 import workloads as wl
from pimd_b_b200.engine import DeviceSim
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = wl.config(name)
x, p = wl.initial_state(cfg, name)
d = DeviceSim(dataclasses.replace(cfg, bosonic=False, obs_bosonic="false")); d.set("x", x); d.set("p", p)
for _ in range(n): d.update_forces()
d.synchronize()
