"""Where the end-to-end step (host buffers in, host buffers out) spends its time on C3: each leg alone, wall clock per call
with a device synchronisation after it, page-locked buffers.   python profiles/e2e_probe.py"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pimd_b_b200 import workloads as wl
from pimd_b_b200.engine import DeviceSim
cfg = wl.config("c3")
x, p = wl.initial_state(cfg, "c3")
sim = DeviceSim(cfg)
pin = torch.empty((3,) + x.shape, dtype=torch.float64, pin_memory=True)
hx, hp, hf = (pin[i].numpy() for i in range(3))
hx[:] = x; hp[:] = p
sim.upload(hx, hp); sim.step(5); sim.synchronize()
def t(fn, n=200):
    fn(); sim.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn(); sim.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
print("upload(x,p) + sync            %7.1f us" % t(lambda: sim.upload(hx, hp)))
print("step(1) + sync                %7.1f us" % t(lambda: sim.step(1)))
print("download(x,p,f)               %7.1f us" % t(lambda: sim.download(hx, hp, hf)))
print("download(p,f)                 %7.1f us" % t(lambda: sim.download(None, hp, hf)))
print("step_download(1,x,p,f)        %7.1f us" % t(lambda: sim.step_download(1, hx, hp, hf)))
print("upload + step_download        %7.1f us" % t(lambda: (sim.upload(hx, hp), sim.step_download(1, hx, hp, hf))))
# raw copies of the same sizes
d = torch.empty_like(pin, device="cuda")
def cp_h2d(): d[:2].copy_(pin[:2], non_blocking=True)
def cp_d2h(): pin[1:].copy_(d[1:], non_blocking=True)
def ts(fn, n=200):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
print("raw H2D 1.57 MB + sync        %7.1f us" % ts(cp_h2d))
print("raw D2H 1.57 MB + sync        %7.1f us" % ts(cp_d2h))
big = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True); dbig = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
print("raw H2D 64 MB: %.1f GB/s" % (64 * 1.048576e-3 / (ts(lambda: dbig.copy_(big, non_blocking=True), 20) * 1e-6) / 1e3 * 1e3 / 1e3))
print("raw D2H 64 MB: %.1f GB/s" % (64 * 1.048576e-3 / (ts(lambda: big.copy_(dbig, non_blocking=True), 20) * 1e-6) / 1e3 * 1e3 / 1e3))
