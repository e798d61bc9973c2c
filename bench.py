#!/usr/bin/env python
"""bench.py — PIMD steps/s of the force-and-propagate hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

* default arm: the CUDA path through the C ABI (libpimdb200.so). One "step" = one iteration of the body of
  Simulation::run (thermostat half step, COM removal, B, A, forces = pair + exchange + springs, B, thermostat half
  step, COM removal) on the headline system (He-4 Aziz, N=512, P=64, PBC); estimators are evaluated every `sfreq`
  steps inside the timed region. State is resident in HBM for `value`; `e2e` drives the same step through the
  host-buffer API (upload x,p / step / download x,p,f every step, page-locked host buffers).
* N>1 (torchrun): beads are sharded over the ranks (pimd_b_b200.distributed.PeerShardedSimulation): halo slices and the
  momentum sums are stored into the peers' memory over NVLink by the step's own kernels, one CUDA-graph replay per rank
  and step, no host collective inside the timed region; time = max over ranks; the system is the same at every N, so
  scaling is "strong". PIMDB_SHARD_MODE=nccl selects the host-driven NCCL choreography of round 1 for comparison.
* `roofline` (pair-tile kernel, FP64 pipe): the kernel timed alone -- CUDA events around 20 x the step's pair-tile launches,
  back to back -- against the DFMA peak measured in this run; `roofline.in_eager_step` = the duration between an event
  pair in an eager step (it also times the exchange kernels that run beside it), `roofline.in_captured_step` = the
  durations of the force-phase kernels inside the graph-replayed step from %globaltimer stamps (N=1 only; a second handle
  created with PIMDB_TIMELINE=1). `roofline_integrator`: achieved GB/s of the fused integrator launches (eager pass).
* The line also carries `c4`: the same measurement on He-4 N=2048 P=128 (BASELINE configs[3], the configuration that
  names bead sharding), at every N.
* --impl reference: the reference's own CPU implementation (oracle/_ref/pimdb_ndim3, unmodified sources, one
  process per bead on all host cores) on the same configuration and initial state.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

from pimd_b_b200 import workloads  # noqa: E402

METRIC = "PIMD steps/s, He-4 Aziz N=512 P=64, 1/2/4/8 B200 vs host-CPU MPI"
L2_FLUSH_BYTES = 256 << 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sfreq", type=int, default=1000, help="estimators every sfreq steps (reference default 1000)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the extra C4 measurement")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps (diagnostic)")
    return ap.parse_args()


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def config_dict(name, cfg):
    """Identical in both arms (the driver compares them)."""
    return {"workload": workloads.DESCRIPTION[name], "natoms": cfg.natoms, "nbeads": cfg.nbeads, "ndim": cfg.ndim}


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines = []
        self.idx = gpu_index
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [s.strip() for s in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        top = sorted(sm)[len(sm) // 2:]   # under load = upper half of the samples
        return {"sm_mhz": float(np.median(top)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- reference arm
def run_reference(args, cfg, x, p, budget_s: float):
    from oracle import refrun
    probe = refrun.time_reference(cfg, x, p, steps=2)
    t = probe["sec_per_step"]
    n = int(max(3, min(args.steps, budget_s / max(t, 1e-9))))
    res = refrun.time_reference(cfg, x, p, steps=n)
    return res, n


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workloads.config(args.workload)
    x, p = workloads.initial_state(cfg, args.workload)
    res, n = run_reference(args, cfg, x, p, budget_s=90.0)
    sps = 1.0 / res["sec_per_step"]
    line = {
        "metric": METRIC, "value": sps, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * res["sec_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": config_dict(args.workload, cfg),
        "details": {"observables": "off"},
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": res["cores"], "kind": res["kind"],
                         "sample": f"{n} MD steps of the full workload after a 2-step warm-up run; "
                                   f"{res['ranks']} processes (one per bead) on {res['cores']} host cores, observables off"},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
class Runner:
    """One workload on this rank's GPU: a full ring (world 1) or this rank's bead shard."""

    def __init__(self, name, world, rank, local_rank, mode):
        import torch
        from pimd_b_b200.engine import DeviceSim
        self.torch = torch
        self.name, self.world, self.rank, self.local_rank = name, world, rank, local_rank
        self.cfg = workloads.config(name)
        self.x, self.p = workloads.initial_state(self.cfg, name)
        self.dev = torch.device("cuda", local_rank)
        self.mode = mode if world > 1 else "single"
        cfg = self.cfg
        if world == 1:
            self.stream = torch.cuda.Stream(device=self.dev)
            self.sim = DeviceSim(cfg, device=local_rank)
            self.sim.set_stream(self.stream.cuda_stream)
            self.lo, self.hi = 0, cfg.nbeads
            self.sim.upload(self.x, self.p)
            self.step = self.sim.step
            self.observe = self.sim.observables
            self.finish = lambda: None
            self.launch_note = "CUDA graph replay (pimdb_step)"
        elif self.mode == "peer":
            from pimd_b_b200.distributed import PeerShardedSimulation
            self.ps = PeerShardedSimulation(cfg, rank, world, local_rank)
            self.sim, self.stream = self.ps.sim, self.ps.stream
            self.lo, self.hi = self.ps.lo, self.ps.hi
            self.ps.set_state(self.x, self.p)
            self.step = self.ps.step
            self.observe = self.ps.observables
            self.finish = lambda: None   # the closing zeroMomentum left open between steps is carried out by whatever reads p
            self.launch_note = ("one CUDA graph replay per rank and step; halo slices and momentum sums stored into the "
                                "peers' memory by the step's kernels (no NCCL / host call inside the step)")
        else:   # round-1 choreography: host-driven phases + NCCL
            from pimd_b_b200.distributed import CudaShard, ShardedSimulation, bead_range
            shard = CudaShard(cfg, rank, world, local_rank)
            self.stream, self.sim = shard.stream, shard.sim
            self.lo, self.hi = bead_range(cfg.nbeads, world, rank)
            self.sim.set("x", self.x[self.lo:self.hi])
            self.sim.set("p", self.p[self.lo:self.hi])
            self.driver = ShardedSimulation(cfg, shard, halo=os.environ.get("PIMDB_SHARD_HALO", "p2p"))
            with torch.cuda.stream(self.stream):
                self.driver.exchange_halos()
            self.step = lambda n=1: self.driver.step(n, finalize=False)
            self.observe = self.driver.observables
            self.finish = self.driver.flush
            self.launch_note = "eager phases + NCCL point-to-point halos and all-reduces (PIMDB_SHARD_MODE=nccl)"

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        import torch.distributed as dist
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        self.barrier()
        self.sim.close()


def measure(run: Runner, K: int, W: int, sfreq: int, flush, clocks_rank0: bool, do_e2e: bool):
    """W untimed steps, exactly K timed steps (CUDA events per step on the launching stream, L2 flushed in between),
    a back-to-back pass, and the end-to-end pass through the host-buffer API."""
    torch = run.torch
    sim, stream = run.sim, run.stream
    out = {}
    with torch.cuda.stream(stream):
        run.step(W)
        run.observe()
        run.barrier()
        clocks = ClockSampler(run.dev.index) if clocks_rank0 and run.rank == 0 else None
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        l0 = sim.launch_count
        run.barrier()
        wall0 = time.perf_counter()
        for i in range(K):
            if flush is not None:
                flush.zero_()
            ev[i][0].record(stream)
            run.step(1)
            if (i + 1) % sfreq == 0:
                run.observe()
            if i == K - 1:
                run.finish()
            ev[i][1].record(stream)
        run.barrier()
        wall1 = time.perf_counter()
        out["gpu_launches"] = int(sim.launch_count - l0)
        total_ms = run.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
        out["clocks"] = clocks.stop() if clocks else None
        out["ms_per_step"] = total_ms / K
        out["value"] = 1e3 * K / total_ms
        out["wall_s_timed_region"] = wall1 - wall0

        # back-to-back replay without flushes (diagnostic: what a production trajectory sees)
        run.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run.step(K)
        e1.record(stream)
        run.barrier()
        out["ms_per_step_back_to_back"] = run.max_over_ranks(e0.elapsed_time(e1) / K)
        out["steps_per_s_back_to_back"] = 1e3 / out["ms_per_step_back_to_back"]

        if do_e2e:
            # the same step through the host-buffer API: upload x,p -> step -> download x,p,f, page-locked host buffers
            # (the library copies straight from / into them), one transpose kernel and one synchronisation per direction
            shape = run.x[run.lo:run.hi].shape
            pin = torch.empty((3,) + tuple(shape), dtype=torch.float64, pin_memory=True)   # x | p | f back to back: one copy each way
            hx, hp, hf = (pin[i].numpy() for i in range(3))
            sim.download(hx, hp, None)
            Ke = max(20, min(K, 300))
            def host_step():
                # what a host loop that keeps its own copy of the state pays per step: two calls
                sim.upload(hx, hp)
                if run.mode == "nccl":
                    run.step(1); run.finish(); sim.download(hx, hp, hf)    # flush the host-driven phases before reading back
                else:
                    sim.step_download(1, hx, hp, hf)      # = step(1) + download(x, p, f), x leaving while the forces are computed
            for _ in range(3):
                host_step()
            run.barrier()
            t0 = time.perf_counter()
            for i in range(Ke):
                host_step()
            run.barrier()
            e2e_s = run.max_over_ranks((time.perf_counter() - t0) / Ke)
            slab = int(np.prod(shape)) * 8
            out["e2e"] = {"value": 1.0 / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": 2 * slab * run.world,
                          "d2h_bytes_per_step": 3 * slab * run.world, "steps": Ke,
                          "call": "pimdb_upload_state(x,p) + pimdb_step_download(1,x,p,f), page-locked host buffers"}
    return out


def rooflines(run: Runner, hbm_peak, peak_src, fp64_peak):
    """Pair-force kernel (FP64 pipe) and fused integrator (HBM), from an eager pass with CUDA events around every
    launch of those kernels (every rank takes part: the steps are collective; rank 0 reports its own kernels)."""
    torch, sim, cfg = run.torch, run.sim, run.cfg
    nsteps = 40
    with torch.cuda.stream(run.stream):
        sim.timing_enable(True)
        for _ in range(4):
            # The eager, event-bracketed launches cost the host more than the kernels cost the device; with the device waiting
            # for the host, an event pair also times the gap until the launch arrives. So the device is parked (~1.5 ms)
            # while the host enqueues ten steps, and then runs them back to back.
            torch.cuda._sleep(3_000_000)
            run.step(nsteps // 4)
        run.barrier()
        pair_ms, npair = sim.timing_get(0)
        step_ms, _ = sim.timing_get(1)
        integ_ms, ninteg = sim.timing_get(2)
        integ_bytes = sim.timing_integrator_bytes()
        sim.timing_enable(False)
    nloc = run.hi - run.lo
    roofline = None
    if cfg.interaction != "free" and npair:
        per_step = max(1, npair // nsteps)
        flops = workloads.pair_flops_per_step(cfg) * nloc / cfg.nbeads / per_step
        # the kernel with the GPU to itself: 20 x (the pair-tile launches of one step) back to back between two CUDA events
        import ctypes as C
        fn = sim.lib.pimdb_debug_pair_tiles_only
        fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_int]
        with torch.cuda.stream(run.stream):
            fn(sim.h, 3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(run.stream); fn(sim.h, 20); e1.record(run.stream)
            e1.synchronize()
        alone_ms = e0.elapsed_time(e1) / 20 / per_step
        in_step_ms = pair_ms
        pair_ms = alone_ms
        ach = flops / (pair_ms * 1e-3) * 1e-12
        traffic, fp64_pipe = None, {}
        try:
            prof = json.loads((ROOT / "profiles" / "r02_ncu_summary.json").read_text())
            if run.name == "c3" and run.world == 1:
                traffic = prof["kernels"]["k_pair_tiles"]["dram_traffic_bytes"]
            fp64_pipe = {k: v.get("fp64_pipe_pct") for k, v in prof["kernels"].items()}
        except Exception:
            pass
        roofline = {"kernel": "k_pair_tiles", "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": ach / fp64_peak if fp64_peak else None, "traffic": traffic,
                    "traffic_source": "profiles/r02_ncu_summary.json (ncu --set full, bytes per launch, N=1 C3 capture)",
                    "peak_source": "DFMA micro-benchmark in this run (MEASURED_PEAKS.json has no FP64 figure)",
                    "flops_per_launch": flops, "ms_per_launch": pair_ms,
                    "timing": "CUDA events around 20 x the step's pair-tile launches, back to back, nothing beside them "
                              "(the kernel timed alone: a step overlaps it with the exchange chain, which an event pair in an "
                              "eager step charges to it -- in_eager_step)",
                    "in_eager_step": {"ms_per_launch": in_step_ms, "frac": flops / (in_step_ms * 1e-3) * 1e-12 / fp64_peak if fp64_peak else None,
                                      "launches_timed": npair, "eager_step_ms": step_ms,
                                      "share_of_step": in_step_ms * (npair / nsteps) / step_ms if step_ms else None},
                    "fp64_pipe_pct_ncu": fp64_pipe}
        if run.world == 1:
            tl = captured_step_timeline(run.name, run.local_rank, torch)
            if tl:
                roofline["in_captured_step"] = {**tl, "frac": flops * per_step / (tl["pair_tiles_ms"] * 1e-3) * 1e-12 / fp64_peak if fp64_peak else None}
    ach_gbs = integ_bytes / (integ_ms * 1e-3) * 1e-9 if integ_ms else None
    integ = {"kernel": "k_integrate", "bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
             "frac": ach_gbs / hbm_peak if ach_gbs else None, "peak_source": peak_src,
             "bytes_per_launch": integ_bytes, "ms_per_launch": integ_ms, "launches_per_step": ninteg / nsteps,
             "note": "algorithmic bytes (p, f, x once per stage group + pair partials where the closing kick assembles the "
                     "forces) / CUDA-event time per launch; at C3 the state is L2-resident and a launch is bounded by "
                     "launch latency (~2 us), not by HBM",
             "bytes_per_step_survey": workloads.integrator_bytes_per_step(cfg) * nloc / cfg.nbeads}
    return roofline, integ


def captured_step_timeline(name, local_rank, torch):
    """How long the kernels of the captured (graph-replayed) step take where they run, overlapped: a second handle created with
    PIMDB_TIMELINE=1 has every block stamp %globaltimer; [first block start, last block end] per kernel, median of 20 steps."""
    import ctypes as C
    from pimd_b_b200.engine import DeviceSim
    cfg = workloads.config(name)
    x, p = workloads.initial_state(cfg, name)
    os.environ["PIMDB_TIMELINE"] = "1"
    try:
        sim = DeviceSim(cfg, device=local_rank)
    finally:
        del os.environ["PIMDB_TIMELINE"]
    try:
        sim.upload(x, p)
        sim.step(30); sim.synchronize()
        fn = sim.lib.pimdb_debug_timeline; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
        fk = sim.lib.pimdb_debug_timeline_kinds; fk.restype = C.c_int; fk.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte)]
        buf, kinds = (C.c_ulonglong * 64)(), (C.c_ubyte * 32)()
        fn(sim.h, buf)
        rows = []
        for _ in range(20):
            sim.step(1); sim.synchronize()
            n = min(fn(sim.h, buf), 32)
            fk(sim.h, kinds)
            t = np.array(buf[:], dtype=np.uint64).reshape(32, 2)[:n].astype(np.int64)
            k = np.array(kinds[:n])
            def span(kind):
                sel = t[k == kind]
                return float(sel[:, 1].max() - sel[:, 0].min()) * 1e-6 if len(sel) else None
            rows.append((span(5), span(2), span(3), span(4), float(t[:, 1].max() - t[:, 0].min()) * 1e-6))
        med = lambda i: float(np.median([r[i] for r in rows])) if rows and rows[0][i] is not None else None
        return {"pair_tiles_ms": med(0), "exchange_factor_tiles_ms": med(1), "exchange_recurrences_ms": med(2),
                "exchange_exterior_forces_ms": med(3), "step_span_ms": med(4),
                "method": "%globaltimer stamps by every block, first start -> last end per kernel kind, captured step, warm"}
    except Exception:
        return None
    finally:
        sim.close()


def scrambled_pair_time(name, local_rank, torch):
    """The pair kernel's far-rotation shortcut depends on how well particle order follows space: time it again with the
    particle order scrambled (same positions, random labels)."""
    from pimd_b_b200.engine import DeviceSim
    cfg = workloads.config(name)
    x, p = workloads.initial_state(cfg, name)
    perm = np.random.default_rng(7).permutation(cfg.natoms)
    sim = DeviceSim(cfg, device=local_rank)
    sim.upload(np.ascontiguousarray(x[:, perm]), np.ascontiguousarray(p[:, perm]))
    st = torch.cuda.Stream(device=local_rank)
    sim.set_stream(st.cuda_stream)
    import ctypes as C
    fn = sim.lib.pimdb_debug_pair_tiles_only
    fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_int]
    with torch.cuda.stream(st):
        sim.step(5)                       # (a few steps so that the beads are no longer on top of each other)
        fn(sim.h, 3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); fn(sim.h, 20); e1.record(st)
        e1.synchronize()
    ms = e0.elapsed_time(e1) / 20
    sim.close()
    return ms


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_main(args)
        return

    import torch
    import torch.distributed as dist
    from pimd_b_b200 import _cabi
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    mode = os.environ.get("PIMDB_SHARD_MODE", "peer")

    hbm_peak, peak_src = load_peaks()
    K, W = args.steps, max(args.warmup, 3)
    flush = None if args.no_flush else torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    lib = _cabi.load()
    pk = C.c_double()
    lib.pimdb_bench_fp64_peak(local_rank, C.byref(pk))
    fp64_peak = pk.value

    # ---- headline workload
    run = Runner(args.workload, world, rank, local_rank, mode)
    cfg = run.cfg
    m = measure(run, K, W, args.sfreq, flush, clocks_rank0=True, do_e2e=True)
    roofline, integ = rooflines(run, hbm_peak, peak_src, fp64_peak) if mode != "nccl" or world == 1 else (None, None)
    launch_note = run.launch_note
    x0, p0 = run.x, run.p
    run.close()
    if roofline is not None and world == 1 and rank == 0:
        try:
            ms_scr = scrambled_pair_time(args.workload, local_rank, torch)
            roofline["ms_per_launch_scrambled_order"] = ms_scr
            roofline["frac_scrambled_order"] = roofline["flops_per_launch"] / (ms_scr * 1e-3) * 1e-12 / fp64_peak
        except Exception as exc:
            roofline["ms_per_launch_scrambled_order"] = f"failed: {exc}"

    # ---- C4 (the configuration that names bead sharding), at every N
    c4 = None
    if args.workload == "c3" and not args.no_c4:
        try:
            run4 = Runner("c4", world, rank, local_rank, mode)
            K4 = max(20, min(K, 200))
            m4 = measure(run4, K4, 5, 10 ** 9, flush, clocks_rank0=False, do_e2e=False)
            r4, i4 = rooflines(run4, hbm_peak, peak_src, fp64_peak) if mode != "nccl" or world == 1 else (None, None)
            c4 = {"config": config_dict("c4", run4.cfg), "steps_per_s": m4["value"], "ms_per_step": m4["ms_per_step"], "steps": K4,
                  "steps_per_s_back_to_back": m4["steps_per_s_back_to_back"], "gpu_launches": m4["gpu_launches"],
                  "frac": r4["frac"] if r4 else None, "roofline": r4, "roofline_integrator": i4}
            run4.close()
        except Exception as exc:   # never lose the headline line to the extra measurement
            c4 = {"failed": repr(exc)}

    # ---- CPU baseline (rank 0, N=1 only): the reference's own implementation on the host cores
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            res, n = run_reference(args, cfg, x0, p0, budget_s=args.cpu_seconds)
            cpu_baseline = {"value": 1.0 / res["sec_per_step"], "unit": "steps/s", "cores": res["cores"],
                            "kind": res["kind"],
                            "sample": f"{n} MD steps of the full workload (same config and initial state), "
                                      f"{res['ranks']} processes (one per bead) on {res['cores']} host cores, observables off"}
        except Exception as exc:  # the baseline is reported, never the target: do not lose the GPU numbers
            cpu_baseline = {"value": None, "unit": "steps/s", "cores": os.cpu_count(), "kind": "reference",
                            "sample": f"failed: {exc}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": m["value"], "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, cfg),
            "details": {"parallelism": f"bead-sharded x{world}" if world > 1 else "single GPU",
                        "l2": "flushed between timed steps (256 MiB write)" if flush is not None else "not flushed",
                        "estimators_every": args.sfreq,
                        "timing": "CUDA events per step on the launching stream, max over ranks", "launch": launch_note},
            "ms_per_step_back_to_back": m["ms_per_step_back_to_back"], "steps_per_s_back_to_back": m["steps_per_s_back_to_back"],
            "wall_s_timed_region": m["wall_s_timed_region"],
            "e2e": m["e2e"], "gpu_launches": m["gpu_launches"], "clocks": m["clocks"],
            "roofline": roofline, "roofline_integrator": integ, "c4": c4, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
