#!/usr/bin/env python
"""bench.py — PIMD steps/s of the force-and-propagate hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

* default arm: the CUDA path through the C ABI (libpimdb200.so). One "step" = one iteration of the body of
  Simulation::run (thermostat half step, COM removal, B, A, forces = pair + exchange + springs, B, thermostat half
  step, COM removal) on the headline system (He-4 Aziz, N=512, P=64, PBC); estimators are evaluated every `sfreq`
  steps inside the timed region. State is resident in HBM for `value`; `e2e` drives the same step through the
  host-buffer API (upload x,p / step / download x,p,f every step).
* N>1 (torchrun): beads are sharded over the ranks (pimd_b_b200.distributed), halo slices and the momentum sums
  move over NCCL; time = max over ranks; the system is the same, so scaling is "strong".
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
        "config": {"workload": workloads.DESCRIPTION[args.workload], "natoms": cfg.natoms, "nbeads": cfg.nbeads,
                   "ndim": cfg.ndim, "observables": "off"},
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": res["cores"], "kind": res["kind"],
                         "sample": f"{n} MD steps of the full workload after a 2-step warm-up run; "
                                   f"{res['ranks']} processes (one per bead) on {res['cores']} host cores, observables off"},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    if args.impl == "reference":
        reference_main(args)
        return

    import torch
    import torch.distributed as dist
    from pimd_b_b200.engine import DeviceSim
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

    cfg = workloads.config(args.workload)
    x, p = workloads.initial_state(cfg, args.workload)
    hbm_peak, peak_src = load_peaks()
    K, W = args.steps, max(args.warmup, 3)

    flush = None if args.no_flush else torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    if world == 1:
        stream = torch.cuda.Stream(device=dev)
        sim = DeviceSim(cfg, device=local_rank)
        sim.set_stream(stream.cuda_stream)
        sim.set("x", x)
        sim.set("p", p)
        stepper = lambda n=1: sim.step(n)
        observe = sim.observables
        launch_count = lambda: sim.launch_count
        lo, hi = 0, cfg.nbeads
    else:
        from pimd_b_b200.distributed import CudaShard, ShardedSimulation, bead_range
        shard = CudaShard(cfg, rank, world, local_rank)
        stream = shard.stream
        sim = shard.sim
        lo, hi = bead_range(cfg.nbeads, world, rank)
        sim.set("x", x[lo:hi])
        sim.set("p", p[lo:hi])
        driver = ShardedSimulation(cfg, shard, halo=os.environ.get("PIMDB_SHARD_HALO", "p2p"))
        with torch.cuda.stream(stream):
            driver.exchange_halos()
        # between consecutive steps the closing zeroMomentum of an iteration is subsumed by the first one of the next
        # (distributed.py: Z O Z = Z O); it is carried out before anything looks at the momenta (observables, the end
        # of the timed region)
        stepper = lambda n=1: driver.step(n, finalize=False)
        observe = driver.observables
        launch_count = lambda: sim.launch_count

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graph_note = "CUDA graph replay (pimdb_step)"
    launches_per_step = None
    with torch.cuda.stream(stream):
        # ---- warm-up (W >= 3 untimed steps; also captures the CUDA graph)
        stepper(W)
        observe()
        barrier()
        if world > 1:
            l_a = launch_count()
            stepper(1)
            launches_per_step = launch_count() - l_a      # kernels of one sharded step (NCCL kernels not counted)
            # Capturing the NCCL point-to-point calls hung on this pool's boxes (round 1): opt-in only.
            ok = driver.enable_graph() if os.environ.get("PIMDB_SHARD_GRAPH") == "1" else False
            if not ok and not hasattr(driver, "_graph_error"):
                driver._graph_error = "disabled (set PIMDB_SHARD_GRAPH=1 to try)"
            flags = torch.tensor([1.0 if ok else 0.0], device=dev)
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            if flags.item() < 0.5:
                driver._graph = None
            graph_note = ("torch CUDA graph of the sharded step incl. NCCL halo + all-reduce" if flags.item() > 0.5
                          else "eager phases + NCCL (graph capture unavailable: %s)" % getattr(driver, "_graph_error", "?"))
            barrier()

        # ---- timed region: exactly K steps, L2 flushed between steps, CUDA events on the launching stream
        clocks = ClockSampler(local_rank) if rank == 0 else None
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        l0 = launch_count()
        barrier()
        wall0 = time.perf_counter()
        for i in range(K):
            if flush is not None:
                flush.zero_()
            ev[i][0].record(stream)
            stepper(1)
            if (i + 1) % args.sfreq == 0:
                observe()
            if i == K - 1 and world > 1:
                driver.flush()       # the pending closing zeroMomentum belongs to the timed region
            ev[i][1].record(stream)
        barrier()
        wall1 = time.perf_counter()
        gpu_launches = launch_count() - l0
        if launches_per_step is not None and getattr(driver, "_graph", None) is not None:
            gpu_launches = launches_per_step * K + (launch_count() - l0)   # graph replays do not pass through the library's counter
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
        clk = clocks.stop() if clocks else None
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        ms_per_step = total_ms / K
        value = 1e3 / ms_per_step

        # ---- back-to-back replay without flushes (diagnostic: what a production trajectory sees)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        stepper(K)
        e1.record(stream)
        barrier()
        b2b_ms = e0.elapsed_time(e1) / K
        if world > 1:
            t = torch.tensor([b2b_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            b2b_ms = float(t.item())

        # ---- e2e: the same step through the host-buffer API (upload x,p -> step -> download x,p,f)
        nloc = hi - lo
        # page-locked host buffers (the library copies straight from / into them; pageable buffers are staged)
        pin = [torch.empty(x[lo:hi].shape, dtype=torch.float64, pin_memory=True) for _ in range(3)]
        hx, hp, hf = (t.numpy() for t in pin)
        hx[...] = x[lo:hi]
        hp[...] = p[lo:hi]
        sim.get("x", hx)
        sim.get("p", hp)
        Ke = max(20, min(K, 200))
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            sim.set("x", hx)
            sim.set("p", hp)
            if world > 1:
                driver.exchange_halos()
            stepper(1)
            sim.get("x", hx)
            sim.get("p", hp)
            sim.get("f", hf)
        barrier()
        e2e_s = (time.perf_counter() - t0) / Ke
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        slab_bytes = nloc * cfg.natoms * cfg.ndim * 8
        e2e = {"value": 1.0 / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": 2 * slab_bytes * world,
               "d2h_bytes_per_step": 3 * slab_bytes * world, "steps": Ke,
               "call": "pimdb_set_state(x,p) + pimdb_step(1) + pimdb_get_state(x,p,f), page-locked host buffers"}

        # ---- roofline of the dominant kernel (pair-force tiles), eager pass with CUDA events per launch
        roofline = None
        integ = None
        if rank == 0:
            lib = _cabi.load()
            pk = C.c_double()
            lib.pimdb_bench_fp64_peak(local_rank, C.byref(pk))
            fp64_peak = pk.value
            if world == 1 and cfg.interaction != "free":
                sim.timing_enable(True)
                sim.step(50)
                pair_ms, npair = sim.timing_get(0)
                step_ms, _ = sim.timing_get(1)
                sim.timing_enable(False)
                flops = workloads.pair_flops_per_step(cfg) / max(1, (npair // 50))
                ach = flops / (pair_ms * 1e-3) * 1e-12
                traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
                try:
                    prof = json.loads((ROOT / "profiles" / "r01_ncu_summary.json").read_text())
                    if args.workload == "c3":
                        traffic = prof["kernels"]["k_pair_tiles"]["dram_traffic_bytes"]
                except Exception:
                    pass
                roofline = {"kernel": "k_pair_tiles", "bound": "fp64", "achieved": ach, "peak": fp64_peak,
                            "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None, "traffic": traffic,
                            "traffic_source": "profiles/r01_ncu_summary.json (ncu --set full, bytes per launch)",
                            "peak_source": "DFMA micro-benchmark in this run (MEASURED_PEAKS.json has no FP64 figure)",
                            "flops_per_launch": flops, "ms_per_launch": pair_ms, "launches_timed": npair,
                            "share_of_step": pair_ms * (npair / 50) / step_ms if step_ms else None,
                            "eager_step_ms": step_ms}
            integ = {"bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
                     "bytes_per_step": workloads.integrator_bytes_per_step(cfg)}

    # ---- CPU baseline (rank 0, N=1 only): the reference's own implementation on the host cores
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            res, n = run_reference(args, cfg, x, p, budget_s=args.cpu_seconds)
            cpu_baseline = {"value": 1.0 / res["sec_per_step"], "unit": "steps/s", "cores": res["cores"],
                            "kind": res["kind"],
                            "sample": f"{n} MD steps of the full workload (same config and initial state), "
                                      f"{res['ranks']} processes (one per bead) on {res['cores']} host cores, observables off"}
        except Exception as exc:  # the baseline is reported, never the target: do not lose the GPU numbers
            cpu_baseline = {"value": None, "unit": "steps/s", "cores": os.cpu_count(), "kind": "reference",
                            "sample": f"failed: {exc}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workloads.DESCRIPTION[args.workload], "natoms": cfg.natoms, "nbeads": cfg.nbeads,
                       "ndim": cfg.ndim, "parallelism": f"bead-sharded x{world}" if world > 1 else "single GPU",
                       "l2": "flushed between timed steps (256 MiB write)" if flush is not None else "not flushed",
                       "estimators_every": args.sfreq, "timing": "CUDA events per step on the launching stream, max over ranks",
                       "launch": graph_note},
            "ms_per_step_back_to_back": b2b_ms, "steps_per_s_back_to_back": 1e3 / b2b_ms,
            "wall_s_timed_region": wall1 - wall0,
            "e2e": e2e, "gpu_launches": int(gpu_launches), "clocks": clk,
            "roofline": roofline, "roofline_integrator": integ, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
