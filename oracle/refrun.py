"""TEST / BENCH INFRASTRUCTURE ONLY: time the reference's own CPU implementation of the hot path.

Runs oracle/_ref/pimdb_ndim<d> (the UNMODIFIED reference sources built by oracle/Makefile against the fork-based
MPI stand-in; one process per bead, as the reference requires) on a given configuration and initial state, and
reads "Wall time per step (sec)" from output/report.txt (reference src/simulation.cpp:566-569; the loop runs
steps+1 iterations, so the figure is corrected by steps/(steps+1)). If the reference binary is absent, falls back
to the single-threaded C restatement (oracle/liboracle.so, kind "port").
"""
from __future__ import annotations

import ctypes as C
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from pimd_b_b200 import io as pio  # noqa: E402  (file formats only)


def ref_binary(ndim: int) -> Path:
    return HERE / "_ref" / f"pimdb_ndim{ndim}"


def time_reference(cfg, x, p, steps: int, observables: bool = False, timeout: float = 3600.0) -> dict:
    """One run of `steps` MD steps; returns {'sec_per_step', 'kind', 'cores', 'ranks', 'steps'}."""
    exe = ref_binary(cfg.ndim)
    cores = os.cpu_count() or 1
    if not exe.exists():
        return time_port(cfg, x, p, steps)
    tmp = Path(tempfile.mkdtemp(prefix="refrun_"))
    try:
        for b in range(cfg.nbeads):
            pio.write_xyz_positions(str(tmp / f"pos_{b}.xyz"), x[b])
            pio.write_manual_velocities(str(tmp / f"vel_{b}.dat"), p[b], cfg.mass)
        c = type(cfg)(**cfg.as_dict())
        c.steps = steps
        c.sfreq = max(steps, 1)
        c.threshold = 0.0
        c.initial_position = "xyz(pos_{}.xyz)"
        c.initial_velocity = "manual(vel_{}.dat)"
        if not observables:
            c.obs_energy = c.obs_classical = c.obs_bosonic = "off"
        (tmp / "config.ini").write_text(c.to_ini())
        env = dict(os.environ, PIMDB_NP=str(cfg.nbeads))
        t0 = time.perf_counter()
        r = subprocess.run([str(exe), "-in", "config.ini"], cwd=tmp, env=env, capture_output=True, text=True,
                           timeout=timeout)
        wall = time.perf_counter() - t0
        rep = tmp / "output" / "report.txt"
        if r.returncode != 0 or not rep.exists():
            raise RuntimeError("reference run failed: " + r.stdout[-400:] + r.stderr[-400:])
        m = re.search(r"Wall time per step \(sec\)\s*:?\s*([0-9.eE+-]+)", rep.read_text())
        if not m:
            raise RuntimeError("could not parse report.txt")
        per_step = float(m.group(1)) * steps / (steps + 1)
        return {"sec_per_step": per_step, "kind": "reference", "cores": cores, "ranks": cfg.nbeads,
                "steps": steps, "wall": wall}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def time_port(cfg, x, p, steps: int) -> dict:
    """Single-threaded C restatement (oracle/liboracle.so)."""
    sys.path.insert(0, str(ROOT))
    from tests.helpers import Oracle
    o = Oracle(cfg)
    o.set("x", x)
    o.set("p", p)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.run_iteration()
    dt = time.perf_counter() - t0
    o.close()
    return {"sec_per_step": dt / max(steps, 1), "kind": "port", "cores": 1, "ranks": 1, "steps": steps, "wall": dt}
