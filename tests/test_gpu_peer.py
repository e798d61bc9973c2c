"""Bead shards coupled through peer memory (pimdb_peer_export / pimdb_peer_attach; csrc/integrator.cu): the sharded CUDA
path -- halo slices and momentum sums stored into the peers' memory by the step's own kernels, one CUDA graph per shard,
no host collective -- against a single handle that owns every bead, and against the oracle.

Reference: getPrev/NextCoords (src/simulation.cpp:299-347, MPI_Sendrecv) and zeroMomentum (:581-603, MPI_Allreduce).
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from tests.helpers import Oracle, relerr
from tests.peer_cases import make_case

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _single(cfg, x, p, nsteps):
    from pimd_b_b200.engine import DeviceSim
    sim = DeviceSim(cfg, device=0)
    sim.set("x", x)
    sim.set("p", p)
    sim.update_forces()
    f0 = sim.get("f")
    sim.step(nsteps)
    out = dict(x=sim.get("x"), p=sim.get("p"), f=sim.get("f"), f0=f0, obs=sim.observables())
    sim.close()
    return out


def _in_process(cfg, x, p, nsteps, world, devices=None):
    """`world` handles in THIS process (plain device pointers instead of cudaIpc mappings), attached to each other."""
    from pimd_b_b200.engine import DeviceSim
    from pimd_b_b200.distributed import bead_range
    devices = devices or [0] * world
    sims, ranges = [], []
    for r in range(world):
        lo, hi = bead_range(cfg.nbeads, world, r)
        sims.append(DeviceSim(cfg, lo, hi, devices[r]))
        ranges.append((lo, hi))
    blobs = b"".join(s.peer_export() for s in sims)
    for r, s in enumerate(sims):
        s.peer_attach(world, r, blobs)
        assert s.peer_attached
    for s, (lo, hi) in zip(sims, ranges):
        s.upload(x[lo:hi], p[lo:hi])
    for s in sims:
        s.update_forces()
    f0 = np.concatenate([s.get("f") for s in sims])
    for s in sims:              # asynchronous: every handle enqueues all its graph replays, the devices sort out the rest
        s.step(nsteps)
    for s in sims:              # reading the momenta is collective and blocking: enqueue its deferred part everywhere first
        s.settle()
    obs = [s.observables() for s in sims]
    out = dict(x=np.concatenate([s.get("x") for s in sims]), p=np.concatenate([s.get("p") for s in sims]),
               f=np.concatenate([s.get("f") for s in sims]), f0=f0,
               obs={k: sum(o[k] for o in obs) for k in obs[0]})
    for s in sims:
        s.close()
    return out


def _compare(got, ref, tol_x=1e-10, tol_p=1e-9, tol_f=1e-9):
    assert relerr(got["f0"], ref["f0"]) < 1e-12          # same kernels on the same positions
    assert relerr(got["x"], ref["x"]) < tol_x
    assert relerr(got["p"], ref["p"]) < tol_p
    assert relerr(got["f"], ref["f"]) < tol_f


@pytest.mark.parametrize("case,world", [("he_langevin", 2), ("he_langevin", 4), ("he_nve_odd", 2), ("he_nve_odd", 3),
                                        ("trap_nofixcom", 2), ("trap_nofixcom", 3), ("dist_nh", 2), ("dist_nh", 4),
                                        ("nm_langevin", 2), ("nm_langevin", 4), ("nm_nve_fixcom", 2), ("nm_nve_fixcom", 4)])
def test_peer_shards_in_one_process_match_a_single_handle(gpu_required, case, world):
    cfg, x, p = make_case(case)
    nsteps = 12
    ref = _single(cfg, x, p, nsteps)
    got = _in_process(cfg, x, p, nsteps, world)
    _compare(got, ref)
    for k, v in ref["obs"].items():
        if np.isfinite(v):
            assert abs(got["obs"][k] - v) <= 1e-8 * max(1.0, abs(v)) + 1e-9 * abs(v), k


def test_peer_shards_follow_the_oracle(gpu_required):
    cfg, x, p = make_case("he_nve_odd")
    got = _in_process(cfg, x, p, 10, 3)
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.set("p", p)
    orc.update_forces()
    assert relerr(got["f0"], orc.get("f")) < 1e-10
    for _ in range(10):
        orc.run_iteration()
    assert relerr(got["x"], orc.get("x")) < 1e-10
    assert relerr(got["p"], orc.get("p")) < 1e-9


def test_peer_steps_split_over_calls_and_state_reads(gpu_required):
    """step(5) + read + step(7) == step(12): the lazily closed zeroMomentum and the halo hand-shake survive call boundaries."""
    from pimd_b_b200.engine import DeviceSim
    from pimd_b_b200.distributed import bead_range
    cfg, x, p = make_case("he_langevin")
    ref = _in_process(cfg, x, p, 12, 2)
    sims = []
    for r in range(2):
        lo, hi = bead_range(cfg.nbeads, 2, r)
        sims.append(DeviceSim(cfg, lo, hi, 0))
    blobs = b"".join(s.peer_export() for s in sims)
    for r, s in enumerate(sims):
        s.peer_attach(2, r, blobs)
    for r, s in enumerate(sims):
        lo, hi = bead_range(cfg.nbeads, 2, r)
        s.upload(x[lo:hi], p[lo:hi])
    for s in sims:
        s.update_forces()       # (as _in_process does: the first kick then uses f(x0) in both runs)
    for s in sims:
        s.step(5)
    # reading the momenta is collective (it carries out the closing zeroMomentum): enqueue on both, then read
    for s in sims:
        s.settle()
    mid = [s.observables() for s in sims]
    assert all(np.isfinite(m["cl_kinetic"]) for m in mid)
    for s in sims:
        s.step(7)
    for s in sims:
        s.settle()
    got_x = np.concatenate([s.get("x") for s in sims])
    got_p = np.concatenate([s.get("p") for s in sims])
    for s in sims:
        s.close()
    assert relerr(got_x, ref["x"]) < 1e-10
    assert relerr(got_p, ref["p"]) < 1e-9


def test_peer_wait_is_bounded(gpu_required, monkeypatch):
    """A peer that never makes its calls surfaces as a RuntimeError after the time-out, not as a hang."""
    from pimd_b_b200.engine import DeviceSim
    from pimd_b_b200.distributed import bead_range
    monkeypatch.setenv("PIMDB_PEER_TIMEOUT_MS", "200")
    cfg, x, p = make_case("trap_nofixcom")
    sims = []
    for r in range(2):
        lo, hi = bead_range(cfg.nbeads, 2, r)
        sims.append(DeviceSim(cfg, lo, hi, 0))
    blobs = b"".join(s.peer_export() for s in sims)
    for r, s in enumerate(sims):
        s.peer_attach(2, r, blobs)
    lo, hi = bead_range(cfg.nbeads, 2, 0)
    sims[0].upload(x[lo:hi], p[lo:hi])       # rank 1 never uploads: rank 0's halo push waits for a hand-shake that never comes
    with pytest.raises(RuntimeError, match="timed out waiting for a peer"):
        sims[0].synchronize()
    for s in sims:
        s.close()


def _run_workers(case, nsteps, world, tmp_path, same_gpu):
    out = tmp_path / "peer_out.npz"
    env = dict(os.environ, PIMDB_TEST_SAME_GPU="1" if same_gpu else "0", PYTHONPATH=str(ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + os.getpid() % 400), str(ROOT / "tests" / "peer_worker.py"), case, str(nsteps), str(out)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    d = np.load(out)
    return dict(x=d["x"], p=d["p"], f=d["f"], f0=d["f0"], obs=dict(zip(d["obs_keys"].tolist(), d["obs"].tolist())))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("case", ["he_langevin", "trap_nofixcom", "nm_langevin"])
def test_peer_shards_in_separate_processes_one_gpu(gpu_required, tmp_path, case):
    """Two processes on cuda:0, coupled through cudaIpc mappings (the multi-process path on a one-GPU box; the two
    contexts are time-sliced, so every hand-shake costs a context switch -- correctness only)."""
    cfg, x, p = make_case(case)
    ref = _single(cfg, x, p, 6)
    got = _run_workers(case, 6, 2, tmp_path, same_gpu=True)
    _compare(got, ref)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_shards_one_process_per_gpu(gpu_required, tmp_path, world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cfg, x, p = make_case("he_langevin")
    ref = _single(cfg, x, p, 12)
    got = _run_workers("he_langevin", 12, world, tmp_path, same_gpu=False)
    _compare(got, ref)
    for k, v in ref["obs"].items():
        if np.isfinite(v):
            assert abs(got["obs"][k] - v) <= 1e-8 * max(1.0, abs(v)) + 1e-9 * abs(v), k
