"""Bead sharding host logic (pimd_b_b200.distributed) on CPU: world_size 2 and 3 over the gloo backend.

The product shard wraps a GPU handle; here a NumPy stand-in implements the same four phases for a system simple
enough to restate in a few lines (distinguishable ring polymers in a harmonic trap, velocity Verlet, fixcom, no
noise). What is under test is the choreography: bead ranges, which slice goes to which neighbour (including the
2-rank case where both neighbours are the same peer), the momentum all-reduce between phases, the observable
all-reduce. The result must equal the single-process oracle run on all beads.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pimd_b_b200.config import SimConfig
from pimd_b_b200.distributed import ShardedSimulation, bead_range
from tests.helpers import FEMTOSECOND, KELVIN, MEV, Oracle, maxwell_momenta


class NumpyShard:
    """Test double of CudaShard: same protocol, CPU tensors, physics restated in NumPy (SoA [bead][axis][atom])."""

    def __init__(self, cfg, x, p, lo, hi):
        self.cfg, self.lo, self.hi = cfg, lo, hi
        n = hi - lo
        S = cfg.ndim * cfg.natoms
        self.xs = torch.zeros((n + 2) * S, dtype=torch.float64)          # halo | owned | halo
        self.x = self.xs.view(n + 2, cfg.ndim, cfg.natoms)
        self.x[1:n + 1] = torch.from_numpy(np.ascontiguousarray(np.transpose(x[lo:hi], (0, 2, 1))))
        self.p = torch.from_numpy(np.ascontiguousarray(np.transpose(p[lo:hi], (0, 2, 1)))).clone()
        self.f = torch.zeros_like(self.p)
        self.send_first = self.xs[S:2 * S]
        self.send_last = self.xs[n * S:(n + 1) * S]
        self.halo_before = self.xs[0:S]
        self.halo_after = self.xs[(n + 1) * S:(n + 2) * S]
        self.com = torch.zeros(4, dtype=torch.float64)
        self.k = cfg.spring_constant
        self.kext = cfg.mass * cfg.ext_omega ** 2

    def _sum_p(self):
        self.com[:self.cfg.ndim] = self.p.sum(dim=(0, 2))

    def _sub_com(self):
        self.p -= (self.com[:self.cfg.ndim] / (self.cfg.natoms * self.cfg.nbeads)).view(1, -1, 1)

    def step_phase(self, k):
        c = self.cfg
        n = self.hi - self.lo
        if k == 0:
            self._sum_p()
        elif k == 1:
            self._sub_com()
            self.p += 0.5 * c.dt * self.f
            self.x[1:n + 1] += c.dt * self.p / c.mass
        elif k == 2:
            xc = self.x[1:n + 1]
            self.f = self.k * (self.x[0:n] + self.x[2:n + 2] - 2 * xc) - self.kext * xc
            self.p += 0.5 * c.dt * self.f
            self._sum_p()
        elif k == 3:
            self._sub_com()

    def observables_partial(self):
        out = torch.zeros(10, dtype=torch.float64)
        out[6] = float((self.p ** 2).sum()) * 0.5 / self.cfg.mass      # cl_kinetic
        return out


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cfg_dict, x, p, steps, out_dir, halo="p2p"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = SimConfig(**cfg_dict)
    lo, hi = bead_range(cfg.nbeads, world, rank)
    shard = NumpyShard(cfg, x, p, lo, hi)
    sim = ShardedSimulation(cfg, shard, halo=halo)
    sim.exchange_halos()
    if halo == "allgather":        # also exercise the pending closing zeroMomentum: left open, subsumed, flushed
        sim.step(2, finalize=False)
        sim.step(steps - 3, finalize=False)
        sim.step(1, finalize=False)
        sim.flush()
    else:
        sim.step(steps)
    obs = sim.observables()
    n = hi - lo
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=shard.x[1:n + 1].numpy(), p=shard.p.numpy(),
             lo=lo, hi=hi, cl_kinetic=obs["cl_kinetic"])
    dist.barrier()
    dist.destroy_process_group()


def test_bead_range_partitions():
    for P in (1, 7, 8, 64, 129):
        for G in (1, 2, 3, 4, 8):
            if G > P:
                with pytest.raises(ValueError):
                    bead_range(P, G, 0)
                continue
            r = [bead_range(P, G, k) for k in range(G)]
            assert r[0][0] == 0 and r[-1][1] == P
            assert all(r[k][1] == r[k + 1][0] for k in range(G - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1


@pytest.mark.parametrize("world,nbeads,halo", [(2, 8, "p2p"), (3, 8, "p2p"), (2, 2, "p2p"), (2, 8, "allgather"),
                                               (3, 7, "allgather")])
def test_sharded_steps_equal_single_process_oracle(world, nbeads, halo, tmp_path):
    cfg = SimConfig(nbeads=nbeads, natoms=6, ndim=3, bosonic=False, fixcom=True, pbc=False,
                    temperature=5.802 * KELVIN, mass=1.0, size=300.0, interaction="free", external="harmonic",
                    ext_omega=3 * MEV, thermostat="none", seed=3, dt=FEMTOSECOND)
    rng = np.random.default_rng(5)
    x = rng.uniform(-20, 20, size=(nbeads, 6, 3))
    p = maxwell_momenta(cfg, rng) + 0.01
    steps = 7
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, cfg.as_dict(), x, p, steps, str(tmp_path), halo), nprocs=world,
                       join=True, start_method="spawn")
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.set("p", p)
    for _ in range(steps):
        orc.run_iteration()
    ref_x, ref_p = orc.get("x"), orc.get("p")
    ke = 0.0
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        lo, hi = int(d["lo"]), int(d["hi"])
        got_x = np.transpose(d["x"], (0, 2, 1))
        got_p = np.transpose(d["p"], (0, 2, 1))
        assert np.max(np.abs(got_x - ref_x[lo:hi])) < 1e-11 * np.max(np.abs(ref_x))
        assert np.max(np.abs(got_p - ref_p[lo:hi])) < 1e-11 * np.max(np.abs(ref_p))
        ke = float(d["cl_kinetic"])
    assert abs(ke - orc.observables()["cl_kinetic"]) < 1e-11 * ke
