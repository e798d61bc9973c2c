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


# ---------------------------------------------------------------------------------------------------------------------
# Peer-memory path (PeerShardedSimulation): the start-up exchange of blobs over gloo, and the device protocol itself --
# sequence-numbered momentum sums in two slots, "done reading" credits before a halo slice is overwritten, halo flags --
# restated in NumPy over memory-mapped files that stand in for the peers' device memory. Same kernels-in-order structure
# as csrc/integrator.cu (k_integrate prologue / epilogue), so an index mix-up or a hand-shake that can dead-lock shows
# up here, on CPU, with world sizes 2 (both neighbours are the same peer) and 3.
class PeerNumpySim:
    BLOB = 256

    def __init__(self, cfg, lo, hi, tmpdir, rank):
        self.cfg, self.lo, self.hi, self.rank = cfg, lo, hi, rank
        self.n = hi - lo
        self.S = cfg.ndim * cfg.natoms
        self.path = os.path.join(tmpdir, f"peer{rank}.bin")
        self.words = 2 * 8 * 4 + 2 * 8 + 4            # com values | com sequence numbers | halo_flag[2], credit[2]
        total = self.words + (self.n + 2) * self.S
        np.zeros(total, dtype=np.float64).tofile(self.path)
        self.mem = np.memmap(self.path, dtype=np.float64, mode="r+")
        self.x = self.mem[self.words:].reshape(self.n + 2, cfg.ndim, cfg.natoms)
        self.p = np.zeros((self.n, cfg.ndim, cfg.natoms))
        self.f = np.zeros_like(self.p)
        self.seq_com = 0
        self.seq_halo = 0
        self.z_owed = False
        self.k = cfg.spring_constant
        self.kext = cfg.mass * cfg.ext_omega ** 2

    # mailbox views of any rank's mapping
    @staticmethod
    def _views(mem):
        com = mem[:64].reshape(2, 8, 4)
        seq = mem[64:80].reshape(2, 8)
        flags = mem[80:82]
        credit = mem[82:84]
        return com, seq, flags, credit

    def peer_export(self) -> bytes:
        rec = f"{self.path}|{self.lo}|{self.hi}".encode()
        return rec + b"\0" * (self.BLOB - len(rec))

    def peer_attach(self, world, rank, blobs):
        assert len(blobs) == world * self.BLOB and rank == self.rank
        self.world = world
        self.prev, self.next = (rank - 1) % world, (rank + 1) % world
        self.box, self.ranges = [], []
        for r in range(world):
            path, lo, hi = blobs[r * self.BLOB:(r + 1) * self.BLOB].rstrip(b"\0").decode().split("|")
            self.ranges.append((int(lo), int(hi)))
            self.box.append(np.memmap(path, dtype=np.float64, mode="r+"))
        assert self.ranges[rank] == (self.lo, self.hi)
        self._push_halos()

    def _wait(self, cond):
        import time
        t0 = time.time()
        while not cond():
            assert time.time() - t0 < 30.0, "peer wait timed out"
            time.sleep(0.0005)

    def _push_halos(self, first=None, last=None):
        """first / last: what to send instead of the current first / last owned bead (the early push of a step)."""
        first = self.x[1] if first is None else first
        last = self.x[self.n] if last is None else last
        k = self.seq_halo + 1
        _, _, _, credit_next = self._views(self.box[self.next])
        _, _, _, credit_prev = self._views(self.box[self.prev])
        credit_next[0] = k            # I am next's previous rank
        credit_prev[1] = k            # I am prev's next rank
        _, _, _, mine = self._views(self.mem)
        self._wait(lambda: mine[0] >= k and mine[1] >= k)
        nprev = self.ranges[self.prev][1] - self.ranges[self.prev][0]
        xprev = self.box[self.prev][self.words:].reshape(nprev + 2, self.cfg.ndim, self.cfg.natoms)
        nnext = self.ranges[self.next][1] - self.ranges[self.next][0]
        xnext = self.box[self.next][self.words:].reshape(nnext + 2, self.cfg.ndim, self.cfg.natoms)
        xprev[nprev + 1] = first              # my first bead -> prev's trailing halo
        xnext[0] = last                       # my last bead  -> next's leading halo
        self._views(self.box[self.prev])[2][1] = k
        self._views(self.box[self.next])[2][0] = k
        self.seq_halo = k

    def _wait_halos(self):
        flags = self._views(self.mem)[2]
        self._wait(lambda: flags[0] >= self.seq_halo and flags[1] >= self.seq_halo)

    def _sum_push(self):
        s = self.seq_com + 1
        sums = self.p.sum(axis=(0, 2))
        for r in range(self.world):
            com, seq, _, _ = self._views(self.box[r])
            com[s & 1, self.rank, :self.cfg.ndim] = sums
            seq[s & 1, self.rank] = s
        self.seq_com = s

    def _subcm_wait(self):
        s = self.seq_com
        com, seq, _, _ = self._views(self.mem)
        self._wait(lambda: all(seq[s & 1, r] == s for r in range(self.world)))
        tot = np.zeros(self.cfg.ndim)
        for r in range(self.world):
            tot += com[s & 1, r, :self.cfg.ndim]
        shift = (tot / (self.cfg.natoms * self.cfg.nbeads)).reshape(1, -1, 1)
        self.p -= shift
        return shift

    def upload(self, x, p):
        self.x[1:self.n + 1] = np.transpose(x, (0, 2, 1))
        self.p[...] = np.transpose(p, (0, 2, 1))
        self.z_owed = False
        self._push_halos()

    def step(self, nsteps):
        c = self.cfg
        for _ in range(nsteps):
            self.z_owed = False                  # subsumed by this iteration's first zeroMomentum
            # kernel 1: momentum sums -> every rank; the boundary beads' coming positions, without the (unknown) shift,
            # -> the ring neighbours:  x~ = x + dt/m (p + dt/2 f)
            self._sum_push()
            xt = lambda b: self.x[1 + b] + c.dt / c.mass * (self.p[b] + 0.5 * c.dt * self.f[b])
            self._push_halos(first=xt(0), last=xt(self.n - 1))
            # kernel 2: wait for the sums, COM removal, B, A, and the uniform shift taken off the two received slices
            shift = self._subcm_wait()
            self.p += 0.5 * c.dt * self.f
            self.x[1:self.n + 1] += c.dt * self.p / c.mass
            self._wait_halos()
            self.x[0] -= c.dt / c.mass * shift[0]
            self.x[self.n + 1] -= c.dt / c.mass * shift[0]
            xc = self.x[1:self.n + 1]
            self.f = self.k * (self.x[0:self.n] + self.x[2:self.n + 2] - 2 * xc) - self.kext * xc
            self.p += 0.5 * c.dt * self.f
            self.z_owed = True

    def observables(self):
        if self.z_owed:
            self._sum_push()
            self._subcm_wait()
            self.z_owed = False
        from pimd_b_b200._cabi import OBS_FIELDS
        out = {n: 0.0 for n in OBS_FIELDS}
        out["cl_kinetic"] = float((self.p ** 2).sum()) * 0.5 / self.cfg.mass
        return out


def _peer_worker(rank, world, port, cfg_dict, x, p, steps, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pimd_b_b200.distributed import PeerShardedSimulation, gather_blobs
    cfg = SimConfig(**cfg_dict)
    # the start-up gather keeps rank order and record size
    rec = bytes([rank]) * 256
    allrec = gather_blobs(rec)
    assert len(allrec) == world * 256 and all(allrec[256 * r] == r and allrec[256 * r + 255] == r for r in range(world))
    sim = PeerShardedSimulation(cfg, make_sim=lambda lo, hi: PeerNumpySim(cfg, lo, hi, out_dir, rank))
    sim.set_state(x, p)
    sim.step(3)
    mid = sim.observables()          # collective: carries out the closing zeroMomentum
    sim.step(steps - 3)
    obs = sim.observables()
    s = sim.sim
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=np.array(s.x[1:s.n + 1]), p=s.p, lo=s.lo, hi=s.hi,
             cl_kinetic=obs["cl_kinetic"], mid=mid["cl_kinetic"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nbeads", [(2, 8), (3, 8), (2, 2), (3, 7)])
def test_peer_protocol_equals_single_process_oracle(world, nbeads, tmp_path):
    cfg = SimConfig(nbeads=nbeads, natoms=6, ndim=3, bosonic=False, fixcom=True, pbc=False,
                    temperature=5.802 * KELVIN, mass=1.0, size=300.0, interaction="free", external="harmonic",
                    ext_omega=3 * MEV, thermostat="none", seed=3, dt=FEMTOSECOND)
    rng = np.random.default_rng(11)
    x = rng.uniform(-20, 20, size=(nbeads, 6, 3))
    p = maxwell_momenta(cfg, rng) + 0.01
    steps = 8
    port = _free_port()
    mp.start_processes(_peer_worker, args=(world, port, cfg.as_dict(), x, p, steps, str(tmp_path)), nprocs=world,
                       join=True, start_method="spawn")
    orc = Oracle(cfg)
    orc.set("x", x)
    orc.set("p", p)
    for _ in range(steps):
        orc.run_iteration()
    ref_x, ref_p = orc.get("x"), orc.get("p")
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        lo, hi = int(d["lo"]), int(d["hi"])
        assert np.max(np.abs(np.transpose(d["x"], (0, 2, 1)) - ref_x[lo:hi])) < 1e-11 * np.max(np.abs(ref_x))
        assert np.max(np.abs(np.transpose(d["p"], (0, 2, 1)) - ref_p[lo:hi])) < 1e-11 * np.max(np.abs(ref_p))
        assert abs(float(d["cl_kinetic"]) - orc.observables()["cl_kinetic"]) < 1e-11 * float(d["cl_kinetic"])
