"""Bead sharding across the GPUs of one box: one process per GPU, torch.distributed for the plumbing.

Two ways to couple the shards:

* ``PeerShardedSimulation`` (the product path): the ranks exchange one 256-byte blob each at start-up
  (``gather_blobs``: cudaIpc handles of the coordinate array and of a small mailbox), after which halo slices and
  momentum sums are *stored into the peers' memory by the step's own kernels* over NVLink (csrc/integrator.cu,
  include/pimdb200.h ``pimdb_peer_attach``). A step is one CUDA-graph replay per rank with no host call or collective
  inside; torch.distributed is only used for the start-up gather and for the all-reduce of the observable partials on
  logging steps.
* ``ShardedSimulation`` (host-driven, kept as the portable cross-check): the host runs NCCL / gloo collectives between
  the four phases of ``pimdb_step_phase``. This is the choreography described next, and what the gloo tests exercise.

The reference runs one MPI rank per bead and moves, every step, one bead slice to each ring neighbour
(MPI_Sendrecv, src/simulation.cpp:299-347), NDIM doubles for zeroMomentum (MPI_Allreduce, :595) and one double
per observable column (src/observables/observable.cpp:105). Here rank r owns the contiguous bead range
``bead_range(P, G, r)``; the same three exchanges become
  * a ring halo exchange of the first / last owned bead slices (NCCL point-to-point over NVLink),
  * an all-reduce of the ndim partial momentum sums (padded to 4 doubles),
  * an all-reduce of the 10 observable partials on logging steps.
Pair forces, springs of interior beads and the integrator never leave the GPU; the exchange recursion runs
redundantly on the two ranks that own bead 1 and bead P (exactly what ranks 0 and P-1 do in the reference).

``ShardedSimulation`` is written against a small *shard* protocol so that the collective choreography can be
tested on CPU with the gloo backend (tests/test_sharding_gloo.py) and runs unchanged on NCCL:

    shard.step_phase(k)            k = 0..3, see include/pimdb200.h
    shard.send_first / send_last   tensors holding the first / last owned bead slice
    shard.halo_before / halo_after tensors receiving the neighbours' slices
    shard.com                      tensor (4,) of partial momentum sums
    shard.observables_partial()    tensor (10,) float64
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

from ._cabi import OBS_FIELDS


def bead_range(nbeads: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced bead ranges (the first ``nbeads % world`` ranks own one extra bead)."""
    if world > nbeads:
        raise ValueError(f"cannot shard {nbeads} beads over {world} ranks")
    base, rem = divmod(nbeads, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DevicePtrView:
    """Zero-copy view of device memory owned by libpimdb200.so (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {
            "shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


class CudaShard:
    """The product shard: a DeviceSim on this rank's GPU; its tensors alias the library's device buffers."""

    def __init__(self, cfg, rank: int, world: int, device: int):
        from .engine import DeviceSim
        lo, hi = bead_range(cfg.nbeads, world, rank)
        torch.cuda.set_device(device)
        self.stream = torch.cuda.Stream(device=device)
        self.sim = DeviceSim(cfg, lo, hi, device)
        self.sim.set_stream(self.stream.cuda_stream)
        dev = torch.device("cuda", device)

        def view(which):
            ptr, cnt = self.sim.halo_ptr(which)
            return torch.as_tensor(_DevicePtrView(ptr, cnt), device=dev)

        self.send_first, self.send_last = view(0), view(1)
        self.halo_before, self.halo_after = view(2), view(3)
        self.com = torch.as_tensor(_DevicePtrView(self.sim.com_ptr(), 4), device=dev)
        self.device = dev

    def step_phase(self, k: int):
        self.sim.step_phase(k)

    def observables_partial(self) -> torch.Tensor:
        o = self.sim.observables()
        return torch.tensor([o[n] for n in OBS_FIELDS], dtype=torch.float64, device=self.device)

    def refresh_local_forces(self):
        self.sim.update_forces()


class ShardedSimulation:
    """Runs Simulation::run's loop body (src/simulation.cpp:246-259) over bead shards."""

    def __init__(self, cfg, shard, group=None, halo: str = "p2p"):
        self.cfg = cfg
        self.shard = shard
        self.group = group
        if halo not in ("p2p", "allgather"):
            raise ValueError("halo must be 'p2p' or 'allgather'")
        self.halo = halo
        self._gather_send = self._gather_recv = None
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.prev = (self.rank - 1) % self.world
        self.next = (self.rank + 1) % self.world

    # Simulation::updateNeighboringCoordinates across ranks
    def exchange_halos(self):
        s = self.shard
        if self.world == 1:
            s.halo_before.copy_(s.send_last)
            s.halo_after.copy_(s.send_first)
            return
        if self.halo == "allgather":
            # One collective instead of four point-to-point operations: every rank contributes its first and last bead
            # slice and picks its two neighbours' out of the result. More bytes than the ring needs (2 G slices), but
            # they are tiny, and a collective can be captured into a CUDA graph together with the kernels of the step.
            n = s.send_first.numel()
            if self._gather_send is None:
                self._gather_send = torch.empty(2 * n, dtype=s.send_first.dtype, device=s.send_first.device)
                self._gather_recv = torch.empty(self.world * 2 * n, dtype=s.send_first.dtype, device=s.send_first.device)
            self._gather_send[:n].copy_(s.send_first)
            self._gather_send[n:].copy_(s.send_last)
            dist.all_gather_into_tensor(self._gather_recv, self._gather_send, group=self.group)
            s.halo_before.copy_(self._gather_recv[(2 * self.prev + 1) * n:(2 * self.prev + 2) * n])   # prev's last slice
            s.halo_after.copy_(self._gather_recv[2 * self.next * n:(2 * self.next + 1) * n])          # next's first slice
            return
        # Order matters when prev == next (two ranks): the k-th send to a peer pairs with its k-th receive.
        ops = [
            dist.P2POp(dist.isend, s.send_first, self.prev, self.group),
            dist.P2POp(dist.isend, s.send_last, self.next, self.group),
            dist.P2POp(dist.irecv, s.halo_after, self.next, self.group),
            dist.P2POp(dist.irecv, s.halo_before, self.prev, self.group),
        ]
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def _allreduce_com(self):
        if self.cfg.fixcom and self.world > 1:
            dist.all_reduce(self.shard.com, op=dist.ReduceOp.SUM, group=self.group)

    # ---- CUDA-graph replay of one sharded step (kernels of all four phases + the NCCL halo / all-reduce calls) ----
    def enable_graph(self) -> bool:
        """Capture one step into a torch CUDA graph (NCCL collectives are capturable); returns False and stays on
        the eager path if capture is not possible in this environment. Call from inside the shard's stream context."""
        if getattr(self, "_graph", None) is not None:
            return True
        if not torch.cuda.is_available() or not hasattr(self.shard, "stream"):
            return False
        self.halo = "allgather"                       # point-to-point NCCL calls inside a capture hung on this stack
        try:
            # warm-up (NCCL communicators, lazy allocations) must not advance the trajectory: put the state back afterwards
            sim = getattr(self.shard, "sim", None)
            saved = None if sim is None else (sim.get("x"), sim.get("p"), sim.get("f"))
            self._step_eager(2)
            torch.cuda.synchronize()
            if saved is not None:
                sim.set("x", saved[0]); sim.set("p", saved[1]); sim.set("f", saved[2])
                self.exchange_halos()
                torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.shard.stream, capture_error_mode="thread_local"):
                self._step_eager(1)
            self._graph = g
            return True
        except Exception as exc:   # pragma: no cover - depends on the NCCL / driver combination
            self._graph = None
            self._graph_error = repr(exc)
            torch.cuda.synchronize()
            return False

    def step(self, nsteps: int = 1, finalize: bool = True):
        """``finalize=False`` leaves the closing zeroMomentum of the last iteration pending (see ``_step_eager``); it is
        carried out by ``flush()``, which ``observables()`` calls, or made redundant by the next ``step``."""
        g = getattr(self, "_graph", None)
        if g is not None:
            for _ in range(nsteps):
                g.replay()
            return
        self._step_eager(nsteps, finalize)

    def _step_eager(self, nsteps: int = 1, finalize: bool = True):
        """One iteration of Simulation::run is  O, Z, B, A, forces, B, O, Z  (O: thermostat half step, Z: zeroMomentum).
        Z is the projection p -> p - mean(p) and O is affine with the same coefficients for every degree of freedom
        (p -> c1 p + c2 xi), so  Z O Z = Z O:  the closing Z of an iteration is subsumed by the first Z of the next one,
        and between consecutive iterations its all-reduce and its kernel are skipped (the positions are updated after
        a Z either way). Only the last iteration before the momenta are looked at carries it out (``flush``). Other
        thermostats do not commute with Z like that, so they keep both."""
        s = self.shard
        skip_closing = self.cfg.fixcom and self.cfg.thermostat in ("langevin", "none") and not self.cfg.nmthermostat
        for it in range(nsteps):
            self._com_pending = False    # a pending closing Z is subsumed by this iteration's first Z
            s.step_phase(0)          # thermostat half step (+ local momentum sums)
            self._allreduce_com()
            s.step_phase(1)          # COM removal, B, A
            self.exchange_halos()
            s.step_phase(2)          # forces, B, thermostat half step (+ local momentum sums)
            if skip_closing and (it + 1 < nsteps or not finalize):
                self._com_pending = True
                continue
            self._allreduce_com()
            s.step_phase(3)          # COM removal

    def flush(self):
        """Carry out a closing zeroMomentum that ``step(..., finalize=False)`` left pending."""
        if getattr(self, "_com_pending", False):
            self._allreduce_com()
            self.shard.step_phase(3)
            self._com_pending = False

    def observables(self) -> dict:
        self.flush()
        part = self.shard.observables_partial()
        if self.world > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        return {n: float(v) for n, v in zip(OBS_FIELDS, part.tolist())}


# ---------------------------------------------------------------------------------------------------------------------
def gather_blobs(blob: bytes, group=None) -> bytes:
    """All-gather one fixed-size byte record per rank, in rank order (the start-up exchange of the peer-memory path).
    Works on any backend: the staging tensor lives where the backend wants it (CUDA for nccl, host for gloo)."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    out = torch.empty(world * len(blob), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return bytes(out.cpu().numpy().tobytes())


class PeerShardedSimulation:
    """Simulation::run's loop body (src/simulation.cpp:246-259) over bead shards coupled through peer memory.

    ``make_sim(lo, hi)`` builds this rank's handle (default: a DeviceSim on ``device``); the handle must offer
    ``peer_export() -> bytes`` and ``peer_attach(world, rank, blobs)`` -- which is all this class needs from it at
    start-up, so the choreography is testable on CPU with a stand-in (tests/test_sharding_gloo.py)."""

    def __init__(self, cfg, rank: int = None, world: int = None, device: int = 0, group=None, make_sim=None):
        self.cfg = cfg
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.lo, self.hi = bead_range(cfg.nbeads, self.world, self.rank)
        if make_sim is None:
            from .engine import DeviceSim
            torch.cuda.set_device(device)
            self.stream = torch.cuda.Stream(device=device)
            self.sim = DeviceSim(cfg, self.lo, self.hi, device)
            self.sim.set_stream(self.stream.cuda_stream)
            self.device = torch.device("cuda", device)
        else:
            self.sim = make_sim(self.lo, self.hi)
            self.stream = None
            self.device = torch.device("cpu")
        blobs = gather_blobs(self.sim.peer_export(), group)
        self.sim.peer_attach(self.world, self.rank, blobs)
        dist.barrier(group)     # every rank is attached (and has pushed its halo slices) before anyone steps

    def set_state(self, x=None, p=None):
        """Global arrays [P][N][D]; this rank uploads its own beads. Collective (the coordinates' halo slices move)."""
        self.sim.upload(None if x is None else x[self.lo:self.hi], None if p is None else p[self.lo:self.hi])

    def step(self, nsteps: int = 1):
        self.sim.step(nsteps)

    def observables(self) -> dict:
        """Sum of the per-rank partial structs, what ObservablesLogger::log does with MPI_Allreduce
        (src/observables/observable.cpp:92-116). Collective."""
        o = self.sim.observables()
        part = torch.tensor([o[n] for n in OBS_FIELDS], dtype=torch.float64, device=self.device)
        if self.world > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        return {n: float(v) for n, v in zip(OBS_FIELDS, part.tolist())}
