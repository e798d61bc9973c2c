"""Worker of tests/test_gpu_peer.py: one process per bead shard (torchrun), coupled through peer memory.

    python -m torch.distributed.run --nproc-per-node G tests/peer_worker.py <case> <nsteps> <out.npz>

Every rank runs its shard with PeerShardedSimulation; rank 0 gathers x, p, f of all shards and writes them (plus the
all-reduced observables) for the parent test to compare with a single-handle run. PIMDB_TEST_SAME_GPU=1 maps every rank
to cuda:0 (the cudaIpc path between processes on a one-GPU box).
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    case, nsteps, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    from tests.peer_cases import make_case
    from pimd_b_b200.distributed import PeerShardedSimulation

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    same = os.environ.get("PIMDB_TEST_SAME_GPU") == "1"
    dev = 0 if same else int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")      # start-up plumbing only; the step itself uses no collective
    cfg, x, p = make_case(case)
    ps = PeerShardedSimulation(cfg, rank, world, dev)
    ps.set_state(x, p)
    ps.sim.update_forces()
    f0 = ps.sim.get("f")
    ps.step(nsteps)
    obs = ps.observables()
    xs, pp, ff = ps.sim.get("x"), ps.sim.get("p"), ps.sim.get("f")
    ps.sim.synchronize()
    parts = [None] * world
    dist.all_gather_object(parts, (ps.lo, ps.hi, xs, pp, ff, f0))
    if rank == 0:
        parts.sort(key=lambda t: t[0])
        np.savez(out, x=np.concatenate([t[2] for t in parts]), p=np.concatenate([t[3] for t in parts]),
                 f=np.concatenate([t[4] for t in parts]), f0=np.concatenate([t[5] for t in parts]),
                 obs=np.array([obs[k] for k in sorted(obs)]), obs_keys=np.array(sorted(obs)))
    dist.barrier()
    ps.sim.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
