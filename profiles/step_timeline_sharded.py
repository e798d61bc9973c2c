"""In-kernel %globaltimer timeline of ONE captured MD step on every rank of a bead-sharded run (peer-memory path).
    python -m torch.distributed.run --nproc-per-node G profiles/step_timeline_sharded.py [c3]
Ranks that own an exterior bead have the timeline slots (PIMDB_TIMELINE needs the exchange buffers); the others print
nothing. Times are relative to the rank's own first kernel of the step."""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["PIMDB_TIMELINE"] = "1"
import torch, torch.distributed as dist
from pimd_b_b200 import workloads as wl
from pimd_b_b200.distributed import PeerShardedSimulation

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = wl.config(name)
x, p = wl.initial_state(cfg, name)
ps = PeerShardedSimulation(cfg, rank, world, lr)
ps.set_state(x, p)
sim = ps.sim
with torch.cuda.stream(ps.stream):
    ps.step(50); sim.synchronize(); dist.barrier()
    fn = sim.lib.pimdb_debug_timeline; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
    buf = (C.c_ulonglong * 64)()
    fn(sim.h, buf)
    for rep in range(2):
        dist.barrier()
        ps.step(1); sim.synchronize()
        n = fn(sim.h, buf)
        dist.barrier()
        for r in range(world):
            if r == rank and n > 0:
                t = np.array(buf[:], dtype=np.uint64).reshape(32, 2)[:n].astype(np.int64)
                t0 = t[:, 0].min()
                print(f"rank {rank} step {rep}: kernels in launch order [start us, end us, duration us]")
                for i, (a, b) in enumerate(t):
                    print(f"   #{i}: {(a - t0) / 1e3:8.2f} {(b - t0) / 1e3:8.2f} {(b - a) / 1e3:8.2f}", flush=True)
                fs = sim.lib.pimdb_debug_integrate_stamps; fs.restype = C.c_int; fs.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
                sb = (C.c_ulonglong * 64)()
                k = fs(sim.h, sb)
                st = np.array(sb[:], dtype=np.uint64).reshape(8, 8).astype(np.int64)
                for i in range(min(k, 8)):
                    print(f"   k_integrate launch {i}: phase stamps (us after the step's first kernel)",
                          " ".join(f"{(v - t0) / 1e3:7.2f}" if v else "      -" for v in st[i]), flush=True)
            dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record(ps.stream); ps.step(500); e1.record(ps.stream); torch.cuda.synchronize()
    print(f"rank {rank}: {e0.elapsed_time(e1) / 500 * 1e3:.1f} us per step back to back", flush=True)
dist.barrier()
sim.close()
dist.destroy_process_group()
