"""Phase times of the native sharded driver (torchrun, one process per GPU): NL_SHARD_PROFILE=1 prints the device time of
prepare / exchange (partition, all-to-all, halo select, halo exchange) / build / count+fill on rank 0.
usage: torchrun --nproc-per-node G scripts/exp_shard_phases.py [atoms_per_gpu] [by-index|slabbed] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NL_SHARD_PROFILE", "1")
import numpy as np, torch, torch.distributed as dist
import neighbourlists_jl_b200 as nl
from importlib import import_module
from bench import DENSITY, CUTOFF, SEED

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "by-index"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sh = import_module("neighbourlists_jl_b200.sharded")
comm = sh.make_nccl_comm()
L = (n * world / DENSITY) ** (1.0 / 3.0)
C = np.eye(3) * L
rng = np.random.Generator(np.random.PCG64(SEED + rank))
X = rng.random((n, 3))
if mode == "slabbed":
    X[:, 2] = (X[:, 2] + rank) / world
X *= L
Xd = torch.from_numpy(X).cuda()
g = torch.arange(rank * n + 1, (rank + 1) * n + 1, dtype=torch.int32).cuda()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
for k in range(reps):
    ev[k].record()
    pl = sh.neighbour_list_sharded_native(Xd, g, CUTOFF, C, (True, True, True), comm, rank, world, with_R=True)
    del pl
ev[reps].record()
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("mode", mode, "world", world, "per-list ms (incl. profile syncs):", " ".join("%.2f" % ev[k].elapsed_time(ev[k + 1]) for k in range(reps)),
          "| NCCL env:", {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}, flush=True)
sh.shard_disconnect(comm)
nl._lib.check(nl._lib.lib().nl_nccl_comm_destroy(comm))
dist.destroy_process_group()
