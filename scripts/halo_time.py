"""Times the SpMM halo exchange between row-block ranks (torchrun, one rank per GPU): a loop of k-column
SpMMs on the P1-FEM pencil, per-class device times; 'gap before axpby' is the neighbour exchange."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch, torch.distributed as dist
from gcge_b200 import api, problems as P
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
api.init(local)
rank, world = api.comm_init_from_torch()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 200
k = int(sys.argv[2]) if len(sys.argv) > 2 else 40
reps = 50
pen = P.p1_fem_kuhn(m)
A = api.Mat(pen.A)
n = pen.A.ncols
X = api.MultiVec(n, k); Y = api.MultiVec(n, k)
api.libc_srand(1); X.set_random(0, k)
for _ in range(3):
    api.mat_dot_multivec(A, X, Y, (0, 0), (k, k))
api.sync(); dist.barrier()
api.prof_enable(True)
api.timer_start()
for _ in range(reps):
    api.mat_dot_multivec(A, X, Y, (0, 0), (k, k))
ms = api.timer_stop()
pr = api.prof_report(); api.prof_enable(False)
if rank == 0:
    tot_gap = sum(v["gap_before_ms"] for v in pr.values())
    print({"world": world, "m": m, "k": k, "per_spmm_ms": round(ms / reps, 4), "spmm_kernel_ms": round(pr["spmm"]["ms"] / reps, 4),
           "axpby_ms": round(pr["axpby"]["ms"] / reps, 4), "gaps_ms": round(tot_gap / reps, 4),
           "env": {e: os.environ[e] for e in os.environ if e.startswith("NCCL_") or e.startswith("B200_")}}, flush=True)
api.comm_finalize()
dist.destroy_process_group()
