"""Worker of tests/test_partition.py::test_world_size_2_gloo_halo_exchange (one process per rank,
torch.distributed gloo on CPU).  TEST INFRASTRUCTURE: emulates the device SpMM with numpy on the
plan the product's host code built; the exchange follows the plan exactly like the NCCL path."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from gcge_b200 import api, problems as P          # noqa: E402
sys.path.insert(0, str(ROOT / "tests"))
from test_partition import extended_x, local_spmm, oracle_spmm          # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
for M in (P.p1_fem_kuhn(6).A, P.laplace3d_7pt(7).A):
    n, k = M.ncols, 4
    p = api.partition_plan(M, rank, world)
    x = np.asfortranarray(np.random.default_rng(11).standard_normal((n, k)))      # same on both ranks
    xloc = x[p["row0"]:p["row0"] + p["nloc"]].copy()
    halo = np.zeros((p["nhalo"], k))
    reqs, recvs = [], []
    for i, q in enumerate(p["nbr"]):
        rows = p["send_rows"][p["send_off"][i]:p["send_off"][i + 1]]
        sbuf = torch.from_numpy(np.ascontiguousarray(xloc[rows]))                  # pack
        rbuf = torch.empty((p["recv_off"][i + 1] - p["recv_off"][i], k), dtype=torch.float64)
        reqs.append(dist.isend(sbuf, int(q))); reqs.append(dist.irecv(rbuf, int(q)))
        recvs.append((i, rbuf))
    for r in reqs:
        r.wait()
    for i, rbuf in recvs:                                                          # unpack in halo-list order
        halo[p["recv_off"][i]:p["recv_off"][i + 1]] = rbuf.numpy()
    xext, shift = extended_x(p, xloc, halo)
    y = local_spmm(p, xext, shift)
    # Gram block: local partial + allreduce (the NCCL path does the same on the device)
    gl = torch.from_numpy(xloc.T @ y); dist.all_reduce(gl)
    parts = [None] * world
    dist.all_gather_object(parts, (p["row0"], y))
    if rank == 0:
        want = oracle_spmm(M, x)
        got = np.vstack([b for _, b in sorted(parts, key=lambda t: t[0])])
        assert np.array_equal(got, want), "distributed SpMM differs from the oracle"
        assert np.allclose(gl.numpy(), x.T @ want, rtol=1e-12, atol=1e-12)
dist.barrier()
if rank == 0:
    print("halo exchange ok")
dist.destroy_process_group()
