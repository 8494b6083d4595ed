"""Multi-GPU parity worker (one process per GPU, launched by torchrun from
tests/test_dist_gpu.py or by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_worker_gpu.py
Every rank builds the SAME global inputs, the library keeps its row block; results are compared
with the oracle (plain-C SpMM, numpy, the reference's golden eigenvalues)."""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from gcge_b200 import api, problems as P          # noqa: E402
from oracle import gcg_numpy as G                  # noqa: E402
from test_partition import oracle_spmm             # noqa: E402

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
api.init(local)
rank, world = api.comm_init_from_torch()
L = api.lib()


def gather_rows(mv, k0, k1, n):
    """global (n x k) block from every rank's slab"""
    full = np.zeros((n, k1 - k0), order="F")
    api._chk(L.b200_mv_download(mv.h, k0, k1, full.ctypes.data_as(api.c_dbl_p), n))   # fills own rows only
    t = torch.from_numpy(np.ascontiguousarray(full)).cuda()
    dist.all_reduce(t)                                                                 # other rows are zero
    return t.cpu().numpy()


def check(cond, msg):
    if not cond:
        raise AssertionError(f"rank {rank}: {msg}")


# ---- partition round trip from the device --------------------------------------------------
pen = P.p1_fem_kuhn(9)
n = pen.A.ncols
A, B = api.Mat(pen.A), api.Mat(pen.B)
row0, nloc, nnzl, nhalo = (C.c_int(), C.c_int(), C.c_int(), C.c_int())
L.b200_mat_local_range.argtypes = [C.c_void_p] + [api.c_int_p] * 4
api._chk(L.b200_mat_local_range(A.h, C.byref(row0), C.byref(nloc), C.byref(nnzl), C.byref(nhalo)))
rp = np.zeros(nloc.value + 1, np.int32); ci = np.zeros(max(nnzl.value, 1), np.int32); va = np.zeros(max(nnzl.value, 1))
L.b200_mat_local_csr.argtypes = [C.c_void_p, api.c_int_p, api.c_int_p, api.c_dbl_p]
api._chk(L.b200_mat_local_csr(A.h, api._ip(rp), api._ip(ci), api._dp(va)))
csr = pen.A.to_scipy().tocsr(); csr.sort_indices()
lo, hi = row0.value, row0.value + nloc.value
check(lo == (n * rank) // world and hi == (n * (rank + 1)) // world, "row range")
check(np.array_equal(rp, csr.indptr[lo:hi + 1] - csr.indptr[lo]), "local row pointers")
check(np.array_equal(ci[:nnzl.value], csr.indices[csr.indptr[lo]:csr.indptr[hi]]), "local columns (global ids)")
check(np.array_equal(va[:nnzl.value], csr.data[csr.indptr[lo]:csr.indptr[hi]]), "local values")

# ---- SpMM with halo exchange: bit-exact against the oracle -----------------------------------
for k in (1, 6, 10, 40, 70):
    x = np.asfortranarray(np.random.default_rng(k).standard_normal((n, k + 2)))
    X = api.MultiVec.from_numpy(x); Y = api.MultiVec(n, k)
    api.mat_dot_multivec(A, X, Y, (1, 0), (1 + k, k))
    got = gather_rows(Y, 0, k, n)
    want = oracle_spmm(pen.A, np.asfortranarray(x[:, 1:1 + k]))
    check(np.array_equal(got, want), f"SpMM k={k} differs from the oracle")

# ---- lattice kernel on plane-aligned slabs: variable coefficients, halo planes from the neighbours ------------
from test_slots_gpu import _lattice_matrix         # noqa: E402
for kind, dims in (("p1", (12, 9, 4 * world)), ("27pt", (10, 6, 3 * world)), ("7pt", (16, 5, 5 * world))):
    M = _lattice_matrix(*dims, kind, seed=7)
    Am = api.Mat(M)
    st = Am.storage()
    check(st["lat_s1"] == dims[0] and st["lat_s2"] == dims[0] * dims[1], f"lattice not recognised on a slab: {st}")
    nl = M.ncols
    for k in (10, 40):
        xl = np.asfortranarray(np.random.default_rng(k).standard_normal((nl, k + 2)))
        Xl = api.MultiVec.from_numpy(xl); Yl = api.MultiVec(nl, k)
        for rep in range(2):                          # second pass: stale halo rows of the first must be replaced
            if rep:
                xl = np.asfortranarray(np.random.default_rng(100 + k).standard_normal((nl, k + 2))); Xl.upload(xl)
            api.mat_dot_multivec(Am, Xl, Yl, (2, 0), (2 + k, k))
            got = gather_rows(Yl, 0, k, nl)
            check(np.array_equal(got, oracle_spmm(M, np.asfortranarray(xl[:, 2:2 + k]))), f"lattice SpMM {kind} k={k} rep={rep}")

# ---- slab-local construction: every rank hands over only its own planes (b200_mat_create_from_local_rows) ------
mloc = 4 * world
penl = P.p1_fem_kuhn(mloc)
nl = mloc ** 3
rowsA, rowsB = P.pencil_rows("p1_fem_kuhn", mloc, (mloc // world) * rank, (mloc // world) * (rank + 1))
Awhole = api.Mat(penl.A); Bwhole = api.Mat(penl.B)
Aloc = api.Mat.from_local_rows(nl, (mloc // world) * rank * mloc * mloc, *rowsA)
Bloc = api.Mat.from_local_rows(nl, (mloc // world) * rank * mloc * mloc, *rowsB)
check(Aloc.storage() == Awhole.storage(), f"storage differs: {Aloc.storage()} vs {Awhole.storage()}")
xl = np.asfortranarray(np.random.default_rng(9).standard_normal((nl, 16)))
Xl = api.MultiVec.from_numpy(xl); Y1 = api.MultiVec(nl, 16); Y2 = api.MultiVec(nl, 16)
api.mat_dot_multivec(Awhole, Xl, Y1, (0, 0), (16, 16)); api.mat_dot_multivec(Aloc, Xl, Y2, (0, 0), (16, 16))
g1 = gather_rows(Y1, 0, 16, nl)
check(np.array_equal(g1, gather_rows(Y2, 0, 16, nl)) and np.array_equal(g1, oracle_spmm(penl.A, xl)), "local-rows SpMM")
o1 = api.gcg_solve(Awhole, Bwhole, nev=8); o2 = api.gcg_solve(Aloc, Bloc, nev=8)
check(o1["num_iter"] == o2["num_iter"] and np.array_equal(o1["eval"], o2["eval"]), "local-rows solve differs from whole-CCS solve")

# malformed rows on ONE rank fail on EVERY rank (the construction is collective: nobody may be left waiting), and the
# library keeps working afterwards
rp_b, ci_b, va_b = (np.array(a, copy=True) for a in rowsA)
if rank == world - 1:
    r_bad = int(np.argmax(np.diff(rp_b) >= 2)); e = int(rp_b[r_bad] - rp_b[0])
    ci_b[e], ci_b[e + 1] = ci_b[e + 1], ci_b[e]                 # columns of one row no longer ascend
failed = False
try:
    api.Mat.from_local_rows(nl, (mloc // world) * rank * mloc * mloc, rp_b, ci_b, va_b)
except RuntimeError as exc:
    failed = True
    check(("ascending" in str(exc)) == (rank == world - 1) and ("another rank" in str(exc)) == (rank != world - 1), f"message: {exc}")
check(failed, "malformed rows on the last rank were accepted")
Aagain = api.Mat.from_local_rows(nl, (mloc // world) * rank * mloc * mloc, *rowsA)
api.mat_dot_multivec(Aagain, Xl, Y2, (0, 0), (16, 16))
check(np.array_equal(g1, gather_rows(Y2, 0, 16, nl)), "local-rows SpMM after a rejected construction")

# ---- the first slab rows lack the farthest sub-diagonal, so the halo plan's extent (from the entries
# those rows have) is SHORTER than the reach of the diagonal image; rows further inside still need halo rows and
# must not be multiplied before the halo has arrived ------------------------------------------------------------
import scipy.sparse as sp                             # noqa: E402
mm = 16
full = P.laplace3d_7pt(mm).A.to_scipy().tolil()
nn = mm ** 3
for q in range(1, world):
    r0 = (nn * q) // world
    for r in range(r0, r0 + 100):                     # rows r0 .. r0+99 lose their (r, r - m^2) entry
        if r - mm * mm >= 0:
            full[r, r - mm * mm] = 0.0
fullc = full.tocsc(); fullc.eliminate_zeros(); fullc.sort_indices()
Mh = P.CCS(nn, nn, fullc.indptr.astype(np.int32), fullc.indices.astype(np.int32), fullc.data.astype(np.float64))
Ah = api.Mat(Mh)
for k in (40,):
    Xh = api.MultiVec(nn, k); Yh = api.MultiVec(nn, k)
    for rep in range(3):
        xh = np.asfortranarray(np.random.default_rng(50 + rep).standard_normal((nn, k)))
        Xh.upload(xh)
        api.mat_dot_multivec(Ah, Xh, Yh, (0, 0), (k, k))
        check(np.array_equal(gather_rows(Yh, 0, k, nn), oracle_spmm(Mh, xh)), f"masked-halo SpMM rep={rep}")

# ---- Gram / dots: globally reduced, identical on every rank -----------------------------------
x = np.asfortranarray(np.random.default_rng(1).standard_normal((n, 24)))
y = np.asfortranarray(np.random.default_rng(2).standard_normal((n, 10)))
X = api.MultiVec.from_numpy(x); Y = api.MultiVec.from_numpy(y)
g = np.zeros((24, 10), order="F")
api.multivec_inner_prod("N", X, Y, (0, 0), (24, 10), g, 24)
check(np.abs(g - x.T @ y).max() < 1e-11 * np.abs(x.T @ y).max(), "Gram block")
t = torch.from_numpy(np.ascontiguousarray(g)).cuda(); t2 = t.clone(); dist.broadcast(t2, 0)
check(bool(torch.equal(t, t2)), "Gram block is not bit-identical across ranks")
d = np.zeros(10)
api.multivec_inner_prod("D", X, Y, (3, 0), (13, 10), d, 1)
check(np.abs(d - np.einsum("ij,ij->j", x[:, 3:13], y)).max() < 1e-11 * n, "column dots")
q = np.zeros((10, 10), order="F"); ws = api.MultiVec(n, 10)
api.multivec_qtap("S", "N", Y, B, Y, (0, 0), (10, 10), q, 10, ws)
Bd = pen.B.to_scipy()
check(np.abs(q - y.T @ (Bd @ y)).max() < 1e-11 * np.abs(q).max(), "QtAP")

# ---- RNG: every rank keeps its rows of the one global glibc stream --------------------------------
R = api.MultiVec(n, 5)
api.libc_srand(0); R.set_random(1, 4)
want = np.zeros((n, 5), order="F"); G.srand(0); G.fill_random(want, 1, 4)
check(np.array_equal(gather_rows(R, 0, 5, n), want), "set_random differs from the glibc stream")

# ---- fused providers ------------------------------------------------------------------------------
xs = np.asfortranarray(np.random.default_rng(5).random((n, 12)))
V = api.MultiVec.from_numpy(xs)
end = api.orth(V, 0, 12, B=B, block_size=8)
v = gather_rows(V, 0, 12, n)
check(end == 12 and np.abs(v.T @ (Bd @ v) - np.eye(12)).max() < 1e-12, "B-orthonormalisation")
Ad = pen.A.to_scipy()
sol = np.asfortranarray(np.random.default_rng(6).random((n, 4))); rhs = np.asfortranarray(Ad @ sol)
Xs = api.MultiVec(n, 4)
api.block_pcg(A, api.MultiVec.from_numpy(rhs), Xs, (0, 0), (4, 4), max_iter=600, rate=1e-30, tol=1e-10)
check(np.abs(Ad @ gather_rows(Xs, 0, 4, n) - rhs).max() < 1e-8, "BlockPCG residual")

# ---- whole solve against the reference's golden output ------------------------------------------------
gold = json.loads((ROOT / "tests" / "golden" / "gcg_reference.json").read_text())["cases"]
for idx in (4, 2):
    case = gold[idx]
    pc = getattr(P, case["generator"])(**case["args"])
    Am = api.Mat(pc.A); Bm = None if pc.B is None else api.Mat(pc.B)
    o = api.gcg_solve(Am, Bm, nev=case["nev"])
    k = min(o["nev_conv"], case["nev_conv"])
    err = np.max(np.abs(o["eval"][:k] - np.array(case["eval"][:k])) / np.abs(np.array(case["eval"][:k])))
    check(o["nev_conv"] >= case["nev"], "not converged")
    check(err < 1e-10, f"eigenvalues differ from the reference: {err:.2e}")
    check(abs(o["num_iter"] - case["num_iter"]) <= 2, f"iterations {o['num_iter']} vs reference {case['num_iter']}")
    vec = gather_rows(o["evec_mv"], 0, k, pc.A.ncols)
    Ax = pc.A.to_scipy() @ vec; Bx = vec if pc.B is None else pc.B.to_scipy() @ vec
    res = np.linalg.norm(Ax - Bx * o["eval"][:k], axis=0)
    check(np.all(res <= 1e-1) and np.all(res <= np.abs(o["eval"][:k]) * 1e-8 * 1.0000001), "residual test")
    ev = torch.from_numpy(o["eval"].copy()).cuda(); ev0 = ev.clone(); dist.broadcast(ev0, 0)
    check(bool(torch.equal(ev, ev0)), "eigenvalues are not bit-identical across ranks")

dist.barrier()
if rank == 0:
    print(f"dist gpu ok: world={world} launches={api.kernel_launches()}")
api.comm_finalize()
dist.destroy_process_group()
