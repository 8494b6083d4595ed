"""CPU tests of the multi-GPU host logic: the 1-D row-block partition of a CCS
matrix, its halo / send lists and the exchange pattern.  The plan is built by the product's own
host code (b200_plan_*, the function b200_mat_create_from_ccs uses on every rank) -- no device is
needed.  The world-size-2 test runs the exchange over torch.distributed/gloo exactly as the GPU
path runs it over NCCL (pack rows -> send/recv per neighbour -> unpack behind the local rows)."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from gcge_b200 import api, problems as P
from oracle import gcg_numpy as G

ROOT = Path(__file__).resolve().parents[1]
ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))


def random_unsymmetric(n=57, density=0.08, seed=3):
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    m = sp.random(n, n, density=density, random_state=rng, format="csc") + sp.identity(n, format="csc") * 2.0
    m = m.tocsc(); m.sort_indices()
    return P.CCS(n, n, m.indptr.astype(np.int32), m.indices.astype(np.int32), m.data.astype(np.float64))


def oracle_spmm(M, x):
    y = np.zeros_like(x, order="F")
    G.clib().oracle_ccs_spmm(M.ncols, ip(M.j_col), ip(M.i_row), dp(M.data), dp(x), dp(y), x.shape[1])
    return y


def global_cols(p):
    """columns of the local CSR slab mapped back to global indices"""
    if p["halo_contiguous"] or not p["nhalo"]:
        return p["ci"] + p["row0"]
    return np.where(p["ci"] < p["nloc"], p["ci"] + p["row0"], p["halo_cols"][np.maximum(p["ci"] - p["nloc"], 0)])


def extended_x(p, xloc, halo_rows):
    """the x block a slab multiplies with, laid out as the device lays it out, plus the index shift:
    banded matrices keep the halo_below rows in FRONT of the local rows (one contiguous window
    of the global vector), otherwise all halo rows follow the local rows in halo-list order."""
    if p["halo_contiguous"]:
        hb = p["halo_below"]
        return np.vstack([halo_rows[:hb], xloc, halo_rows[hb:]]), hb
    return np.vstack([xloc, halo_rows]), 0


def local_spmm(plan, xext, shift=0):
    """Row-by-row gather in entry order with separate multiply and add -- the arithmetic of the
    device SpMM kernel (gcge_b200/csrc/b200_spmm.cu)."""
    y = np.zeros((plan["nloc"], xext.shape[1]))
    rp, ci, va = plan["rp"], plan["ci"], plan["va"]
    for r in range(plan["nloc"]):
        acc = np.zeros(xext.shape[1])
        for e in range(rp[r], rp[r + 1]):
            acc = acc + va[e] * xext[ci[e] + shift]
        y[r] = acc
    return y


CASES = [("7pt", lambda: P.laplace3d_7pt(6).A), ("p1", lambda: P.p1_fem_kuhn(5).B), ("27pt", lambda: P.q1_27pt(4).A),
         ("unsym", random_unsymmetric), ("1d", lambda: P.laplace1d_pencil(40).A)]


@pytest.mark.parametrize("name,make", CASES)
@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
def test_partition_plan_roundtrip_and_exchange_lists(name, make, nranks):
    M = make()
    n = M.ncols
    plans = [api.partition_plan(M, r, nranks) for r in range(nranks)]
    # row blocks tile [0, n) in rank order with the documented split
    assert plans[0]["row0"] == 0 and plans[-1]["row0"] + plans[-1]["nloc"] == n
    for r, p in enumerate(plans):
        assert p["row0"] == (n * r) // nranks and p["nloc"] == (n * (r + 1)) // nranks - (n * r) // nranks
    # bit-exact round trip: the slabs, columns mapped back, are the CSR image of the matrix
    csr = M.to_scipy().tocsr(); csr.sort_indices()
    rp_all, ci_all, va_all = [0], [], []
    for p in plans:
        rp_all.extend((p["rp"][1:] + rp_all[-1]).tolist()); ci_all.append(global_cols(p)); va_all.append(p["va"])
    assert np.array_equal(np.array(rp_all), csr.indptr)
    assert np.array_equal(np.concatenate(ci_all), csr.indices)
    assert np.array_equal(np.concatenate(va_all), csr.data)          # identical bits
    # what a sends to b is exactly what b expects from a, in the same order
    for a, pa in enumerate(plans):
        assert sorted(pa["nbr"]) == list(pa["nbr"]) and a not in pa["nbr"]
        for i, b in enumerate(pa["nbr"]):
            pb = plans[b]
            assert a in pb["nbr"], (a, b)
            j = list(pb["nbr"]).index(a)
            sent = pa["send_rows"][pa["send_off"][i]:pa["send_off"][i + 1]] + pa["row0"]
            expected = pb["halo_cols"][pb["recv_off"][j]:pb["recv_off"][j + 1]]
            assert np.array_equal(sent, expected)
    # emulated distributed SpMM == the oracle's global SpMM, bit for bit
    x = np.asfortranarray(np.random.default_rng(7).standard_normal((n, 3)))
    want = oracle_spmm(M, x)
    for p in plans:
        xext, shift = extended_x(p, x[p["row0"]:p["row0"] + p["nloc"]], x[p["halo_cols"]] if p["nhalo"] else x[:0])
        got = local_spmm(p, xext, shift)
        assert np.array_equal(got, want[p["row0"]:p["row0"] + p["nloc"]])
    # all ranks of one matrix use the same halo mode (the exchange would not match otherwise)
    assert len({p["halo_contiguous"] for p in plans}) == 1


def test_partition_rejects_rectangular_across_ranks():
    M = P.laplace1d_pencil(10).A
    rect = P.CCS(12, 10, M.j_col, M.i_row, M.data)
    with pytest.raises(api.B200Error, match="square"):
        api.partition_plan(rect, 0, 2)
    assert api.partition_plan(rect, 0, 1)["nloc"] == 12


def test_world_size_2_gloo_halo_exchange():
    """Two processes, gloo: each builds its own plan, packs the rows its neighbour needs, exchanges
    them with send/recv in the plan's neighbour order, multiplies its slab; rank 0 gathers and
    checks against the oracle.  Same control flow as b200k_spmm's NCCL path."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", PYTHONPATH=str(ROOT))
    procs = [subprocess.Popen([sys.executable, str(ROOT / "tests" / "dist_worker_cpu.py")],
                              env=dict(env, RANK=str(r), WORLD_SIZE="2"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
    assert "halo exchange ok" in outs[0]


@pytest.mark.parametrize("name,m", [("p1_fem_kuhn", 9), ("q1_27pt", 8), ("laplace3d_7pt", 10)])
def test_slab_generators_equal_the_whole_matrix_columns(name, m):
    """problems.pencil_rows (what a rank generates for itself with bench.py --local-gen) == the CCS columns of the
    whole-matrix generator for the same planes, bit for bit."""
    pen = getattr(P, name)(m)
    for (k0, k1) in ((0, m), (2, 5), (m - 1, m)):
        rowsA, rowsB = P.pencil_rows(name, m, k0, k1)
        c0, c1 = k0 * m * m, k1 * m * m
        for M, rows in ((pen.A, rowsA), (pen.B, rowsB)):
            if M is None:
                assert rows is None
                continue
            rp, ci, va = rows
            assert np.array_equal(rp, M.j_col[c0:c1 + 1] - M.j_col[c0])
            assert np.array_equal(ci, M.i_row[M.j_col[c0]:M.j_col[c1]])
            assert np.array_equal(va, M.data[M.j_col[c0]:M.j_col[c1]])
