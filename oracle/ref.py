"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/_ref/libgcge_ref.so.

The library is the UNMODIFIED reference (CCS + OpenMP path) compiled by
oracle/Makefile from /root/reference, plus the thin entry points of
oracle/ref_driver.c.  Nothing in the product (gcge_b200/) may import this
module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs do.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_ref" / "libgcge_ref.so"
_lib = None

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)


def available() -> bool:
    return _LIB_PATH.exists()


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(f"{_LIB_PATH} missing: run `make -C oracle` where /root/reference exists")
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")  # reference calls BLAS inside omp regions
        # OpenBLAS (opencv wheel) needs its bundled libquadmath/libgfortran; RUNPATH of our
        # library does not cover its dependencies, so preload them globally.
        import glob
        import sysconfig

        blasdir = Path(sysconfig.get_paths()["purelib"]) / "opencv_python_headless.libs"
        for pat in ("libquadmath-*.so*", "libgfortran-*.so*", "libopenblasp-*.so"):
            for f in sorted(glob.glob(str(blasdir / pat))):
                C.CDLL(f, mode=C.RTLD_GLOBAL)
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.ref_set_threads.restype = C.c_int
        _lib.ref_get_max_threads.restype = C.c_int
    return _lib


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dbl_p)


def _se(start, end):
    s = (C.c_int * 2)(*start)
    e = (C.c_int * 2)(*end)
    return s, e


def set_threads(n: int) -> int:
    return lib().ref_set_threads(int(n))


def max_threads() -> int:
    return lib().ref_get_max_threads()


def srand(seed: int = 0):
    lib().ref_srand(C.c_uint(seed))


def _F(a):
    """Column-major float64 view required by LAPACKVEC (ldd == nrows)."""
    assert a.dtype == np.float64 and a.flags.f_contiguous, "multivectors must be Fortran-ordered float64"
    return a


def mat_dot_multivec(M, x, y, start, end):
    s, e = _se(start, end)
    n = x.shape[0]
    if M is None:
        lib().ref_mat_dot_multivec(n, None, None, None, _dp(_F(x)), x.shape[1], _dp(_F(y)), y.shape[1], s, e)
    else:
        lib().ref_mat_dot_multivec(n, _ip(M.j_col), _ip(M.i_row), _dp(M.data), _dp(_F(x)), x.shape[1],
                                   _dp(_F(y)), y.shape[1], s, e)


def multivec_axpby(alpha, x, beta, y, start, end):
    s, e = _se(start, end)
    n = y.shape[0]
    lib().ref_multivec_axpby(n, C.c_double(alpha), _dp(None if x is None else _F(x)),
                             0 if x is None else x.shape[1], C.c_double(beta), _dp(_F(y)), y.shape[1], s, e)


def multivec_linear_comb(x, y, start, end, coef, ldc, beta, incb):
    s, e = _se(start, end)
    n = y.shape[0]
    lib().ref_multivec_linear_comb(n, _dp(None if x is None else _F(x)), 0 if x is None else x.shape[1],
                                   _dp(_F(y)), y.shape[1], s, e, _dp(coef), int(ldc), _dp(beta), int(incb))


def multivec_inner_prod(nsd, x, y, start, end, ip, ld):
    s, e = _se(start, end)
    lib().ref_multivec_inner_prod(x.shape[0], C.c_char(nsd.encode()), _dp(_F(x)), x.shape[1],
                                  _dp(_F(y)), y.shape[1], s, e, _dp(ip), int(ld))


def multivec_qtap(ntsA, ntsdQAP, q, M, p, start, end, qAp, ld, ws):
    s, e = _se(start, end)
    lib().ref_multivec_qtap(q.shape[0], C.c_char(ntsA.encode()), C.c_char(ntsdQAP.encode()),
                            _dp(_F(q)), q.shape[1],
                            _ip(None if M is None else M.j_col), _ip(None if M is None else M.i_row),
                            _dp(None if M is None else M.data),
                            _dp(_F(p)), p.shape[1], s, e, _dp(qAp), int(ld), _dp(_F(ws)), ws.shape[1])


def multivec_set_random(x, start, end):
    lib().ref_multivec_set_random(x.shape[0], _dp(_F(x)), x.shape[1], int(start), int(end))


def multivec_orth(x, start_x, end_x, B=None, method="mgs", block_size=-1, max_reorth=2,
                  orth_zero_tol=2 * np.finfo(float).eps):
    n, nc = x.shape
    ws = np.zeros((n, nc), order="F")
    dbl_ws = np.zeros(4 * nc * nc + 16 * nc + 64)
    end = C.c_int(end_x)
    lib().ref_multivec_orth(n, 1 if method == "bgs" else 0, int(block_size), int(max_reorth),
                            C.c_double(orth_zero_tol), _dp(_F(x)), nc, int(start_x), C.byref(end),
                            _ip(None if B is None else B.j_col), _ip(None if B is None else B.i_row),
                            _dp(None if B is None else B.data), _dp(ws), nc, _dp(dbl_ws))
    return end.value


def block_pcg(M, b, x, start, end, max_iter=30, rate=1e-2, tol=1e-14, tol_type="abs"):
    s, e = _se(start, end)
    n = x.shape[0]
    k = end[0] - start[0]
    r = np.zeros((n, k), order="F"); p = np.zeros((n, k), order="F"); w = np.zeros((n, k), order="F")
    niter = C.c_int(0); res = C.c_double(0)
    lib().ref_block_pcg(n, _ip(M.j_col), _ip(M.i_row), _dp(M.data), _dp(_F(b)), b.shape[1],
                        _dp(_F(x)), x.shape[1], s, e, int(max_iter), C.c_double(rate), C.c_double(tol),
                        tol_type.encode(), _dp(r), _dp(p), _dp(w), k, C.byref(niter), C.byref(res))
    return niter.value, res.value


def gcg_solve(A, B=None, nev=10, nev_max=0, block_size=0, nev_init=0, tol=(1e-1, 1e-8), max_iter=500,
              argv=(), quiet=True, want_evec=True, evec_given=None):
    """Runs the reference GCG (reference src/ops_eig_sol_gcg.c:1253) through the sequence of
    reference test/test_eig_sol_gcg.c:28-169.  Returns dict(eval, evec, num_iter, nev_conv, seconds)."""
    n = A.ncols
    nev_max_eff = nev_max if nev_max > 0 else 2 * nev
    ev = np.zeros(nev_max_eff)
    evec = np.zeros((n, nev_max_eff), order="F") if want_evec else None
    num_iter = C.c_int(0); nev_conv = C.c_int(0); secs = C.c_double(0)
    args = [b"ref"] + [str(a).encode() for a in argv]
    argv_c = (C.c_char_p * len(args))(*args)
    lib().ref_gcg_solve(n, _ip(A.j_col), _ip(A.i_row), _dp(A.data),
                        _ip(None if B is None else B.j_col), _ip(None if B is None else B.i_row),
                        _dp(None if B is None else B.data),
                        int(nev), int(nev_max), int(block_size), int(nev_init),
                        C.c_double(tol[0]), C.c_double(tol[1]), int(max_iter),
                        len(args), argv_c, 1 if quiet else 0,
                        _dp(ev), _dp(evec), C.byref(num_iter), C.byref(nev_conv), C.byref(secs),
                        0 if evec_given is None else int(evec_given.shape[1]),
                        _dp(None if evec_given is None else _F(np.asfortranarray(evec_given, dtype=np.float64))))
    return {"eval": ev, "evec": evec, "num_iter": num_iter.value, "nev_conv": nev_conv.value,
            "seconds": secs.value}


def block_amg_dense(A_levels, P_levels, b, x, max_iter, rate, tol):
    """The reference's BlockAMG (src/ops_lin_sol.c:466-715) over its dense LAPACK back end (the CCS back end
    cannot run it: no MultiGridCreate, transposed multiply assumes symmetry).  A_levels[l]: dense n_l x n_l,
    P_levels[l]: dense n_l x n_{l+1}; x is updated in place."""
    L = len(A_levels)
    A = [np.asfortranarray(a, dtype=np.float64) for a in A_levels]
    P = [np.asfortranarray(p, dtype=np.float64) for p in P_levels]
    n_l = (C.c_int * L)(*[a.shape[0] for a in A])
    Ap = (c_dbl_p * L)(*[_dp(a) for a in A])
    Pp = (c_dbl_p * L)(*([_dp(p) for p in P] + [None]))
    mi = (C.c_int * len(max_iter))(*max_iter); ra = (C.c_double * len(rate))(*rate); to = (C.c_double * len(tol))(*tol)
    lib().ref_block_amg_dense(L, n_l, Ap, Pp, int(x.shape[1]), _dp(_F(b)), _dp(_F(x)), mi, ra, to)
    return x
