"""BASELINE config 5: kernel sweep.  SpMM (Y = A X), Gram (X^T Y, 3k x k), LinearComb (X C), axpby, dots and
the B-orthogonalisation of a k-block against a 2k-block (TestOrth's shape) at k = 16 ... 512 on an n-row P1-FEM pencil, timed with CUDA
events on the library stream (L2 flushed between repetitions), reported as achieved HBM GB/s
or FP64 TFLOP/s from the ALGORITHMIC bytes/flops of SURVEY.md §8d."""
import argparse, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gcge_b200 import api, problems as P


def sweep(a):
    """a: namespace with m, gen, ks, reps, ops, p, noflush; returns the header dict and the list of rows (also printed)."""
    rows = []
    say = (lambda *x, **kw: None) if getattr(a, "quiet", False) else print
    api.init(0)
    pen = getattr(P, a.gen)(a.m)
    n, nnz = pen.A.ncols, pen.A.nnz
    A = api.Mat(pen.A)
    say(f"# {a.gen} m={a.m} n={n} nnz={nnz}", flush=True)
    peaks = {}
    try:
        peaks = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)

    def timeit(fn, reps):
        fn(); api.sync()
        ts = []
        for _ in range(reps):
            if not a.noflush:
                api.flush_l2()
            api.timer_start(); fn(); ts.append(api.timer_stop())
        return float(np.median(ts)), float(np.min(ts))

    ops = a.ops.split(",")
    for k in [int(v) for v in a.ks.split(",")]:
        p = a.p if a.p else 3 * k
        X = api.MultiVec(n, max(p, k)); Y = api.MultiVec(n, k)
        api.libc_srand(1)
        X.upload(np.asfortranarray(np.random.default_rng(0).random((n, 4))), 0)   # cheap fill: replicate 4 columns
        for c in range(4, max(p, k), 4):
            w = min(4, max(p, k) - c)
            api.multivec_axpby(1.0, X, 0.0, X, (0, c), (w, c + w))
        row = {"k": k, "p": p}
        if "spmm" in ops:
            med, best = timeit(lambda: api.mat_dot_multivec(A, X, Y, (0, 0), (k, k)), a.reps)
            byt = nnz * 12 + (n + 1) * 4 + 16 * n * k
            row["spmm_ms"] = round(med, 4); row["spmm_GBs"] = round(byt / med / 1e6, 1); row["spmm_frac"] = round(byt / med / 1e6 / hbm, 3)
        if "axpby" in ops:
            med, best = timeit(lambda: api.multivec_axpby(0.5, X, 1.5, Y, (0, 0), (k, k)), a.reps)
            byt = 24 * n * k
            row["axpby_ms"] = round(med, 4); row["axpby_GBs"] = round(byt / med / 1e6, 1); row["axpby_frac"] = round(byt / med / 1e6 / hbm, 3)
        if "dots" in ops:
            d = np.zeros(k)
            med, best = timeit(lambda: api.multivec_inner_prod("D", X, Y, (0, 0), (k, k), d, 1), a.reps)
            byt = 16 * n * k
            row["dots_ms"] = round(med, 4); row["dots_GBs"] = round(byt / med / 1e6, 1); row["dots_frac"] = round(byt / med / 1e6 / hbm, 3)
        if "gram" in ops:
            g = np.zeros((p, k), order="F")
            med, best = timeit(lambda: api.multivec_inner_prod("N", X, Y, (0, 0), (p, k), g, p), a.reps)
            fl = 2.0 * n * p * k; byt = 8 * n * (p + k)
            row["gram_ms"] = round(med, 4); row["gram_TF"] = round(fl / med / 1e9, 2); row["gram_GBs"] = round(byt / med / 1e6, 1)
        if "lincomb" in ops:
            coef = np.asfortranarray(np.random.default_rng(1).random((p, k)))
            med, best = timeit(lambda: api.multivec_linear_comb(X, Y, (0, 0), (p, k), coef, p, None, 0), a.reps)
            fl = 2.0 * n * p * k; byt = 8 * n * (p + k)
            row["lincomb_ms"] = round(med, 4); row["lincomb_TF"] = round(fl / med / 1e9, 2); row["lincomb_GBs"] = round(byt / med / 1e6, 1)
        if "orth" in ops and p >= 3 * k:
            # B-orthonormalise the k columns [2k, 3k) against the 2k columns in front of them (reference TestOrth,
            # test/test_orth.c:44-111, at scale): 2k-block prepared once, the k-block refilled before every repetition
            Bm = api.Mat(pen.B) if pen.B is not None else None
            ws = api.MultiVec(n, min(k, 80))
            api.libc_srand(3); X.set_random(0, 3 * k)
            end0 = api.orth(X, 0, 2 * k, B=Bm, ws=ws, block_size=80)
            ts = []
            for _ in range(max(2, a.reps // 2)):
                X.set_random(2 * k, 3 * k)
                api.sync()
                api.timer_start(); end = api.orth(X, 2 * k, 3 * k, B=Bm, ws=ws, block_size=80); ts.append(api.timer_stop())
            med = float(np.median(ts))
            nb = (k + 79) // 80                                        # blocks of 80 columns, two rounds each
            # algorithmic traffic: per block and round one B x (SpMM bytes), Gram + update against everything
            # in front (read X0 twice, read + write the block), panel Gram + update
            byt = 0.0; fl = 0.0
            for b in range(nb):
                kb = min(80, k - 80 * b); m0 = 2 * k + 80 * b
                spmm_b = (pen.B.nnz * 12 + (n + 1) * 4 + 16 * n * kb) if pen.B is not None else 0
                byt += 2 * (2 * spmm_b + 8 * n * (2 * m0 + 3 * kb) + 8 * n * 3 * kb)
                fl += 2 * (4.0 * n * m0 * kb + 4.0 * n * kb * kb)
            row["orth_ms"] = round(med, 3); row["orth_end"] = [int(end0), int(end)]
            row["orth_GBs"] = round(byt / med / 1e6, 1); row["orth_TF"] = round(fl / med / 1e9, 2)
            ws.close()
            if Bm is not None:
                Bm.close()
        say(json.dumps(row), flush=True)
        rows.append(row)
        X.close(); Y.close()
    A.close()
    return {"gen": a.gen, "m": a.m, "n": n, "nnz": nnz, "hbm_gbs": hbm}, rows


def arguments(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=200)
    ap.add_argument("--gen", default="p1_fem_kuhn")
    ap.add_argument("--ks", default="16,32,64,128,256,512")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ops", default="spmm,axpby,gram,lincomb,dots,orth")
    ap.add_argument("--p", type=int, default=0, help="Gram/LinearComb inner width (default 3k)")
    ap.add_argument("--noflush", type=int, default=0)
    return ap.parse_args(argv)


if __name__ == "__main__":
    sweep(arguments())
