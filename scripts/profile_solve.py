"""Phase timing of the device GCG on the BASELINE configs (development aid)."""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gcge_b200 import api, problems as P

ap = argparse.ArgumentParser()
ap.add_argument("--gen", default="laplace3d_7pt"); ap.add_argument("--m", type=int, default=100)
ap.add_argument("--nev", type=int, default=50); ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--max-iter", type=int, default=0); ap.add_argument("--verbose", type=int, default=0)
a = ap.parse_args()
api.init(0)
t = time.time(); pen = getattr(P, a.gen)(a.m); tg = time.time() - t
t = time.time(); A = api.Mat(pen.A); B = None if pen.B is None else api.Mat(pen.B); tu = time.time() - t
print(f"{a.gen} m={a.m} n={pen.A.ncols} nnz={pen.A.nnz} gen {tg:.1f}s upload {tu:.1f}s", flush=True)
for r in range(a.reps):
    t = time.time()
    o = api.gcg_solve(A, B, nev=a.nev, max_iter=a.max_iter, verbose=bool(a.verbose))
    dt = time.time() - t
    st = o["stats"]
    print(json.dumps({"wall": round(dt, 3), "num_iter": o["num_iter"], "nev_conv": o["nev_conv"],
                      **{k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()}}), flush=True)
    print("eval[:3]", o["eval"][:3], flush=True)
    o["evec_mv"].close()
