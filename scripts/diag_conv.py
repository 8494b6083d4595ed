import sys, os, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gcge_b200 import api, problems as P
api.init(0)
m = int(sys.argv[1]); nev = int(sys.argv[2]); cap = int(sys.argv[3]) if len(sys.argv) > 3 else 80
pen = P.p1_fem_kuhn(m)
A, B = api.Mat(pen.A), api.Mat(pen.B)
t = time.time()
o = api.gcg_solve(A, B, nev=nev, numIterMax=cap)
print({"m": m, "nev": nev, "num_iter": o["num_iter"], "nev_conv": o["nev_conv"], "eval0": float(o["eval"][0]), "eval_last": float(o["eval"][nev - 1]),
       "s": round(time.time() - t, 2), "env": {e: os.environ[e] for e in os.environ if e.startswith("B200_")}}, flush=True)
