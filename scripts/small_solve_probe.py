"""Repeated device solves of a small pencil (default P1 m = 40, nev = 200): device time, launches and the phase table of
every call -- the same_size leg of bench.py in isolation."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gcge_b200 import api, problems as P

m = int(sys.argv[1]) if len(sys.argv) > 1 else 40
nev = int(sys.argv[2]) if len(sys.argv) > 2 else 200
api.init(0)
pen = P.p1_fem_kuhn(m)
A, B = api.Mat(pen.A), api.Mat(pen.B)
ev = api.MultiVec(pen.A.ncols, 2 * nev)
for i in range(4):
    w0 = time.time()
    api.timer_start()
    o = api.gcg_solve(A, B, nev=nev, evec=ev, seed=0)
    ms = api.timer_stop()
    print(i, f"{ms:.1f} ms device, {time.time() - w0:.3f} s wall", o["num_iter"], o["nev_conv"], {k: (round(v, 3) if isinstance(v, float) else v) for k, v in o["stats"].items()}, flush=True)
