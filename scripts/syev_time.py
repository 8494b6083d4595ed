"""Times the device eigen-solve of a projected-matrix-shaped problem (order n) with CUDA events."""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gcge_b200 import api
api.init(0)
for n in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "120,240,480").split(",")]:
    rng = np.random.default_rng(n)
    nx = (n * 5) // 6
    a = np.zeros((n, n)); a[np.arange(nx), np.arange(nx)] = np.sort(rng.uniform(30, 650, nx))
    bw = n - nx
    e = rng.standard_normal((nx, bw)) * 5
    c = rng.standard_normal((bw, bw)); c = c @ c.T * 50 + np.eye(bw) * 700
    a[:nx, nx:] = e; a[nx:, :nx] = e.T; a[nx:, nx:] = c
    ts = []
    for rep in range(3):
        api.prof_enable(True)            # resets the per-class counters
        w, z, sw = api.dense_syev(a)
        api.sync()
        ts.append(api.prof_report()["syev_jacobi"]["ms"])
        api.prof_enable(False)
    wl = np.linalg.eigvalsh(a)
    print({"n": n, "sweeps": sw, "ms": [round(t, 3) for t in ts], "err": float(np.abs(w - wl).max() / np.abs(wl).max())}, flush=True)
