"""MultiVecSetRandomValue on the device: time of the fill at the headline shape (n = 8 M, 400 columns) and at one
block (40 columns); checks the first and last values against glibc's rand() stream replayed on the host."""
import ctypes as C, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gcge_b200 import api

api.init(0)
libc = C.CDLL("libc.so.6")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
for k in (400, 40):
    X = api.MultiVec(n, k)
    api.libc_srand(0); X.set_random(0, k); api.sync()
    ts = []
    for _ in range(3):
        api.libc_srand(0)
        api.timer_start(); X.set_random(0, k); ts.append(api.timer_stop())
    head = X.numpy(0, 1)[:5, 0]; 
    libc.srand(0); want = np.array([libc.rand() / 2147483648.0 for _ in range(5)])
    print(f"n={n} k={k}: {min(ts):.2f} ms  ({8.0 * n * k / min(ts) / 1e6:.0f} GB/s written)  first values match glibc: {np.array_equal(head, want)}", flush=True)
    X.close()
