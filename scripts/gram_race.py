"""Run-to-run bitwise comparison of the Gram kernel over shapes (see lincomb_race.py)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gcge_b200 import api
api.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65553
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rng = np.random.default_rng(2)
x = np.asfortranarray(rng.standard_normal((n, 484))); y = np.asfortranarray(rng.standard_normal((n, 230)))
X = api.MultiVec.from_numpy(x); Y = api.MultiVec.from_numpy(y)
for (xo, p, yo, q) in [(0, 480, 0, 40), (2, 440, 4, 40), (178, 302, 6, 222), (2, 40, 2, 40), (0, 64, 0, 30), (4, 480, 0, 8)]:
    ref = None; nbad = 0
    for r in range(reps):
        g = np.zeros((p, q), order="F")
        api.multivec_inner_prod("N", X, Y, (xo, yo), (xo + p, yo + q), g, p)
        if ref is None:
            ref = g.copy()
            want = x[:, xo:xo + p].T @ y[:, yo:yo + q]
            scale = np.abs(x[:, xo:xo + p]).T @ np.abs(y[:, yo:yo + q])
            print((xo, p, yo, q), "err vs numpy", float((np.abs(g - want) / scale).max()), flush=True)
        elif not np.array_equal(g, ref):
            nbad += 1
    print((xo, p, yo, q), "runs that differ from the first:", nbad, "of", reps - 1, flush=True)
