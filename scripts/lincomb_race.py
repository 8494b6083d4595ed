import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gcge_b200 import api
api.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rng = np.random.default_rng(1)
x = np.asfortranarray(rng.standard_normal((n, 480)))
X = api.MultiVec.from_numpy(x)
import ast
shapes = ast.literal_eval(sys.argv[3]) if len(sys.argv) > 3 else [(120, 360, 280), (178, 302, 222), (0, 480, 400), (40, 440, 360)]
for (off, p, q) in shapes:
    coef = np.asfortranarray(rng.standard_normal((p, q)))
    Y = api.MultiVec(n, 400)
    ref = None; nbad = 0; worst = 0.0
    for r in range(reps):
        api.multivec_linear_comb(X, Y, (off, off), (off + p, off + q), coef, p, None, 0)
        got = Y.numpy()[:, off:off + q]
        if ref is None:
            ref = got.copy()
            want = x[:2048, off:off + p] @ coef
            print((off, p, q), "first-run err vs numpy (2048 rows)", float(np.abs(got[:2048] - want).max()), flush=True)
        else:
            d = np.abs(got - ref)
            if d.max() > 0:
                nbad += 1; worst = max(worst, float(d.max()))
                if nbad <= 2:
                    idx = np.argwhere(d > 0)
                    rows = np.unique(idx[:, 0]); cols = np.unique(idx[:, 1])
                    print("   run", r, "differs in", len(idx), "entries; rows mod 128:", sorted(set((rows % 128).tolist()))[:20], "n rows", len(rows),
                          "tiles", sorted(set((rows // 128).tolist()))[:8], "cols", cols.min(), "..", cols.max(), flush=True)
    print((off, p, q), "runs that differ from the first:", nbad, "of", reps - 1, "worst", worst, flush=True)
