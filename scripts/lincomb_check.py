import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from gcge_b200 import api
api.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5001
for (p, q) in [(302, 222), (304, 222), (302, 224), (302, 30), (302, 62), (288, 30), (16, 30), (32, 30), (64, 30), (80, 30), (302, 14), (302, 46), (302, 16), (302, 32)]:
    rng = np.random.default_rng(p * 1000 + q)
    x = np.asfortranarray(rng.standard_normal((n, 480))); y = np.asfortranarray(np.zeros((n, 400)))
    coef = np.asfortranarray(rng.standard_normal((p, q)))
    X = api.MultiVec.from_numpy(x); Y = api.MultiVec.from_numpy(y)
    api.multivec_linear_comb(X, Y, (0, 0), (p, q), coef, p, None, 0)
    got = Y.numpy()[:, :q]
    want = x[:, :p] @ coef
    err = np.abs(got - want)
    bad_cols = np.nonzero(err.max(axis=0) > 1e-9)[0]
    print((p, q), "max err", float(err.max()), "bad cols", bad_cols[:12].tolist(), "n bad rows", int((err.max(axis=1) > 1e-9).sum()), flush=True)
