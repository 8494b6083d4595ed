#!/bin/bash
# compute-sanitizer over the hand-written kernels at small sizes (run under gpurun):
#   scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [what]
# what: dense (Gram + LinearComb shapes of the solve, TMA-fed and cp.async kernels), spmm (lattice, 1-D diagonal and CSR
# kernels incl. the fused dot), solve (a whole small GCG solve).  racecheck sees shared-memory hazards between
# threads of a CTA; hazards against the async (TMA) proxy are outside its model -- scripts/lincomb_race.py /
# scripts/gram_race.py (bitwise run-to-run comparison under load) are the tools for those.
tool=${1:-memcheck}; what=${2:-dense}
export PYTHONPATH=$(cd "$(dirname "$0")/.." && pwd)
case $what in
dense) prog='
import numpy as np
from gcge_b200 import api
api.init(0)
n = 4099
rng = np.random.default_rng(2)
x = np.asfortranarray(rng.standard_normal((n, 132))); y = np.asfortranarray(rng.standard_normal((n, 70)))
X = api.MultiVec.from_numpy(x); Y = api.MultiVec.from_numpy(y)
for (xo, p, yo, q) in [(0, 130, 0, 40), (2, 100, 4, 40), (1, 64, 3, 30), (4, 128, 0, 8), (0, 40, 0, 40)]:
    g = np.zeros((p, q), order="F")
    api.multivec_inner_prod("N", X, Y, (xo, yo), (xo + p, yo + q), g, p)
    assert np.abs(g - x[:, xo:xo + p].T @ y[:, yo:yo + q]).max() < 1e-9
    c = np.asfortranarray(rng.standard_normal((p, q)))
    api.multivec_linear_comb(X, Y, (xo, yo), (xo + p, yo + q), c, p, None, 0)
    assert np.abs(Y.numpy()[:, yo:yo + q] - x[:, xo:xo + p] @ c).max() < 1e-9
gs = np.zeros((40, 40), order="F")
api.multivec_inner_prod("S", X, X, (0, 0), (40, 40), gs, 40)
print("dense ok")
';;
spmm) prog='
import numpy as np
from gcge_b200 import api, problems as P
api.init(0)
for pen in (P.p1_fem_kuhn(14), P.laplace3d_7pt(12), P.q1_27pt(10), P.laplace1d_pencil(2001)):
    A = api.Mat(pen.A); n = pen.A.ncols
    for k in (3, 10, 40, 70):
        x = np.asfortranarray(np.random.default_rng(k).standard_normal((n, k)))
        X = api.MultiVec.from_numpy(x); Y = api.MultiVec(n, k)
        api.mat_dot_multivec(A, X, Y, (0, 0), (k, k))
        assert np.abs(Y.numpy() - pen.A.to_scipy() @ x).max() < 1e-9
    b = api.MultiVec.from_numpy(np.asfortranarray(np.random.default_rng(1).random((n, 10)))); xs = api.MultiVec(n, 10)
    api.block_pcg(A, b, xs, (0, 0), (10, 10), max_iter=5)
print("spmm ok")
';;
solve) prog='
from gcge_b200 import api, problems as P
api.init(0)
pen = P.p1_fem_kuhn(10)
o = api.gcg_solve(api.Mat(pen.A), api.Mat(pen.B), nev=6)
assert o["nev_conv"] >= 6
print("solve ok", o["num_iter"])
';;
esac
compute-sanitizer --tool $tool --error-exitcode 9 python -c "$prog"
echo "sanitizer $tool $what rc=$?"
