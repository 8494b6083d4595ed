#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu7.log
tail -12 gpurun_out/pytest_gpu7.log
python - <<'PY' 2>&1 | tee gpurun_out/matbuild_time.log
import time, os, numpy as np
from gcge_b200 import api, problems as P
api.init(0)
pen = P.p1_fem_kuhn(200)
for h in (pen.A.j_col, pen.A.i_row, pen.A.data): api.host_register(h)
for mode in ("device", "host", "device"):
    if mode == "host": os.environ["B200_HOST_BUILD"] = "1"
    else: os.environ.pop("B200_HOST_BUILD", None)
    t = time.time(); A = api.Mat(pen.A); api.sync(); print(mode, "build of A (nnz %d): %.3f s" % (pen.A.nnz, time.time() - t), flush=True)
    A.close()
PY
