import ctypes as C, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gcge_b200 import api
api.init(0)
t = C.c_double(0)
assert api.lib().b200_measure_dmma_peak(C.byref(t)) == 0
print("DMMA register-resident peak TFLOP/s:", round(t.value, 2))
out = (C.c_double * 5)()
assert api.lib().b200_measure_fp64_peaks(out) == 0
print("FP64 pipes TFLOP/s: dmma_shared_operands %.2f  dmma_2Ax8B_interleaved %.2f  dmma_2Ax8B_rowwise %.2f  dfma %.2f  dmma+dfma %.2f" % tuple(out))
import torch
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
torch.matmul(a, b); torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("cuBLAS DGEMM 8192^3 TFLOP/s:", round(2 * 8192**3 / best / 1e9, 2))
