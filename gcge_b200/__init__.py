"""gcge_b200 -- B200-native block GCG eigensolver hot path.

The product is the C-ABI shared library ``gcge_b200/lib/libgcge_b200.so`` (CUDA kernels
for sm_100a + host C drivers, ``include/gcge_b200.h``) and the OPS adaptor
``gcge_b200/app/app_b200.c``.  This Python package is only a ctypes harness over that ABI
for tests and ``bench.py``; it contains no numerical code and NO fallback: if the library
is missing, or no CUDA device is present, calls raise.
"""
from .api import (  # noqa: F401
    B200Error, GCGParams, Mat, MultiVec, block_pcg, dense_syev, device_count, gcg_solve, init,
    kernel_launches, lib, lib_path, libc_srand, orth, sync, wtime,
)
