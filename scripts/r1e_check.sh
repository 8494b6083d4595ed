#!/bin/bash
# round-1 session-3 check: GPU tests, syev timing, SpMM variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
tail -15 gpurun_out/pytest_gpu3.log
python scripts/syev_time.py 120,240,480 2>&1 | tee gpurun_out/syev_time.log
for v in 0 1 2 3 4 5; do
  echo "variant $v"; B200_SPMM_VARIANT=$v python scripts/kernel_sweep.py --m 100 --ops spmm --ks 40 --reps 5
done 2>&1 | tee gpurun_out/spmm_variants.log
B200_SPMM_OLD_DIA=1 python scripts/kernel_sweep.py --m 100 --ops spmm --ks 10,16,20,32,40,64 --reps 5 2>&1 | tee gpurun_out/spmm_old.log
python scripts/kernel_sweep.py --m 100 --ops spmm --ks 10,16,20,32,40,64 --reps 5 2>&1 | tee gpurun_out/spmm_new.log
