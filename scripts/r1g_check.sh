#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_l3_gpu.py tests/test_slots_gpu.py -m gpu -x -q -k "syev or spmm or pcg" > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log
tail -5 gpurun_out/pytest_gpu5.log
B200_SYEV_PROF=1 python scripts/syev_time.py 240,480 2>&1 | tee gpurun_out/syev_time3.log
for cfg in "2 256 2" "2 256 3" "2 128 2" "2 128 3" "2 128 4" "1 256 2" "1 256 3" "1 128 3"; do
  set -- $cfg
  echo "CP=$1 NT=$2 CTAS=$3"; B200_SPMM_CP=$1 B200_SPMM_NT=$2 B200_SPMM_CTAS=$3 python scripts/kernel_sweep.py --m 100 --ops spmm --ks 40 --reps 7
done 2>&1 | grep -v "^#" | tee gpurun_out/spmm_variants3.log
for cp in 1 2; do for c in 2 3; do echo "CP=$cp CTAS=$c"; B200_SPMM_CP=$cp B200_SPMM_CTAS=$c python scripts/kernel_sweep.py --m 100 --ops spmm --ks 16,20,24,32,40,48,64 --reps 5; done; done 2>&1 | tee gpurun_out/spmm_new3.log
