#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_l3_gpu.py tests/test_slots_gpu.py -m gpu -x -q -k "syev or spmm or pcg" > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
tail -5 gpurun_out/pytest_gpu4.log
B200_SYEV_PROF=1 python scripts/syev_time.py 240,480 2>&1 | tee gpurun_out/syev_time2.log
for cfg in "2 0 0" "3 0 0" "4 0 0" "2 8 0" "1 8 8" "2 0 8" "3 0 2" "4 0 2" "2 0 2"; do
  set -- $cfg
  echo "CTAS=$1 NS=$2 RB=$3"; B200_SPMM_CTAS=$1 B200_SPMM_NS=$2 B200_SPMM_RB=$3 python scripts/kernel_sweep.py --m 100 --ops spmm --ks 40 --reps 7
done 2>&1 | grep -v "^#" | tee gpurun_out/spmm_variants2.log
for c in 2 3; do echo "CTAS=$c"; B200_SPMM_CTAS=$c python scripts/kernel_sweep.py --m 100 --ops spmm --ks 8,10,16,20,24,32,40,48,50,64 --reps 5; done 2>&1 | tee gpurun_out/spmm_new2.log
bash scripts/ncu_spmm.sh r1e 40 2>&1 | tail -3
