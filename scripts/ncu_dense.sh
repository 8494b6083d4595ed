#!/bin/bash
# FP64 peak probe + dense sweep + ncu --set full of the Gram / LinearComb kernels (one GPU)
set -u
python scripts/fp64_peak.py 2>&1 | tee gpurun_out/fp64_peak.log
CMD="python scripts/kernel_sweep.py --m 100 --ops gram,lincomb --p 480 --ks 40,400 --reps 3"
$CMD > gpurun_out/dense_sweep.log 2>&1 || { tail -5 gpurun_out/dense_sweep.log; exit 1; }
cat gpurun_out/dense_sweep.log
ncu --set full --clock-control none --import-source on -k regex:"gram_partial|lincomb_kernel" -c 6 -o gpurun_out/ncu_${1:-r1b}_dense -f $CMD > gpurun_out/ncu_${1:-r1b}_dense.log 2>&1
echo "ncu rc=$?"
