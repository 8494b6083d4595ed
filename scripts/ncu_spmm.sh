#!/bin/bash
# ncu --set full of the SpMM kernel at k=40 (n = 1 M P1-FEM pencil), after a plain run
set -u
CMD="python scripts/kernel_sweep.py --m 100 --ops spmm --ks ${2:-40} --reps 3"
$CMD > gpurun_out/spmm_plain.log 2>&1 || { tail -5 gpurun_out/spmm_plain.log; exit 1; }
cat gpurun_out/spmm_plain.log
ncu --set full --clock-control none --import-source on -k regex:"spmm_" -s 1 -c 2 -o gpurun_out/ncu_${1:-r1c}_spmm -f $CMD > gpurun_out/ncu_${1:-r1c}_spmm.log 2>&1
echo "ncu rc=$?"
