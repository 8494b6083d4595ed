#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, one GPU):  scripts/ncu_r2.sh [tag]
#  1. plain run of the truncated bench command (must exit 0 before any ncu pass), then its launch list
#  2. --set full captures of the shipped hot kernels at the headline shapes (n = 8.0 M, k = 40):
#     BlockPCG update kernels + lattice SpMM with the fused dot (inside a BlockPCG loop), lattice SpMM alone,
#     TMA-fed Gram and LinearComb (p = 480, q = 40)
set -u
TAG=${1:-r2b}
mkdir -p gpurun_out
CMD="python bench.py --m 200 --max-iter 3 --warmup 0 --steps 1 --e2e-steps 0 --no-cpu --no-parity"
$CMD > gpurun_out/ncu_${TAG}_plain_bench_line.json 2> gpurun_out/ncu_${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_${TAG}_plain.err; exit 1; }
cut -c1-300 gpurun_out/ncu_${TAG}_plain_bench_line.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/ncu_${TAG}_launches.csv $CMD > gpurun_out/ncu_${TAG}_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/ncu_${TAG}_launches.csv)"
C1="python scripts/bpcg_time.py 200 40"
$C1 > gpurun_out/ncu_${TAG}_bpcg_plain.log 2>&1 || { tail -5 gpurun_out/ncu_${TAG}_bpcg_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"bpcg_update_px|bpcg_update_r|spmm_lat" -s 9 -c 6 -o gpurun_out/ncu_${TAG}_bpcg_iteration -f $C1 > gpurun_out/ncu_${TAG}_bpcg_iteration.log 2>&1
echo "bpcg capture rc=$?"
C2="python scripts/kernel_sweep.py --m 200 --ops spmm,gram,lincomb --p 480 --ks 40 --reps 2"
$C2 > gpurun_out/ncu_${TAG}_dense_plain.log 2>&1 || { tail -5 gpurun_out/ncu_${TAG}_dense_plain.log; exit 1; }
cat gpurun_out/ncu_${TAG}_dense_plain.log
ncu --set full --clock-control none --import-source on -k regex:"gram_tma2|lincomb_tma|spmm_lat" -s 3 -c 6 -o gpurun_out/ncu_${TAG}_dense -f $C2 > gpurun_out/ncu_${TAG}_dense.log 2>&1
echo "dense capture rc=$?"; ls -la gpurun_out/*${TAG}*
