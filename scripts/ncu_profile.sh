#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, one GPU).  Usage: scripts/ncu_profile.sh <tag> <kernel-regex> [m] [max-iter]
# 1. plain run of the truncated bench command (exit 0 required before any ncu pass)
# 2. launch list: every kernel launch of the same command with its device time
# 3. one --set full capture of the kernel named by <kernel-regex>
set -u
TAG=${1:-r1}; KREGEX=${2:-bpcg_update_xr}; M=${3:-200}; IT=${4:-3}
CMD="python bench.py --m $M --max-iter $IT --warmup 0 --steps 1 --e2e-steps 0 --no-cpu"
mkdir -p gpurun_out
$CMD > gpurun_out/ncu_${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/ncu_${TAG}_plain.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/ncu_${TAG}_launches.csv $CMD > gpurun_out/ncu_${TAG}_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/ncu_${TAG}_launches.csv)"
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 20 -c 2 -o gpurun_out/ncu_${TAG}_full -f $CMD > gpurun_out/ncu_${TAG}_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/ | tail -8
