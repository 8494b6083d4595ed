#!/bin/bash
# (inside gpurun, one GPU) launch list of the truncated bench command on the current build:
# plain run first (must exit 0), then the gpu__time_duration pass;  scripts/ncu_launches.sh [tag]
set -u
TAG=${1:-r2g}
mkdir -p gpurun_out
CMD="python bench.py --m 200 --max-iter 3 --warmup 0 --steps 1 --e2e-steps 0 --no-cpu --no-parity"
$CMD > gpurun_out/ncu_${TAG}_plain_bench_line.json 2> gpurun_out/ncu_${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_${TAG}_plain.err; exit 1; }
cut -c1-200 gpurun_out/ncu_${TAG}_plain_bench_line.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/ncu_${TAG}_launches.csv $CMD > gpurun_out/ncu_${TAG}_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/ncu_${TAG}_launches.csv)"
python scripts/ncu_summarise.py gpurun_out/ncu_${TAG}_launches.csv > gpurun_out/ncu_${TAG}_launches_summary.txt 2>&1; head -30 gpurun_out/ncu_${TAG}_launches_summary.txt
