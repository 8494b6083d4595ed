#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu9.log
tail -4 gpurun_out/pytest_gpu9.log
timeout 120 python scripts/lincomb_race.py 65536 8 "[(2,302,30),(178,302,222),(0,480,400)]" 2>&1 | grep "runs that differ" | tee gpurun_out/lincomb_race.log
timeout 100 python scripts/kernel_sweep.py --m 200 --ops spmm --ks 40 --p 40 --reps 5 2>&1 | tail -1
timeout 400 python bench.py --warmup 1 --steps 1 --no-cpu > gpurun_out/bench_full_8.log 2>&1
tail -1 gpurun_out/bench_full_8.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['result'], d['e2e'])
print(d['phases_s'])
for k,v in d['kernel_classes'].items(): print(k, v)
"
