#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu6.log
tail -4 gpurun_out/pytest_gpu6.log
python scripts/kernel_sweep.py --m 100 --ops gram,lincomb --p 480 --ks 16,40,400 --reps 3 2>&1 | tee gpurun_out/dense_sweep2.log
python bench.py --warmup 1 --steps 1 --no-cpu > gpurun_out/bench_full_5.log 2>&1
tail -1 gpurun_out/bench_full_5.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['result'], d['e2e'], d['roofline']['kernel'], round(d['roofline']['frac'],3))
print(d['phases_s'])
for k,v in d['kernel_classes'].items(): print(k, v)
"
