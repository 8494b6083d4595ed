#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_dist_gpu.py -m gpu -x -q -k "8" 2>&1 | tail -3
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --warmup 1 --steps 1 --no-cpu > gpurun_out/bench_full_${N}gpu.log 2>&1
tail -1 gpurun_out/bench_full_${N}gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=',d['n_gpus'], d['value'], d['result'], d['e2e'], d['roofline']['kernel'], round(d['roofline']['frac'],3))
print(d['phases_s'])
for k,v in d['kernel_classes'].items(): print(k, v)
"
done
