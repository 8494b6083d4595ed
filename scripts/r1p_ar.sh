#!/bin/bash
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 200 $TR 29581 tests/dist_worker_gpu.py 2>&1 | tail -2
timeout 240 $TR 29582 bench.py --gpus $N --warmup 1 --steps 1 --no-cpu > gpurun_out/bench_full_${N}gpu_ar.log 2>&1
tail -1 gpurun_out/bench_full_${N}gpu_ar.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=',d['n_gpus'], d['value'], d['result'], d['e2e'])
for k,v in d['kernel_classes'].items(): print(k, v)
" || tail -20 gpurun_out/bench_full_${N}gpu_ar.log
