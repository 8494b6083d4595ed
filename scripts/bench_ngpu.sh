#!/bin/bash
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
if [ "$2" != "benchonly" ]; then timeout 200 $TR 29571 tests/dist_worker_gpu.py 2>&1 | tail -2; fi
timeout 240 $TR 29572 bench.py --gpus $N --warmup 1 --steps 1 --no-cpu > gpurun_out/bench_full_${N}gpu_p2p.log 2>&1
tail -1 gpurun_out/bench_full_${N}gpu_p2p.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=',d['n_gpus'], d['value'], d['result'], d['e2e'])
print(d['phases_s'])
for k,v in d['kernel_classes'].items(): print(k, v)
" || tail -20 gpurun_out/bench_full_${N}gpu_p2p.log
