#!/bin/bash
# (inside `gpurun --gpus N`)  multi-GPU parity worker, then bench.py --gpus N under torchrun; logs under gpurun_out/
N=${1:-4}; TAG=${3:-r2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
if [ "$2" != "benchonly" ]; then
	timeout 300 $TR 29571 tests/dist_worker_gpu.py > gpurun_out/dist_worker_${N}gpu_${TAG}.log 2>&1
	echo "worker rc=$?"; tail -3 gpurun_out/dist_worker_${N}gpu_${TAG}.log
fi
timeout 400 $TR 29572 bench.py --gpus $N --warmup 1 --steps 1 --no-cpu > gpurun_out/bench_${N}gpu_${TAG}.log 2>&1
tail -1 gpurun_out/bench_${N}gpu_${TAG}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=',d['n_gpus'], d['value'], d['result'], d['e2e'])
print(d['phases_s']); print(d['parity_vs_golden'])
for k,v in d['kernel_classes'].items(): print(k, v)
" || tail -20 gpurun_out/bench_${N}gpu_${TAG}.log
