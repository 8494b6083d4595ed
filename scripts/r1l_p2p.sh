#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 200 $TR 29561 tests/dist_worker_gpu.py 2>&1 | tail -4
echo "--- halo timing: p2p / nccl"
timeout 120 $TR 29562 scripts/halo_time.py 200 40 2>&1 | grep -E "world|rror" | head -5
B200_NO_P2P=1 timeout 120 $TR 29563 scripts/halo_time.py 200 40 2>&1 | grep -E "world|rror" | head -5
echo "--- bench N=2"
timeout 200 $TR 29564 bench.py --gpus 2 --warmup 1 --steps 1 --no-cpu > gpurun_out/bench_full_2gpu_c.log 2>&1
tail -1 gpurun_out/bench_full_2gpu_c.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=',d['n_gpus'], d['value'], d['result'], d['e2e'])
for k,v in d['kernel_classes'].items(): print(k, v)
" || tail -20 gpurun_out/bench_full_2gpu_c.log
