#!/bin/bash
# (inside `gpurun --gpus 2`)  slab-local construction: single-GPU tests, the 2-rank parity worker, then the headline
# solve with every rank generating only its own planes (bench.py --local-gen)
TAG=${1:-r2}
timeout 600 python -m pytest tests/test_slots_gpu.py -x -q -m gpu -k "local_rows or axpby or lattice or constant" > gpurun_out/local_rows_tests_${TAG}.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/local_rows_tests_${TAG}.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 300 $TR 29571 tests/dist_worker_gpu.py > gpurun_out/dist_worker_2gpu_${TAG}.log 2>&1
echo "worker rc=$?"; tail -3 gpurun_out/dist_worker_2gpu_${TAG}.log
timeout 400 $TR 29572 bench.py --gpus 2 --warmup 1 --steps 1 --no-cpu --local-gen > gpurun_out/bench_2gpu_localgen_${TAG}.log 2>&1
tail -1 gpurun_out/bench_2gpu_localgen_${TAG}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=',d['n_gpus'], d['value'], d['result'], d['e2e'])
print(d['phases_s']); print(d['parity_vs_golden'])
" || tail -20 gpurun_out/bench_2gpu_localgen_${TAG}.log
