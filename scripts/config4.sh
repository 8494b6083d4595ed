#!/bin/bash
# (inside `gpurun --gpus 8`)  BASELINE config 4: 27-point operator, n = 400^3 = 64 M, nev = 100, row-sharded over 8 GPUs;
# every rank generates and uploads only its own 50 lattice planes (bench.py --local-gen).  Log under gpurun_out/.
N=${1:-8}; M=${2:-400}; NEV=${3:-100}; TAG=${4:-r2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29574"
timeout 1500 $TR bench.py --gpus $N --workload q1_27pt --lattice $M --nev $NEV --local-gen --warmup 1 --steps 1 --no-cpu \
	> gpurun_out/bench_config4_q1_27pt_m${M}_nev${NEV}_${N}gpu_${TAG}.log 2>&1
echo "config4 rc=$?"
tail -1 gpurun_out/bench_config4_q1_27pt_m${M}_nev${NEV}_${N}gpu_${TAG}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=',d['n_gpus'], d['value'], d['result'], d['e2e'], d['setup_s'])
print(d['phases_s']); print(d['parity_at_full_size'])
for k,v in d['kernel_classes'].items(): print(k, v)
" || tail -20 gpurun_out/bench_config4_q1_27pt_m${M}_nev${NEV}_${N}gpu_${TAG}.log
