#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$TR 29551 scripts/halo_time.py 200 40 2>&1 | grep world
NCCL_MIN_P2P_NCHANNELS=8 $TR 29552 scripts/halo_time.py 200 40 2>&1 | grep world
NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 $TR 29553 scripts/halo_time.py 200 40 2>&1 | grep world
NCCL_NCHANNELS_PER_PEER=8 $TR 29554 scripts/halo_time.py 200 40 2>&1 | grep world
NCCL_P2P_USE_CUDA_MEMCPY=1 $TR 29555 scripts/halo_time.py 200 40 2>&1 | grep world
