#!/bin/bash
# (inside gpurun) the three arms side by side: reference on CCS+OpenMP, the reference's GCG over OPS_B200_Set (tier A),
# the device GCG through the OPS table (tier B); tests/test_gcg_gpu.py::test_tiers_side_by_side_at_size
TAG=${1:-r2}
OUT=gpurun_out/tiers_${TAG}.log
: > $OUT
GCGE_TIERS_AT=64,50 timeout 900 python -m pytest tests/test_gcg_gpu.py -q -s -m gpu -k tiers_side_by_side 2>&1 | grep -E "TIERS|passed|failed|Error" >> $OUT
GCGE_TIERS_AT=128,100 GCGE_TIERS_REF=0 timeout 900 python -m pytest tests/test_gcg_gpu.py -q -s -m gpu -k tiers_side_by_side 2>&1 | grep -E "TIERS|passed|failed|Error" >> $OUT
cat $OUT
