#!/bin/bash
# Variants of the lattice SpMM (b200_spmm_lat.cu) inside the BlockPCG loop at n = m^3, k columns.
#   scripts/tune_lat.sh [m] [k]      (run under gpurun; prints two lines per variant)
m=${1:-200}; k=${2:-40}
run() { echo "## $*"; env "$@" B200_LAT_VERBOSE=1 timeout 120 python scripts/bpcg_time.py $m $k 2>&1 | grep -E "spmm_lat k=$k dot=1|spmm_ms" | sort -u | tail -2; }
run B200_X=0
run B200_LAT_NS=5
run B200_LAT_NS=6
run B200_LAT_TI=12 B200_LAT_TJ=7 B200_LAT_NS=5
run B200_LAT_TI=16 B200_LAT_TJ=6
run B200_LAT_TI=8 B200_LAT_TJ=12
run B200_LAT_TI=24 B200_LAT_TJ=4
run B200_LAT_NO_CONST=1
run B200_NO_LAT=1
