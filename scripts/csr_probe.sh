#!/bin/bash
# (inside gpurun) CSR SpMM on matrices without a diagonal image: the P1 lattice pencil with the image switched off and
# the mesh-ordered (unstructured) P1 pencil on data/cube4.dat's mesh after 5 refinements.
# profiles/csr_probe_r2g.log also holds the same sweep with the x gathers as ld.global.nc.L1::no_allocate (an experiment
# build, option csr_no_l1, not shipped: 1.3-1.6 x SLOWER -- the kernel is bound by L2 -> SM traffic, L1 hits matter).
TAG=${1:-r2}
OUT=gpurun_out/csr_probe_${TAG}.log
: > $OUT
echo "## B200_NO_DIA=1 p1_fem_kuhn m=100" >> $OUT
B200_NO_DIA=1 python scripts/kernel_sweep.py --m 100 --ks 16,32,40,64 --ops spmm >> $OUT 2>&1
echo "## cube4_p1_mesh refine=5" >> $OUT
python scripts/kernel_sweep.py --gen cube4_p1_mesh --m 5 --ks 16,32,40,64 --ops spmm >> $OUT 2>&1
cat $OUT
