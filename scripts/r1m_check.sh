#!/bin/bash
python -m pytest tests/test_l3_gpu.py -m gpu -x -q -k "syev" 2>&1 | tail -2
B200_SYEV_PROF=1 python scripts/syev_time.py 240,480 2>&1 | tee gpurun_out/syev_time4.log
CMD="python scripts/kernel_sweep.py --m 100 --ops gram,lincomb --p 480 --ks 40,400 --reps 2"
ncu --set full --clock-control none --import-source on -k regex:"lincomb_kernel|gram_partial" -s 2 -c 2 -o gpurun_out/ncu_r1f_dense_k40 -f python scripts/kernel_sweep.py --m 100 --ops gram,lincomb --p 480 --ks 40 --reps 2 > gpurun_out/ncu_r1f_dense_k40.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lincomb_kernel" -s 1 -c 1 -o gpurun_out/ncu_r1f_lincomb_k400 -f python scripts/kernel_sweep.py --m 100 --ops lincomb --p 480 --ks 400 --reps 2 > gpurun_out/ncu_r1f_lincomb_k400.log 2>&1
ls -la gpurun_out/*r1f*
