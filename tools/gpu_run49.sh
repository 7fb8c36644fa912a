#!/bin/bash
# ncu --set full of the three filter kernels at L = 1M (forward, tensor-core last-Linear backward, saved-trunk backward)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_filter_fwd_fast|k_filter_out_bwd$|k_filter_trunk_bwd' -s 6 -c 3 -o gpurun_out/filter_kernels -f python tools/prof_filter.py 1000000 256 2 > gpurun_out/ncu_filter.log 2>&1
ncu -i gpurun_out/filter_kernels.ncu-rep --page raw --csv > gpurun_out/filter_kernels_raw.csv 2>/dev/null
tail -4 gpurun_out/ncu_filter.log
