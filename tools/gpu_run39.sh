#!/bin/bash
# ncu --set full of the TC filter-out backward kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_filter_out_bwd -c 2 -o gpurun_out/filter_out_bwd -f python tools/prof_filter.py 1000000 256 1 > gpurun_out/ncu_fob.log 2>&1
ncu -i gpurun_out/filter_out_bwd.ncu-rep --page raw --csv > gpurun_out/filter_out_bwd_raw.csv 2>/dev/null
tail -5 gpurun_out/ncu_fob.log
