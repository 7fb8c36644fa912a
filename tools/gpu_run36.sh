#!/bin/bash
# what bounds the filter kernels?  ncu --set full (raw page) of the fused forward and the fused trunk backward at L = 1M
mkdir -p gpurun_out
python tools/prof_filter.py 1000000 256 2 > gpurun_out/plain_f.log 2>&1 || { tail -5 gpurun_out/plain_f.log; exit 1; }
timeout 900 ncu --set full --clock-control none -k regex:'k_filter' -s 6 -c 4 -o /tmp/prof_filter python tools/prof_filter.py 1000000 256 2 > gpurun_out/ncu_f.log 2>&1
ncu -i /tmp/prof_filter.ncu-rep --page raw --csv > gpurun_out/prof_filter_raw.csv 2>/dev/null
ncu -i /tmp/prof_filter.ncu-rep --page details > gpurun_out/details_filter.txt 2>/dev/null
tail -4 gpurun_out/plain_f.log
