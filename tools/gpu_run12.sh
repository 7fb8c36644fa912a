#!/bin/bash
mkdir -p gpurun_out
python tools/prof_conv.py 1000000 24 1 bf16 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k k_col_fwd -s 7 -c 1 -o /tmp/p_a python tools/prof_conv.py 1000000 24 1 bf16 1 > gpurun_out/ncu_a.log 2>&1
ncu -i /tmp/p_a.ncu-rep --page source --csv --print-source sass > /tmp/a_sass.csv 2>/dev/null
python tools/ncu_top_sass.py /tmp/a_sass.csv 70 > gpurun_out/top_sass_colfwd.txt 2>&1
ncu -i /tmp/p_a.ncu-rep --page details > gpurun_out/details_colfwd.txt 2>/dev/null
python tools/prof_conv.py 1000000 24 1 bf16 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k k_col_inv -s 6 -c 1 -o /tmp/p_c python tools/prof_conv.py 1000000 24 1 bf16 1 > gpurun_out/ncu_c.log 2>&1
ncu -i /tmp/p_c.ncu-rep --page source --csv --print-source sass > /tmp/c_sass.csv 2>/dev/null
python tools/ncu_top_sass.py /tmp/c_sass.csv 50 > gpurun_out/top_sass_colinv.txt 2>&1
ncu -i /tmp/p_c.ncu-rep --page details > gpurun_out/details_colinv.txt 2>/dev/null
head -5 gpurun_out/top_sass_colfwd.txt
