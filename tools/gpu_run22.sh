#!/bin/bash
mkdir -p gpurun_out
python tools/prof_conv.py 1000000 128 1 bf16 1 > gpurun_out/plain_c.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_row_conv|k_col_fwd|k_col_inv' -s 20 -c 10 -o /tmp/prof_conv1m python tools/prof_conv.py 1000000 128 1 bf16 1 > gpurun_out/ncu_c.log 2>&1
ncu -i /tmp/prof_conv1m.ncu-rep --page raw --csv > gpurun_out/prof_conv1m_raw.csv 2>/dev/null
ncu -i /tmp/prof_conv1m.ncu-rep --page source --csv --print-source cuda > /tmp/src_cuda.csv 2>/dev/null
python tools/ncu_top_lines.py /tmp/src_cuda.csv 60 > gpurun_out/top_lines_cuda.txt 2>&1
ls -la /tmp/*.csv gpurun_out | tail -8; tail -3 gpurun_out/plain_c.log
