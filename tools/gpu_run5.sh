#!/bin/bash
mkdir -p gpurun_out
python tools/prof_conv.py 1000000 12 1 bf16 1 > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_row_conv|k_col_fwd|k_col_inv' -s 56 -c 9 -o /tmp/prof_conv1m python tools/prof_conv.py 1000000 12 1 bf16 1 > gpurun_out/ncu_b.log 2>&1
ncu -i /tmp/prof_conv1m.ncu-rep --page raw --csv > gpurun_out/prof_conv1m_raw.csv 2>/dev/null
ncu -i /tmp/prof_conv1m.ncu-rep --page source --csv --print-source sass > /tmp/src_sass.csv 2>/dev/null
ncu -i /tmp/prof_conv1m.ncu-rep --page source --csv --print-source cuda > /tmp/src_cuda.csv 2>/dev/null
head -c 3000 /tmp/src_cuda.csv > gpurun_out/src_cuda_head.txt
python tools/ncu_top_lines.py /tmp/src_cuda.csv 40 > gpurun_out/top_lines_cuda.txt 2>&1
python tools/ncu_top_lines.py /tmp/src_sass.csv 40 > gpurun_out/top_lines_sass.txt 2>&1
gzip -c /tmp/src_cuda.csv > gpurun_out/src_cuda.csv.gz
ls -la /tmp/*.csv gpurun_out
