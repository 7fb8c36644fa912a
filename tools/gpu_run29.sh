#!/bin/bash
# where do the row-transform kernels (phase B) stall?  ncu --set full with source attribution, one launch each of
# the forward (mode 0) and the saved-spectrum backward (mode 5) at L = 1M, 128 rows
mkdir -p gpurun_out
python tools/prof_conv.py 1000000 128 1 bf16 1 > gpurun_out/plain.log 2>&1 || exit 1
# launch order in one run(): spectrum A,B | fwd A,B,C | bwd A,B,C | dk B,C  (1 group at 128 rows); warm-up runs = 2
ncu --set full --clock-control none --import-source on -k regex:k_row_conv -s 8 -c 4 -o /tmp/p_b python tools/prof_conv.py 1000000 128 1 bf16 1 > gpurun_out/ncu_b.log 2>&1
ncu -i /tmp/p_b.ncu-rep --page source --csv --print-source sass > /tmp/b_sass.csv 2>/dev/null
python tools/ncu_top_sass.py /tmp/b_sass.csv 60 > gpurun_out/top_sass_rowconv.txt 2>&1
ncu -i /tmp/p_b.ncu-rep --page source --csv --print-source cuda > /tmp/b_src.csv 2>/dev/null
python tools/ncu_top_lines.py /tmp/b_src.csv 40 > gpurun_out/top_lines_rowconv.txt 2>&1
ncu -i /tmp/p_b.ncu-rep --page details > gpurun_out/details_rowconv.txt 2>/dev/null
ncu -i /tmp/p_b.ncu-rep --page raw --csv > gpurun_out/raw_rowconv.csv 2>/dev/null
head -50 gpurun_out/top_lines_rowconv.txt
