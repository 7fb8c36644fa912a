#!/bin/bash
# executed-instruction histogram (per opcode and per source line) of the bf16 column kernels at L = 1M
mkdir -p gpurun_out
python tools/prof_conv.py 1000000 128 1 bf16 1 > gpurun_out/plain.log 2>&1 || exit 1
ncu --section SourceCounters --section WarpStateStats --clock-control none --import-source on -k regex:'k_col_fwd|k_col_inv' -s 10 -c 6 -o /tmp/p_col python tools/prof_conv.py 1000000 128 1 bf16 1 > gpurun_out/ncu_col.log 2>&1
ncu -i /tmp/p_col.ncu-rep --page source --csv --print-source sass > gpurun_out/col_sass.csv 2>/dev/null
ncu -i /tmp/p_col.ncu-rep --page source --csv --print-source cuda > gpurun_out/col_src.csv 2>/dev/null
gzip -f gpurun_out/col_sass.csv gpurun_out/col_src.csv
ls -la gpurun_out/col_*.gz; tail -3 gpurun_out/ncu_col.log
