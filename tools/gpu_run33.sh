#!/bin/bash
# ncu --set full of one group (128 rows) of every long-conv phase kernel at L = 1M (after the plain run exits 0)
mkdir -p gpurun_out
python tools/prof_conv.py 1000000 128 1 bf16 1 > gpurun_out/plain_c.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_row_conv|k_col_fwd|k_col_inv' -s 20 -c 10 -o /tmp/prof_conv1m python tools/prof_conv.py 1000000 128 1 bf16 1 > gpurun_out/ncu_c.log 2>&1
ncu -i /tmp/prof_conv1m.ncu-rep --page raw --csv > gpurun_out/prof_conv1m_raw.csv 2>/dev/null
python tools/ncu_family_columns.py gpurun_out/prof_conv1m_raw.csv gpurun_out/ncu_full_longconv_family_1m_128rows.csv
cut -d, -f1-9 gpurun_out/ncu_full_longconv_family_1m_128rows.csv; tail -2 gpurun_out/plain_c.log
