#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_step.py hyenadna-large-1m gpurun_out/step_profile_1m.txt > gpurun_out/prof_step.log 2>&1
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True timeout 600 python tools/prof_step.py hyenadna-large-1m gpurun_out/step_profile_1m_expseg.txt > gpurun_out/prof_step_expseg.log 2>&1
python tools/prof_conv.py 1000000 12 1 bf16 1 > gpurun_out/plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_conv1m.csv python tools/prof_conv.py 1000000 12 1 bf16 1 > gpurun_out/ncu_a.log 2>&1
python tools/prof_conv.py 1000000 12 1 bf16 1 > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_row_conv|k_col_fwd|k_col_inv' -s 56 -c 9 -o /tmp/prof_conv1m python tools/prof_conv.py 1000000 12 1 bf16 1 > gpurun_out/ncu_b.log 2>&1
ncu -i /tmp/prof_conv1m.ncu-rep --page raw --csv > gpurun_out/prof_conv1m_raw.csv 2>/dev/null
ncu -i /tmp/prof_conv1m.ncu-rep --page details --csv > gpurun_out/prof_conv1m_details.csv 2>/dev/null
ncu -i /tmp/prof_conv1m.ncu-rep --page source --csv > /tmp/src.csv 2>/dev/null; gzip -c /tmp/src.csv > gpurun_out/prof_conv1m_source.csv.gz
ls -la /tmp/prof_conv1m.ncu-rep gpurun_out/
sz=$(stat -c %s /tmp/prof_conv1m.ncu-rep); if [ "$sz" -lt 30000000 ]; then cp /tmp/prof_conv1m.ncu-rep gpurun_out/; fi
head -6 gpurun_out/prof_step.log; head -6 gpurun_out/prof_step_expseg.log
