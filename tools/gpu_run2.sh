#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for args in "1000000 24 1 bf16 3" "1000000 24 1 fp32 3" "32768 256 8 bf16 3" "160000 256 1 bf16 3" "1024 128 64 bf16 5" "4096 256 8 bf16 5"; do
  timeout 300 python tools/prof_conv.py $args >> gpurun_out/prof_conv.log 2>&1
done
timeout 600 python tools/prof_step.py hyenadna-large-1m gpurun_out/step_profile_1m.txt > gpurun_out/prof_step.log 2>&1
python tools/prof_conv.py 1000000 12 1 bf16 1 > gpurun_out/plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_conv1m.csv python tools/prof_conv.py 1000000 12 1 bf16 1 > gpurun_out/ncu_a.log 2>&1
python tools/prof_conv.py 1000000 12 1 bf16 1 > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_row_conv|k_col_fwd|k_col_inv' -s 52 -c 26 -o gpurun_out/prof_conv1m python tools/prof_conv.py 1000000 12 1 bf16 1 > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/prof_conv.log; head -8 gpurun_out/prof_step.log
