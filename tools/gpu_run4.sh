#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for args in "1000000 24 1 bf16 3" "1000000 24 1 fp32 3" "32768 256 8 bf16 3" "160000 256 1 bf16 3" "1024 128 64 bf16 5" "4096 256 8 bf16 5"; do
  timeout 300 python tools/prof_conv.py $args >> gpurun_out/prof_conv.log 2>&1
done
for mb in 24 32 64 96; do echo "L2 budget $mb MB" >> gpurun_out/prof_conv.log; HY_L2_MB=$mb timeout 300 python tools/prof_conv.py 1000000 48 1 bf16 3 >> gpurun_out/prof_conv.log 2>&1; done
timeout 600 python tools/prof_step.py hyenadna-large-1m gpurun_out/step_profile_1m.txt > gpurun_out/prof_step.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1; echo "rc=$?" >> gpurun_out/bench_1m.log
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/prof_conv.log; head -4 gpurun_out/prof_step.log; tail -2 gpurun_out/bench_1m.log
