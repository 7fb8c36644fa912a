#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_1m.log 2> gpurun_out/bench_1m.err; echo "rc=$?" >> gpurun_out/bench_1m.err
for mb in 48 96 192 512 1024 2048; do
  echo "== L2_MB=$mb" >> gpurun_out/sweep.log
  HY_L2_MB=$mb python tools/prof_conv.py 1000000 256 1 bf16 3 >> gpurun_out/sweep.log 2>&1
done
echo "== 32K B=8" >> gpurun_out/sweep.log
python tools/prof_conv.py 32768 256 8 bf16 3 >> gpurun_out/sweep.log 2>&1
echo "== 160K B=1" >> gpurun_out/sweep.log
python tools/prof_conv.py 160000 256 1 bf16 3 >> gpurun_out/sweep.log 2>&1
timeout 600 python tools/prof_step.py hyenadna-large-1m gpurun_out/step_profile_1m.txt > gpurun_out/prof_step.log 2>&1
tail -c 600 gpurun_out/bench_1m.log; grep -E "==|long-conv" gpurun_out/sweep.log
