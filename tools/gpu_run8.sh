#!/bin/bash
mkdir -p gpurun_out
for mb in 96 148 296 592 1184; do echo "L2 budget $mb MB" >> gpurun_out/prof_l2.log; HY_L2_MB=$mb timeout 300 python tools/prof_conv.py 1000000 148 1 bf16 3 >> gpurun_out/prof_l2.log 2>&1; done
HY_BENCH_DEBUG=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err
cat gpurun_out/prof_l2.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_a.log; grep -o '"clocks": {[^}]*}' gpurun_out/bench_a.log; tail -3 gpurun_out/bench_a.err
