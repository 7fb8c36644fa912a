#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_1m.log 2> gpurun_out/bench_1m.err; echo "rc=$?" >> gpurun_out/bench_1m.err
timeout 600 python tools/prof_step.py hyenadna-large-1m gpurun_out/step_profile_1m.txt > gpurun_out/prof_step.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -c 1200 gpurun_out/bench_1m.log
