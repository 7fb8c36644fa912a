#!/bin/bash
# TC filter-out backward: parity + timing + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/prof_filter.py 1000000 256 5 > gpurun_out/prof_filter_tc.txt 2>&1
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_1m.log 2> gpurun_out/bench_1m.err; echo "rc=$?" >> gpurun_out/bench_1m.err
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/prof_filter_tc.txt | tail -5; tail -c 1500 gpurun_out/bench_1m.log
