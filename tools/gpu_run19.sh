#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "filter or module or golden or operator" > gpurun_out/pytest_small.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_small.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_1m.log 2> gpurun_out/bench_1m.err; echo "rc=$?" >> gpurun_out/bench_1m.err
tail -3 gpurun_out/pytest_small.log; tail -c 1200 gpurun_out/bench_1m.log; tail -2 gpurun_out/bench_1m.err
