#!/bin/bash
# re-entry check of HEAD: GPU tests, the default bench line, the operator sweep against the reference's GPU path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_1m.log 2> gpurun_out/bench_1m.err; echo "rc=$?" >> gpurun_out/bench_1m.err
timeout 900 python tools/sweep_operator.py gpurun_out/sweep_operator.jsonl > gpurun_out/sweep_operator.txt 2>&1; echo "rc=$?" >> gpurun_out/sweep_operator.txt
tail -3 gpurun_out/pytest_gpu.log; tail -c 600 gpurun_out/bench_1m.log; tail -1 gpurun_out/bench_1m.err; tail -25 gpurun_out/sweep_operator.txt
