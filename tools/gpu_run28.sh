#!/bin/bash
# quick A/B of a kernel change: GPU parity tests + the long-conv micro-benchmark at the headline shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/prof_conv.py 1000000 256 1 bf16 5 2>&1 | tee gpurun_out/prof_conv_1m.txt | tail -8
