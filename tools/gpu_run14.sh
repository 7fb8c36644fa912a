#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/prof_conv.py 1000000 64 1 bf16 3 > gpurun_out/ab.log 2>&1
python tools/prof_conv.py 32768 256 8 bf16 3 >> gpurun_out/ab.log 2>&1
python tools/prof_conv.py 160000 256 1 bf16 3 >> gpurun_out/ab.log 2>&1
python tools/prof_conv.py 1000000 64 1 fp32 3 >> gpurun_out/ab.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/ab.log
