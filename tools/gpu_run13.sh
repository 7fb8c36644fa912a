#!/bin/bash
mkdir -p gpurun_out
echo "== default" > gpurun_out/ab.log
python tools/prof_conv.py 1000000 64 1 bf16 3 >> gpurun_out/ab.log 2>&1
python tools/prof_conv.py 32768 256 8 bf16 3 >> gpurun_out/ab.log 2>&1
python tools/prof_conv.py 160000 256 1 bf16 3 >> gpurun_out/ab.log 2>&1
echo "== small CTAs" >> gpurun_out/ab.log
export HYENA_B200_LIB=$PWD/dna_b200/lib/libhyena_b200_small.so
python tools/prof_conv.py 1000000 64 1 bf16 3 >> gpurun_out/ab.log 2>&1
python tools/prof_conv.py 32768 256 8 bf16 3 >> gpurun_out/ab.log 2>&1
python tools/prof_conv.py 160000 256 1 bf16 3 >> gpurun_out/ab.log 2>&1
timeout 300 python -m pytest tests -m gpu -x -q -k "four_step or one_million_direct" > gpurun_out/pytest_small.log 2>&1
cat gpurun_out/ab.log; tail -3 gpurun_out/pytest_small.log
