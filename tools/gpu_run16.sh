#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "four_step or one_million" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for cfg in "1 1024" "2 64" "3 96" "4 128" "3 48" "4 64" "2 128" "4 256" "3 192"; do set -- $cfg; echo "== nstream $1 budget $2 MB" >> gpurun_out/pipe.log; HY_NSTREAM=$1 HY_L2_MB=$2 timeout 300 python tools/prof_conv.py 1000000 128 1 bf16 3 >> gpurun_out/pipe.log 2>&1; done
echo "== 32k: nstream 1" >> gpurun_out/pipe.log; HY_NSTREAM=1 HY_L2_MB=1024 python tools/prof_conv.py 32768 256 8 bf16 3 >> gpurun_out/pipe.log 2>&1
echo "== 32k: nstream 3 / 96" >> gpurun_out/pipe.log; HY_NSTREAM=3 HY_L2_MB=96 python tools/prof_conv.py 32768 256 8 bf16 3 >> gpurun_out/pipe.log 2>&1
HY_NSTREAM=3 HY_L2_MB=96 timeout 600 python -m pytest tests -m gpu -x -q -k "four_step or one_million" >> gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; grep -E "==|long-conv|conv_fwd|conv_bwd" gpurun_out/pipe.log
