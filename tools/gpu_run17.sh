#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/pipe.log
for cfg in "1 1024 0" "3 96 1" "4 96 1" "4 64 1" "2 64 1" "3 48 1"; do set -- $cfg; echo "== nstream $1 budget $2 MB persist $3" >> gpurun_out/pipe.log; if [ "$3" = "1" ]; then export HY_PERSIST=1; else unset HY_PERSIST; fi; HY_NSTREAM=$1 HY_L2_MB=$2 timeout 300 python tools/prof_conv.py 1000000 128 1 bf16 3 >> gpurun_out/pipe.log 2>&1; done
grep -E "==|long-conv|conv_fwd|conv_bwd|persisting" gpurun_out/pipe.log
