#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/prof_conv.py 1000000 256 1 bf16 5 2>&1 | tail -8
