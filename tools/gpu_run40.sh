#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "filter" 2>&1 | tail -8
python tools/prof_filter.py 1000000 256 5 2>&1 | tail -4
