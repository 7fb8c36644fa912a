#!/bin/bash
mkdir -p gpurun_out
python tools/prof_filter.py 1000000 256 5 > gpurun_out/filter.log 2>&1
HY_TRUNK_MINB=2 python tools/prof_filter.py 1000000 256 5 >> gpurun_out/filter.log 2>&1
cat gpurun_out/filter.log
