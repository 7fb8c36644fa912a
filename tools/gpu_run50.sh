#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "filter or operator or model" 2>&1 | tail -3
python tools/prof_filter.py 1000000 256 5 2>&1 | tail -3
HY_FWD_TF=4 python tools/prof_filter.py 1000000 256 5 2>&1 | tail -3
