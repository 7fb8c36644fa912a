#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/bench_1m.log 2>gpurun_out/bench_1m.err; tail -c 700 gpurun_out/bench_1m.log
echo; echo "== sgemm emulation probe"
python tools/prof_filter.py 1000000 256 5 2>&1 | tail -4
CUBLAS_EMULATE_SINGLE_PRECISION=1 python tools/prof_filter.py 1000000 256 5 2>&1 | tail -4
CUBLAS_EMULATE_SINGLE_PRECISION=1 CUBLAS_EMULATION_STRATEGY=performant python tools/prof_filter.py 1000000 256 5 2>&1 | tail -4
python -c "import torch; print(torch.backends.cuda.preferred_blas_library(), torch.version.cuda); import ctypes; print([l for l in open('/proc/self/maps').read().split() if 'cublas' in l][:2])"
