#!/bin/bash
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; uptime >> gpurun_out/host.txt; head -20 /proc/cpuinfo | grep "model name" | head -1 >> gpurun_out/host.txt
HY_BENCH_DEBUG=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err
HY_BENCH_DEBUG=1 HY_NO_CLOCK_SAMPLER=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.log 2> gpurun_out/bench_b.err
HY_BENCH_DEBUG=1 HY_NO_CLOCK_SAMPLER=1 PYTORCH_CUDA_ALLOC_CONF=expandable_segments:False timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c.log 2> gpurun_out/bench_c.err
uptime >> gpurun_out/host.txt
cat gpurun_out/host.txt; for f in a b c; do echo "== $f"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_$f.log; tail -3 gpurun_out/bench_$f.err; done
