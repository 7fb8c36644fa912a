#!/bin/bash
mkdir -p gpurun_out
HY_BENCH_DEBUG=1 timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err
( time HY_BENCH_DEBUG=1 timeout 900 python bench.py ) > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
timeout 600 python bench.py --workload hyenadna-small-32k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small.log 2>&1
timeout 600 python bench.py --workload hyenadna-medium-160k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_medium.log 2>&1
timeout 600 python bench.py --workload hyenadna-tiny-1k --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tiny.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_a.log; grep -o '"clocks": {[^}]*}' gpurun_out/bench_a.log; tail -3 gpurun_out/bench_a.err
tail -c 1500 gpurun_out/bench_default.log; tail -5 gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_ref.log; tail -4 gpurun_out/bench_ref.err
for f in small medium tiny; do grep -o '"value": [0-9.]*, "unit": "nt/s", "n_gpus": 1, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' gpurun_out/bench_$f.log; grep -o '"frac": [0-9.]*' gpurun_out/bench_$f.log; done
