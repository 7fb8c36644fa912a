#!/bin/bash
# first GPU pass: tests, smoke, small + headline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --workload hyenadna-small-32k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small.log 2>&1; echo "rc=$?" >> gpurun_out/bench_small.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1; echo "rc=$?" >> gpurun_out/bench_1m.log
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_small.log gpurun_out/bench_1m.log
