#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_1m.log 2> gpurun_out/bench_1m.err; echo "rc=$?" >> gpurun_out/bench_1m.err
tail -2 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench_1m.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_1m.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks']['sm_mhz'],d['clocks']['reasons'],d['roofline']['frac'],d['roofline']['ms_per_launch'],d['cpu_baseline']['value'])"
