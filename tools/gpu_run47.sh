#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/prof_filter.py 1000000 256 5 2>&1 | tail -4
HYENA_B200_TRUNK_SAVE_MAX_MB=0 python tools/prof_filter.py 1000000 256 5 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_1m.log 2> gpurun_out/bench_1m.err; echo "rc=$?" >> gpurun_out/bench_1m.err
tail -1 gpurun_out/bench_1m.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_1m.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks']['sm_mhz'],d['config']['peak_mem_gib'],d['roofline']['breakdown_ms_per_step'])"
