#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m cProfile -o /tmp/tiny.prof bench.py --workload hyenadna-tiny-1k --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tiny.log 2>&1
python - <<'PY' > gpurun_out/cprofile_tiny.txt 2>&1
import pstats
p = pstats.Stats('/tmp/tiny.prof'); p.sort_stats('tottime').print_stats(35)
PY
timeout 600 python -m cProfile -o /tmp/med.prof bench.py --workload hyenadna-medium-160k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_medium.log 2>&1
python - <<'PY' > gpurun_out/cprofile_medium.txt 2>&1
import pstats
p = pstats.Stats('/tmp/med.prof'); p.sort_stats('tottime').print_stats(35)
PY
head -60 gpurun_out/cprofile_tiny.txt; head -50 gpurun_out/cprofile_medium.txt
