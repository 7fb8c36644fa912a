#!/bin/bash
# round evidence r01n: tests, smoke, every bench workload, reference arm, ncu launch list of the bench command, step profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_1m.log 2> gpurun_out/bench_1m.err; echo "rc=$?" >> gpurun_out/bench_1m.err
for w in hyenadna-medium-160k hyenadna-small-32k hyenadna-tiny-1k; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "rc=$?" >> gpurun_out/bench_$w.err
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?" >> gpurun_out/bench_ref.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
gzip -f gpurun_out/launches_bench.csv
timeout 600 python tools/prof_step.py hyenadna-large-1m gpurun_out/step_profile_1m.txt > gpurun_out/prof_step.log 2>&1
tail -2 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; for f in gpurun_out/bench_*.err; do echo $f; tail -1 $f; done; tail -c 900 gpurun_out/bench_1m.log
