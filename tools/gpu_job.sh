#!/bin/bash
# One parameterised GPU job for `gpurun -- bash tools/gpu_job.sh <what> [args]` (replaces the numbered one-offs of
# round 1).  Everything it writes goes to gpurun_out/.
#   tests [pytest args]     python -m pytest tests -m gpu -q ...
#   smoke                   __graft_entry__.smoke()
#   bench [bench args]      python bench.py ...            -> gpurun_out/bench.json (+ .err)
#   benchn N [bench args]   torchrun --nproc-per-node N bench.py --gpus N ...
#   launches [bench args]   plain run, then the ncu launch list of the same command -> gpurun_out/launches.csv
#   ncu <regex> <cmd...>    plain run of <cmd>, then ncu --set full of kernels matching <regex> -> gpurun_out/prof.ncu-rep
#   py <script> [args]      python <script> ... > gpurun_out/<script>.log
set -u
mkdir -p gpurun_out
what=$1; shift
case "$what" in
  tests)   timeout 1500 python -m pytest tests -m gpu -q -x "$@" 2>&1 | tee gpurun_out/pytest_gpu.log | tail -15 ;;
  smoke)   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log | tail -5 ;;
  bench)   timeout 900 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err ;;
  benchn)  N=$1; shift
           timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 2953"$N" \
             bench.py --gpus "$N" "$@" > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
           tail -c 1500 gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err ;;
  launches) timeout 600 python bench.py "$@" > gpurun_out/launches_plain.log 2>&1 &&
           timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
             python bench.py "$@" > gpurun_out/launches_ncu.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv ;;
  ncu)     rx=$1; shift
           timeout 600 "$@" > gpurun_out/ncu_plain.log 2>&1 &&
           timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c ${NCU_COUNT:-6} -f -o gpurun_out/prof "$@" > gpurun_out/ncu.log 2>&1
           echo "rc=$?"; tail -3 gpurun_out/ncu.log ;;
  py)      s=$1; shift; timeout 1200 python "$s" "$@" 2>&1 | tee gpurun_out/$(basename "$s" .py).log | tail -40 ;;
  *) echo "unknown job $what"; exit 2 ;;
esac
