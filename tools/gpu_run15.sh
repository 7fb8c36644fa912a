#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_1m.log 2> gpurun_out/bench_1m.err; echo "rc=$?" >> gpurun_out/bench_1m.err
timeout 600 python tools/prof_step.py hyenadna-large-1m gpurun_out/step_profile_1m.txt > gpurun_out/prof_step.log 2>&1
# ncu launch list of the bench command (contract): plain run first, then the same command under ncu
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# DRAM traffic of one layer's long-conv family (256 rows @ 1M): full set on the 11 launches of the 3rd run()
python tools/prof_conv.py 1000000 256 1 bf16 1 > gpurun_out/plain_conv.log 2>&1 && \
ncu --set full --clock-control none -k regex:'k_col_fwd|k_row_conv|k_col_inv' -s 22 -c 11 -o /tmp/fam python tools/prof_conv.py 1000000 256 1 bf16 1 > gpurun_out/ncu_fam.log 2>&1
ncu -i /tmp/fam.ncu-rep --page raw --csv > gpurun_out/fam_raw.csv 2>/dev/null
gzip -f gpurun_out/launches_bench.csv
tail -c 1200 gpurun_out/bench_1m.log; tail -2 gpurun_out/bench_1m.err; ls -la gpurun_out
