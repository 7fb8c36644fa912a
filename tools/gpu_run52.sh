#!/bin/bash
# N-GPU weak-scaling line; usage: bash tools/gpu_run52.sh N
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m_${N}gpu.log 2> gpurun_out/bench_1m_${N}gpu.err; echo "rc=$?" >> gpurun_out/bench_1m_${N}gpu.err
tail -2 gpurun_out/bench_1m_${N}gpu.err; tail -c 300 gpurun_out/bench_1m_${N}gpu.log
