#!/bin/bash
# 8-GPU weak-scaling line (one 1M-nt sequence per rank, flat NCCL gradient all-reduce)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m_8gpu.log 2> gpurun_out/bench_1m_8gpu.err; echo "rc=$?" >> gpurun_out/bench_1m_8gpu.err
tail -3 gpurun_out/bench_1m_8gpu.err; tail -c 600 gpurun_out/bench_1m_8gpu.log
