#!/bin/bash
# 2-GPU weak-scaling check of the bench (torchrun, NCCL gradient all-reduce) + the reference arm launched the same way
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_1m_2gpu.log 2> gpurun_out/bench_1m_2gpu.err; echo "rc=$?" >> gpurun_out/bench_1m_2gpu.err
tail -1 gpurun_out/bench_1m_2gpu.err; tail -c 1500 gpurun_out/bench_1m_2gpu.log | head -c 700
