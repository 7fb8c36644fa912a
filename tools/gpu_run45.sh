#!/bin/bash
# channel-partitioned operator: slab test + strong scaling 1/2/4/8 GPUs
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "channel_slabs" 2>&1 | tail -3
: > gpurun_out/channel_slabs_scaling.jsonl
timeout 300 python tools/bench_channel_slabs.py >> gpurun_out/channel_slabs_scaling.jsonl 2> gpurun_out/channel_slabs.err
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n tools/bench_channel_slabs.py >> gpurun_out/channel_slabs_scaling.jsonl 2>> gpurun_out/channel_slabs.err
done
tail -5 gpurun_out/channel_slabs.err; cut -c1-330 gpurun_out/channel_slabs_scaling.jsonl
