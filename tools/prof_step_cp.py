"""Kernel-level breakdown of one channel-partition step on rank 0 (torch.profiler) under torchrun.
usage: torchrun --nproc-per-node G tools/prof_step_cp.py [seqlen] [out.txt]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F, torch.distributed as dist
import bench
from dna_b200.standalone import HyenaDNAModel
from dna_b200.dp import FlatGradAllReduce, ChannelPartition, set_channel_partition

cfg = dict(bench.WORKLOADS["hyenadna-large-1m"])
L = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["seqlen"]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(2222)
model = HyenaDNAModel(d_model=cfg["d_model"], n_layer=cfg["n_layer"], d_inner=cfg["d_inner"], vocab_size=12, pad_vocab_size_multiple=8,
                      embed_dropout=0.0, lm_head=True, layer=dict(l_max=L + 2, **bench.LAYER_CFG)).to(dev).train()
part = ChannelPartition()
set_channel_partition(model, part)
red = FlatGradAllReduce(model.parameters())
opt = torch.optim.AdamW(model.parameters(), lr=6e-4, fused=True)
Lc = L // world
ids = torch.randint(7, 11, (1, Lc + 1), device=dev)

def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(ids[:, :-1])
    loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), ids[:, 1:].reshape(-1), reduction="sum") / L
    red.zero(); loss.backward(); red.allreduce(average=False); opt.step()
    return loss

for _ in range(3):
    step()
torch.cuda.synchronize(); dist.barrier()
ts = []
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append((round((t1 - t0) * 1e3, 1), round((t2 - t0) * 1e3, 1)))
if rank == 0:
    print("cpu-enqueue ms / total ms:", ts, flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
if rank == 0:
    txt = prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=60, max_name_column_width=60)
    print(txt)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
dist.barrier(); dist.destroy_process_group()
