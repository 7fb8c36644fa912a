"""Exchange step of the channel partition in isolation: GB/s of the pull kernel variants vs NCCL all_to_all_single.
usage: torchrun --nproc-per-node G tools/bench_exchange.py [L]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from dna_b200 import _lib
from dna_b200.dp import ChannelPartition

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
D = 256
Lc = L // world
lib = _lib.lib()
cases = [("uT bf16 [3D, L/G] -> [3w, L]", (1, 3 * D, Lc), torch.bfloat16, 3, True),
         ("z  bf16 [w, L] -> [D, L/G]", (1, D // world, L), torch.bfloat16, 1, False),
         ("k  fp32 [D, L/G] -> [w, L]", (1, D, Lc), torch.float32, 1, True)]
for backend, mode in (("nccl", 0), ("peer", 0), ("peer", 1)):
    part = ChannelPartition(backend=backend)
    lib.hy_debug_set_peer_mode(mode)
    for name, shape, dt, n, to_ch in cases:
        x = torch.randn(shape, device=dev).to(dt)
        ref = None
        for _ in range(3):
            y = part._exchange(x, n, to_ch)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            y = part._exchange(x, n, to_ch)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # cross-check against the NCCL result
        chk = ChannelPartition(backend="nccl")._exchange(x, n, to_ch)
        ok = torch.equal(chk, y)
        if rank == 0:
            nbytes = x.numel() * x.element_size()
            print(f"{backend:4s} mode {mode}  {name:32s} {float(t):7.3f} ms  out {nbytes / 1e6:7.1f} MB  remote {nbytes * (world - 1) / world / float(t) / 1e6:7.1f} GB/s  equal_to_nccl={ok}", flush=True)
    if backend == "peer":
        part.peer(dev).check()
dist.barrier(); dist.destroy_process_group()
