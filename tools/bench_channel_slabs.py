"""BASELINE.json configs[3], the operator view: ONE 1 M-nt sequence, the fused long-conv + gating operator
(fwd + bwd) with its d_model channels partitioned over the ranks (dna_b200.dp.channel_slab, SURVEY 8e "Channels,
B = 1"): the long conv, the gates, the D skip and the filter gradient are per channel, so there is NO collective
on this path — strong scaling of one sequence, time = max over ranks (CUDA events on the launching stream).

    python tools/bench_channel_slabs.py [--steps 5] [--warmup 3] [--seqlen 1000000] [--d-model 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_channel_slabs.py

Rank 0 prints one JSON line: nt/s of the whole sequence, algorithmic GB/s (11*s*B*D*L + 12*D*L over all channels,
DESIGN section 5) summed over ranks, and its fraction of N x the measured HBM peak.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from dna_b200.dp import channel_slab
from dna_b200.fftconv import fftconv_func

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--seqlen", type=int, default=1_000_000)
ap.add_argument("--d-model", type=int, default=256)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
L, Dm = args.seqlen, args.d_model
lo, hi = channel_slab(Dm, rank, world)
H = hi - lo
g = torch.Generator(device=dev).manual_seed(1234)          # the same full tensors on every rank, sliced to the slab
mk = lambda *s: torch.randn(*s, device=dev, generator=g)
x0, x1, v, dz = (mk(1, Dm, L)[:, lo:hi].to(torch.bfloat16).contiguous() for _ in range(4))
k = (mk(Dm, L) * torch.exp(-torch.arange(L, device=dev) / (L / 8.0))[None] / 8)[lo:hi].contiguous()
D = mk(Dm)[lo:hi].contiguous()
inputs = [t.requires_grad_(True) for t in (x0, x1, v, k, D)]


def step():
    for t in inputs:
        t.grad = None
    out = fftconv_func(x1, k, D, dropout_mask=None, gelu=False, v=v, q=x0)
    out.backward(dz)
    return out


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    out = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
chk = torch.stack([out.float().abs().sum(), inputs[3].grad.abs().sum()])   # a checksum of the slab's results
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(chk)
if rank == 0:
    peak = 6545.0
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    t = float(ms)
    alg = 11 * 2 * Dm * L + 12 * Dm * L
    print(json.dumps({"metric": "nucleotides/sec fwd+bwd fused long-conv + gating operator @1M bp, channel-partitioned",
                      "value": L / t * 1e3, "unit": "nt/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": t, "scaling": "strong", "dtype": "bf16 activations / fp32 FFT", "data": "synthetic",
                      "config": {"workload": "longconv-1m-channel-slabs", "seqlen": L, "d_model": Dm, "batch": 1,
                                 "channels_per_gpu": H, "collectives": "none"},
                      "roofline": {"bound": "hbm", "achieved": alg / t / 1e6, "peak": peak * world, "unit": "GB/s",
                                   "frac": alg / t / 1e6 / (peak * world), "algorithmic_bytes": alg},
                      "checksum": [float(c) for c in chk]}))
if world > 1:
    dist.destroy_process_group()
