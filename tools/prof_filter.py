"""Micro-benchmark of the implicit-filter kernels (CUDA events).  usage: python tools/prof_filter.py L D [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dna_b200 import kernels as K, _lib
from dna_b200.hyena import HyenaFilter

L = int(sys.argv[1]); D = int(sys.argv[2]); iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
if os.environ.get("HY_TRUNK_MINB"):
    print("trunk minb:", _lib.lib().hy_debug_set_trunk_minb(int(os.environ["HY_TRUNK_MINB"])))
if os.environ.get("HY_FWD_TF"):
    print("fwd tf:", _lib.lib().hy_debug_set_filter_fwd_tf(int(os.environ["HY_FWD_TF"])))
torch.manual_seed(0)
f = HyenaFilter(D, emb_dim=5, order=64, seq_len=L, w=10, lr_pos_emb=0.0).cuda()
dk = torch.randn(D, L, device="cuda")

def run():
    for p in f.parameters():
        p.grad = None
    k = f.filter_cm(L)          # [D, L] channel-major, as the operator consumes it
    k.backward(dk)

for _ in range(2):
    run()
torch.cuda.synchronize()
K.enable_timing(True); K.drain_timing()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    run()
e1.record()
t = K.drain_timing()
for tag, (c, ms) in t.items():
    print(f"{tag:14s} {ms / iters:9.3f} ms/iter ({c // iters} calls)")
print(f"L={L} D={D}: filter fwd+bwd wall {e0.elapsed_time(e1) / iters:.3f} ms/iter")
