"""Micro-benchmark of the add + LayerNorm kernels.  usage: python tools/prof_addln.py rows D [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dna_b200 import kernels as K

rows = int(sys.argv[1]); D = int(sys.argv[2]); iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
x = torch.randn(rows, D, device="cuda").bfloat16(); r = torch.randn(rows, D, device="cuda")
g = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
dy = torch.randn(rows, D, device="cuda").bfloat16(); dr = torch.randn(rows, D, device="cuda")
def run():
    y, ro, mean, rstd = K.add_ln_fwd(x, r, g, b, 1e-5, torch.bfloat16, torch.float32, True)
    K.add_ln_bwd(dy, dr, ro, mean, rstd, g, torch.bfloat16, True, True)
for _ in range(3): run()
torch.cuda.synchronize(); K.enable_timing(True); K.drain_timing()
for _ in range(iters): run()
t = K.drain_timing()
n = rows * D
for tag, byt in (("add_ln_fwd", 12 * n), ("add_ln_bwd", 16 * n)):
    ms = t[tag][1] / iters
    print(f"{tag}: {ms:.3f} ms, {byt / ms / 1e6:.0f} GB/s ({byt / ms / 1e6 / 6545 * 100:.0f}% of 6545)")
