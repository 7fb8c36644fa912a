"""One fwd+bwd of the gated operator in the single-kernel regime (L <= 4096), timed per C-ABI call; run under ncu with
-k regex:k_fused to capture the kernels.  usage: python tools/prof_fused.py L D B [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dna_b200 import kernels as K
from dna_b200.fftconv import fftconv_func

L, D, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = "cuda"
torch.manual_seed(0)
x0, x1, v = (torch.randn(B, D, L, device=dev).to(torch.bfloat16).requires_grad_(True) for _ in range(3))
k = (torch.randn(D, L, device=dev) * torch.exp(-torch.arange(L, device=dev) / (L / 4.0))).requires_grad_(True)
Dp = torch.randn(D, device=dev).requires_grad_(True)
dz = torch.randn(B, D, L, device=dev).to(torch.bfloat16)
for _ in range(2):
    fftconv_func(x1, k, Dp, None, False, v=v, q=x0).backward(dz)
torch.cuda.synchronize()
K.enable_timing(True); K.drain_timing()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fftconv_func(x1, k, Dp, None, False, v=v, q=x0).backward(dz)
e1.record(); torch.cuda.synchronize()
kt = K.drain_timing()
ms = e0.elapsed_time(e1) / iters
alg = 11 * 2 * B * D * L + 12 * D * L
print(f"L={L} D={D} B={B}: {ms:.3f} ms fwd+bwd, {alg / ms / 1e6:.1f} GB/s algorithmic; per call (ms):",
      {t: round(kt[t][1] / iters, 3) for t in kt})
