"""Micro-benchmark of the long-conv entry points (CUDA events) — also the ncu target.
usage: python tools/prof_conv.py L H B [dtype] [iters]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dna_b200 import kernels as K
from dna_b200._lib import IN_SHORTCONV, OUT_SHORTCONV

L = int(sys.argv[1]); H = int(sys.argv[2]); B = int(sys.argv[3])
dt = torch.bfloat16 if (len(sys.argv) < 5 or sys.argv[4] == "bf16") else torch.float32
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dev = "cuda"
if os.environ.get('HY_L2_MB') or os.environ.get('HY_NSTREAM'):
    from dna_b200 import _lib
    _lib.lib().hy_set_pipeline(int(os.environ.get('HY_NSTREAM', '1')), int(os.environ.get('HY_L2_MB', '1024')) << 20)
torch.manual_seed(0)
uT = torch.randn(B, 3 * H, L, device=dev).to(dt)
sw = torch.randn(3 * H, 3, device=dev) * 0.5
sb = torch.randn(3 * H, device=dev)
pb = torch.randn(3 * H, device=dev)
k = torch.randn(H, L, device=dev) * torch.exp(-torch.arange(L, device=dev) / (L / 4.0))[None]
D = torch.randn(H, device=dev)
dz = torch.randn(B, H, L, device=dev).to(dt)
s = 2 if dt == torch.bfloat16 else 4
if os.environ.get('HY_PERSIST'):
    from dna_b200 import _lib
    import ctypes
    print('persisting L2 MB:', _lib.lib().hy_debug_set_persist(1))

def run():
    Kf = K.filter_spectrum(k, D, L)
    gs = None if os.environ.get('HY_NO_GSAVE') else K.conv_gsave_alloc(B, H, L, dev)
    z, ys = K.conv_fwd(uT, Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=sw, sb=sb, pb=pb, save_y=True, gsave=gs)
    defer = (not os.environ.get('HY_NO_DEFER')) and K.shortconv_gate_supported(uT, dz, ys)   # as dna_b200.hyena does
    dX, _, _, dKacc, dD = K.conv_bwd(dz, uT, Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=sw, sb=sb, pb=pb, ysave=ys,
                                     gsave=gs, defer_dx0=defer)
    dk = K.conv_dk(dKacc, L)
    duT = K.shortconv_bwd(uT, dX, sw, pb, L, dout=dz if defer else None, ysave=ys if defer else None)
    return z

for _ in range(2):
    run()
torch.cuda.synchronize()
K.enable_timing(True); K.drain_timing()
n0 = K.launch_count()
for _ in range(iters):
    run()
t = K.drain_timing()
nl = (K.launch_count() - n0) / iters
alg = 11 * s * B * H * L + 12 * H * L
tot = 0.0
for tag, (c, ms) in t.items():
    print(f"{tag:14s} {ms / iters:9.3f} ms/iter")
    if tag in ("spectrum", "conv_fwd", "conv_bwd", "conv_dk"):
        tot += ms / iters
print(f"L={L} H={H} B={B} {dt}: long-conv family {tot:.3f} ms/iter, {alg / tot / 1e6:.1f} GB/s algorithmic "
      f"({alg / tot / 1e6 / 6545 * 100:.1f}% of 6545), {tot * 1e3 / (B * H):.1f} us/row, launches/iter {nl:.0f}")
