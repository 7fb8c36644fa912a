"""Schedule statistics of the persistent long-conv pipeline (hy_conv_pipe.cuh): per C-ABI call, how many cycles the
CTAs spent waiting on a dependency (per phase), inside the phase bodies, and in claim + publish.
usage: python tools/prof_pipe.py L H [dtype]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dna_b200 import kernels as K, _lib
from dna_b200._lib import IN_SHORTCONV, OUT_SHORTCONV

L = int(sys.argv[1]); H = int(sys.argv[2]); B = 1
dt = torch.bfloat16 if (len(sys.argv) < 4 or sys.argv[3] == "bf16") else torch.float32
dev = "cuda"
lib = _lib.lib()
torch.manual_seed(0)
uT = torch.randn(B, 3 * H, L, device=dev).to(dt)
sw = torch.randn(3 * H, 3, device=dev) * 0.5
sb = torch.randn(3 * H, device=dev); pb = torch.randn(3 * H, device=dev)
k = torch.randn(H, L, device=dev) * torch.exp(-torch.arange(L, device=dev) / (L / 4.0))[None]
D = torch.randn(H, device=dev)
dz = torch.randn(B, H, L, device=dev).to(dt)
stats = (ctypes.c_ulonglong * 16)()


def timed(name, fn):
    for _ in range(2):
        out = fn()
    lib.hy_debug_pipe_stats(1, None)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record()
    torch.cuda.synchronize()
    lib.hy_debug_pipe_stats(0, stats)
    s = list(stats)
    ms = e0.elapsed_time(e1)
    items = max(s[6], 1)
    print(f"{name:9s} {ms:7.3f} ms  items {s[6]:6d}  body {s[7] / items:8.0f} cyc/item  claim+publish {s[8] / items:6.0f} cyc/item  "
          f"waits: A {s[1]:5d} items {s[0] / max(s[1], 1):7.0f} cyc | B {s[3]:5d} items {s[2] / max(s[3], 1):7.0f} cyc | "
          f"C {s[5]:5d} items {s[4] / max(s[5], 1):7.0f} cyc | wait share of CTA time {100 * (s[0] + s[2] + s[4]) / max(s[0] + s[2] + s[4] + s[7] + s[8], 1):.1f}%",
          flush=True)
    return out


def family():
    Kf = timed("spectrum", lambda: K.filter_spectrum(k, D, L))
    gs = K.conv_gsave_alloc(B, H, L, dev)
    z, ys = timed("conv_fwd", lambda: K.conv_fwd(uT, Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=sw, sb=sb, pb=pb, save_y=True, gsave=gs))
    defer = K.shortconv_gate_supported(uT, dz, ys)
    res = timed("conv_bwd", lambda: K.conv_bwd(dz, uT, Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=sw, sb=sb, pb=pb, ysave=ys, gsave=gs, defer_dx0=defer))
    timed("conv_dk", lambda: K.conv_dk(res[3], L))


lib.hy_debug_set_conv_pipe(0)
print("--- per-phase launches over row groups (pipeline off)")
family()
lib.hy_debug_set_conv_pipe(1)
for lag in (0, 1, 2, 3, 4):
    lib.hy_debug_set_pipe_lag(lag)
    print(f"--- persistent pipeline, lag {lag} (0 = automatic)")
    family()
sys.exit(0)
Kf = timed("spectrum", lambda: K.filter_spectrum(k, D, L))
gs = K.conv_gsave_alloc(B, H, L, dev)
z, ys = timed("conv_fwd", lambda: K.conv_fwd(uT, Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=sw, sb=sb, pb=pb, save_y=True, gsave=gs))
defer = K.shortconv_gate_supported(uT, dz, ys)
res = timed("conv_bwd", lambda: K.conv_bwd(dz, uT, Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=sw, sb=sb, pb=pb, ysave=ys, gsave=gs, defer_dx0=defer))
timed("conv_dk", lambda: K.conv_dk(res[3], L))
