"""Long-conv family (spectrum + forward + backward + dk, CUDA events) over a list of sequence lengths — used to compare
column-tile choices of the 3 * 2^a / 5 * 2^a transform lengths and the padded power-of-two lengths.
usage: python tools/sweep_lengths.py H L1,L2,... [iters]     (HYENA_B200_LIB / HYENA_B200_POW2_ONLY select the variant)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dna_b200 import kernels as K
from dna_b200._lib import IN_SHORTCONV, OUT_SHORTCONV

H = int(sys.argv[1]); Ls = [int(x) for x in sys.argv[2].split(",")]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev, dt, B = "cuda", torch.bfloat16, 1
for L in Ls:
    torch.manual_seed(0)
    uT = torch.randn(B, 3 * H, L, device=dev).to(dt)
    sw = torch.randn(3 * H, 3, device=dev) * 0.5
    sb = torch.randn(3 * H, device=dev); pb = torch.randn(3 * H, device=dev)
    k = torch.randn(H, L, device=dev) * torch.exp(-torch.arange(L, device=dev) / (L / 4.0))[None]
    D = torch.randn(H, device=dev)
    dz = torch.randn(B, H, L, device=dev).to(dt)

    def run():
        Kf = K.filter_spectrum(k, D, L)
        gs = K.conv_gsave_alloc(B, H, L, dev)
        z, ys = K.conv_fwd(uT, Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=sw, sb=sb, pb=pb, save_y=True, gsave=gs)
        defer = K.shortconv_gate_supported(uT, dz, ys)
        dX, _, _, dKacc, dD = K.conv_bwd(dz, uT, Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=sw, sb=sb, pb=pb,
                                         ysave=ys, gsave=gs, defer_dx0=defer)
        return K.conv_dk(dKacc, L)

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    K.enable_timing(True); K.drain_timing()
    for _ in range(iters):
        run()
    t = K.drain_timing()
    K.enable_timing(False)
    parts = {tag: ms / iters for tag, (c, ms) in t.items() if tag in ("spectrum", "conv_fwd", "conv_bwd", "conv_dk")}
    tot = sum(parts.values())
    M = K.fft_len(L)
    alg = 11 * 2 * B * H * L + 12 * H * L
    print(f"L={L:8d} M={M:8d} (M1={M // 4096 if M > 4096 else 1:3d}) H={H}: family {tot:7.3f} ms  "
          f"{alg / tot / 1e6:7.1f} GB/s alg ({alg / tot / 1e6 / 65.45:4.1f}%)  " + "  ".join(f"{k_} {v:.3f}" for k_, v in parts.items()), flush=True)
