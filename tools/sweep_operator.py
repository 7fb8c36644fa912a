"""BASELINE.json configs[4]: isolated fftconv + gating operator sweep, seqlen 1 k .. 1 M x d_model 128 .. 512,
our fused kernels against the reference's own GPU path (`fftconv_ref`: torch.fft / cuFFT + elementwise gates,
src/models/sequence/hyena.py:60-92 with the gates of hyena.py:481,496) on the SAME device and inputs.

    python tools/sweep_operator.py [out.jsonl] [--max-elems 27] [--dtype bf16|fp32]

Per (L, D): B is chosen so that B*D*L ~ 2^max_elems (SURVEY 8d uses 2^28; 2^27 keeps the reference path's autograd
temporaries comfortable).  Times are CUDA-event means of fwd+bwd after warm-up; parity is the max-abs difference
of the outputs / gradients relative to the reference's output scale (printed, asserted only loosely here —
the parity gates live in tests/).  `k` is a decaying random filter, `D ~ N(0,1)`.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dna_b200.fftconv import fftconv_func


def fftconv_ref(u, k, D, dropout_mask=None, gelu=False):
    """The GPU baseline being timed: the reference's eager long convolution (torch.fft / cuFFT), restated from
    src/models/sequence/hyena.py:60-92 (fft_size = 2L, fp32 FFT, cast back to u.dtype). Measurement-only."""
    L = u.shape[-1]
    n = 2 * L
    k_f = torch.fft.rfft(k, n=n) / n
    u_f = torch.fft.rfft(u.to(k.dtype), n=n)
    y = torch.fft.irfft(u_f * k_f, n=n, norm="forward")[..., :L]
    return (y + u * D.unsqueeze(-1)).to(u.dtype)

args = sys.argv[1:]
out_path = args[0] if args and not args[0].startswith("--") else None
max_elems = int(args[args.index("--max-elems") + 1]) if "--max-elems" in args else 27
dt = torch.float32 if ("--dtype" in args and args[args.index("--dtype") + 1] == "fp32") else torch.bfloat16
dev = "cuda"
PEAK = 6545.0
if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", PEAK)


def ref_op(x0, x1, v, k, D):
    g = v * x1                                             # gate #1 (hyena.py:481)
    y = fftconv_ref(g, k, D, dropout_mask=None, gelu=False)  # long conv + skip (hyena.py:60-92; gelu off, l.265)
    return y * x0                                          # gate #2 (hyena.py:496)


def our_op(x0, x1, v, k, D):
    return fftconv_func(x1, k, D, dropout_mask=None, gelu=False, v=v, q=x0)


def timed(fn, inputs, dz, iters):
    for _ in range(2):
        for t in inputs:
            t.grad = None
        fn(*inputs).backward(dz)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        for t in inputs:
            t.grad = None
        out = fn(*inputs)
        out.backward(dz)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out.detach(), [t.grad.detach().clone() for t in inputs]


rows = []
print(f"{'L':>8} {'D':>4} {'B':>5} {'ours ms':>9} {'ref ms':>9} {'x':>6} {'alg GB/s':>9} {'frac':>6}  max rel diff (out, dx0, dx1, dv, dk, dD)")
for L in (1024, 4096, 16384, 65536, 262144, 1_000_000):
    for Dm in (128, 256, 512):
        B = max(1, (1 << max_elems) // (Dm * L))
        torch.manual_seed(0)
        mk = lambda: torch.randn(B, Dm, L, device=dev).to(dt).requires_grad_(True)
        x0, x1, v = mk(), mk(), mk()
        k = (torch.randn(Dm, L, device=dev) * torch.exp(-torch.arange(L, device=dev) / (L / 8.0))[None] / 8).requires_grad_(True)
        D = torch.randn(Dm, device=dev).requires_grad_(True)
        dz = torch.randn(B, Dm, L, device=dev).to(dt)
        inputs = [x0, x1, v, k, D]
        iters = 5 if B * Dm * L >= (1 << 26) else 20
        t_ours, o_ours, g_ours = timed(our_op, inputs, dz, iters)
        t_ref, o_ref, g_ref = timed(ref_op, inputs, dz, max(2, iters // 2))
        rel = lambda a, b: float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))
        diffs = [rel(o_ours, o_ref)] + [rel(a, b) for a, b in zip(g_ours, g_ref)]
        s = 2 if dt == torch.bfloat16 else 4
        alg = 11 * s * B * Dm * L + 12 * Dm * L
        gbs = alg / t_ours / 1e6
        row = dict(L=L, D=Dm, B=B, dtype=str(dt).split(".")[-1], ours_ms=t_ours, ref_ms=t_ref, speedup=t_ref / t_ours,
                   alg_bytes=alg, alg_gbs=gbs, frac_of_hbm=gbs / PEAK, rel_diff=dict(zip(["out", "dx0", "dx1", "dv", "dk", "dD"], diffs)))
        rows.append(row)
        print(f"{L:8d} {Dm:4d} {B:5d} {t_ours:9.3f} {t_ref:9.3f} {t_ref / t_ours:6.2f} {gbs:9.1f} {gbs / PEAK:6.3f}  "
              + " ".join(f"{d:.1e}" for d in diffs), flush=True)
        del x0, x1, v, k, D, dz, inputs, o_ours, o_ref, g_ours, g_ref
        torch.cuda.empty_cache()
if out_path:
    with open(out_path, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
