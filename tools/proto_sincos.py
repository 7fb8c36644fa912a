"""Numerical check (numpy, fp32 arithmetic emulated through fp64 FMAs) of hy_sincos_core in dna_b200/csrc/hy_common.cuh:
three-term Cody-Waite reduction by pi/2 + Cephes-style polynomials, against fp64 sin / cos up to |x| = 1e5.
    python tools/proto_sincos.py"""
import numpy as np
f32 = np.float32
hi = f32(np.pi / 2); mid = f32(np.pi / 2 - float(hi)); lo = f32(np.pi / 2 - float(hi) - float(mid))
print("pi/2 =", repr(float(hi)), "+", repr(float(mid)), "+", repr(float(lo)))
fma = lambda a, b, c: (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)
k = lambda r, v: np.full_like(r, f32(v))


def sincos(x):
    x = x.astype(f32)
    j = np.rint((x * f32(0.636619772367581343)).astype(f32)).astype(f32)
    r = fma(j, k(j, -hi), x); r = fma(j, k(j, -mid), r); r = fma(j, k(j, -lo), r)
    q = j.astype(np.int64)
    r2 = (r * r).astype(f32)
    ps = fma(r2, k(r, -1.9515295891e-4), k(r, 8.3321608736e-3)); ps = fma(ps, r2, k(r, -1.6666654611e-1)); ps = fma((ps * r2).astype(f32), r, r)
    pc = fma(r2, k(r, 2.443315711809948e-5), k(r, -1.388731625493765e-3)); pc = fma(pc, r2, k(r, 4.166664568298827e-2))
    pc = fma(pc, r2, k(r, -0.5)); pc = fma(pc, r2, k(r, 1.0))
    sw = (q & 1) == 1
    ss, cc = np.where(sw, pc, ps), np.where(sw, ps, pc)
    return np.where((q & 2) != 0, -ss, ss), np.where(((q + 1) & 2) != 0, -cc, cc)


rng = np.random.default_rng(0)
for scale in (3, 50, 300, 3e4, 1e5):
    x = rng.uniform(-scale, scale, 2_000_000).astype(f32)
    s, c = sincos(x)
    x64 = x.astype(np.float64)
    print(f"|x| < {scale:g}: max abs err sin {np.abs(s - np.sin(x64)).max():.2e}  cos {np.abs(c - np.cos(x64)).max():.2e}  "
          f"(numpy float32 sin: {np.abs(np.sin(x).astype(np.float64) - np.sin(x64)).max():.2e})")
