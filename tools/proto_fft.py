"""Numpy prototype of the index algebra the CUDA kernels use (design aid, not product code).

Validates: (1) in-place mixed-radix DIF (natural in -> digit-reversed out) and its DIT inverse,
(2) real-sequence packing N real -> M=N/2 complex with on-the-fly unpack/multiply/repack on
(k, M-k) pairs, (3) the M = M1 x M2 four-step split with row pairing (k1, M1-k1).
"""
import numpy as np

def plan(S):
    r = []
    while S > 1:
        for c in (16, 8, 4, 2):
            if S % c == 0 and S >= c:
                r.append(c); S //= c; break
    return r

def dif_inplace(x, radices):
    """x: [..., S] complex. natural in -> digit-reversed out (position p holds freq(p))."""
    x = x.copy(); S = x.shape[-1]; P = S
    for R in radices:
        sub = P // R
        y = x.reshape(x.shape[:-1] + (S // P, R, sub))        # [blk, m, j]
        F = np.exp(-2j * np.pi * np.outer(np.arange(R), np.arange(R)) / R)  # [q, m]
        z = np.einsum('qm,...bmj->...bqj', F, y)
        tw = np.exp(-2j * np.pi * np.outer(np.arange(R), np.arange(sub)) / P)  # [q, j]
        z = z * tw
        x = z.reshape(x.shape)
        P = sub
    return x

def dit_inplace_inv(x, radices):
    """inverse of dif_inplace (unnormalised: returns S * original)."""
    x = x.copy(); S = x.shape[-1]
    spans = []
    P = S
    for R in radices:
        spans.append((P, R)); P //= R
    for (P, R) in reversed(spans):
        sub = P // R
        y = x.reshape(x.shape[:-1] + (S // P, R, sub))        # [blk, q, j]
        tw = np.exp(+2j * np.pi * np.outer(np.arange(R), np.arange(sub)) / P)
        y = y * tw
        F = np.exp(+2j * np.pi * np.outer(np.arange(R), np.arange(R)) / R)  # [m, q]
        z = np.einsum('mq,...bqj->...bmj', F, y)
        x = z.reshape(x.shape)
    return x

def freq_of_pos(p, radices):
    """frequency index stored at position p after dif_inplace."""
    S = int(np.prod(radices)); k = 0; w = 1; rem = S
    for R in radices:
        rem //= R
        q = (p // rem) % R
        k += q * w; w *= R
    return k

def pos_of_freq(k, radices):
    S = int(np.prod(radices)); p = 0; rem = S
    for R in radices:
        rem //= R
        q = k % R; k //= R
        p += q * rem
    return p

def pair_pointwise(Zk, Zm, Kk, Km, W):
    """packed-domain multiply for the pair (k, M-k), k != 0.  W = exp(-2*pi*i*k/N).
    Returns (Wk, Wm) = packed spectrum of the product y at k and M-k."""
    E = 0.5 * (Zk + np.conj(Zm)); O = 0.5 * (Zk - np.conj(Zm)); T = W * O
    Xk = E - 1j * T
    Xm = np.conj(E + 1j * T)
    Yk = Kk * Xk; Ym = Km * Xm
    E2 = 0.5 * (Yk + np.conj(Ym)); O2 = 0.5 * (Yk - np.conj(Ym)); T2 = np.conj(W) * O2
    return E2 + 1j * T2, np.conj(E2 - 1j * T2)

def conv_packed_single(g, k, Mc):
    """causal conv y[t]=sum_s k[s] g[t-s] via one M-point complex FFT of the packed sequence."""
    L = len(g); N = 2 * Mc
    rad = plan(Mc)
    gp = np.zeros(N); gp[:L] = g
    z = gp[0::2] + 1j * gp[1::2]
    Z = dif_inplace(z, rad)
    kp = np.zeros(N); kp[:L] = k
    K = np.fft.rfft(kp)                     # true spectrum K[0..M]
    Wout = np.zeros(Mc, complex)
    for p in range(Mc):
        f = freq_of_pos(p, rad)
        if f == 0:
            X0 = Z[p].real + Z[p].imag; XM = Z[p].real - Z[p].imag
            Y0 = K[0].real * X0; YM = K[Mc].real * XM
            Wout[p] = 0.5 * (Y0 + YM) + 0.5j * (Y0 - YM)
        else:
            pm = pos_of_freq(Mc - f, rad)
            Wtw = np.exp(-2j * np.pi * f / N)
            a, b = pair_pointwise(Z[p], Z[pm], K[f], K[Mc - f], Wtw)
            Wout[p] = a
    w = dit_inplace_inv(Wout, rad) / Mc
    y = np.empty(N); y[0::2] = w.real; y[1::2] = w.imag
    return y[:L]

def conv_packed_2d(g, k, M1, M2):
    """four-step: n = n1*M2 + n2, k = k1 + M1*k2; rows k1 paired with M1-k1."""
    L = len(g); Mc = M1 * M2; N = 2 * Mc
    r1 = plan(M1); r2 = plan(M2)
    gp = np.zeros(N); gp[:L] = g
    z = (gp[0::2] + 1j * gp[1::2]).reshape(M1, M2)
    # phase A: column FFTs over n1 (in place, k1 at pos1), twiddle W_M^{n2*k1}
    A = dif_inplace(z.T.copy(), r1).T.copy()           # [pos1, n2]
    k1_of = np.array([freq_of_pos(p, r1) for p in range(M1)])
    A = A * np.exp(-2j * np.pi * np.outer(k1_of, np.arange(M2)) / Mc)
    # phase B: row FFTs over n2 (k2 at pos2)
    Bm = dif_inplace(A, r2)                            # [pos1, pos2]
    kp = np.zeros(N); kp[:L] = k
    K = np.fft.rfft(kp)
    k2_of = np.array([freq_of_pos(p, r2) for p in range(M2)])
    Wm = np.zeros_like(Bm)
    for p1 in range(M1):
        k1 = k1_of[p1]
        for p2 in range(M2):
            k2 = k2_of[p2]; f = k1 + M1 * k2
            if f == 0:
                Z0 = Bm[p1, p2]
                X0 = Z0.real + Z0.imag; XM = Z0.real - Z0.imag
                Y0 = K[0].real * X0; YM = K[Mc].real * XM
                Wm[p1, p2] = 0.5 * (Y0 + YM) + 0.5j * (Y0 - YM); continue
            fm = Mc - f
            k1m = fm % M1; k2m = fm // M1
            if k1 != 0:
                assert k1m == M1 - k1 and k2m == M2 - 1 - k2
                assert pos_of_freq(k2m, r2) == M2 - 1 - p2
            else:
                assert k1m == 0 and k2m == (M2 - k2) % M2
            q1 = pos_of_freq(k1m, r1); q2 = pos_of_freq(k2m, r2)
            Wtw = np.exp(-2j * np.pi * f / N)
            a, b = pair_pointwise(Bm[p1, p2], Bm[q1, q2], K[f], K[fm], Wtw)
            Wm[p1, p2] = a
    # inverse row FFTs, conj twiddle, inverse column FFTs
    C = dit_inplace_inv(Wm, r2)
    C = C * np.exp(+2j * np.pi * np.outer(k1_of, np.arange(M2)) / Mc)
    w = dit_inplace_inv(C.T.copy(), r1).T.reshape(-1) / Mc
    y = np.empty(N); y[0::2] = w.real; y[1::2] = w.imag
    return y[:L]

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for S in (16, 64, 128, 512, 1024, 4096):
        x = rng.standard_normal(S) + 1j * rng.standard_normal(S)
        rad = plan(S)
        X = dif_inplace(x, rad)
        ref = np.fft.fft(x)
        f = np.array([freq_of_pos(p, rad) for p in range(S)])
        assert np.allclose(X, ref[f]), S
        assert all(pos_of_freq(freq_of_pos(p, rad), rad) == p for p in range(S))
        assert np.allclose(dit_inplace_inv(X, rad) / S, x)
    for (L, Mc) in ((100, 128), (128, 128), (1000, 1024), (37, 64)):
        g = rng.standard_normal(L); k = rng.standard_normal(L)
        ref = np.convolve(g, k)[:L]
        assert np.allclose(conv_packed_single(g, k, Mc), ref), (L, Mc)
    for (L, M1, M2) in ((1000, 4, 256), (2048, 8, 256), (512, 2, 256), (4000, 16, 256), (8192, 32, 256), (3000, 64, 64)):
        g = rng.standard_normal(L); k = rng.standard_normal(L)
        ref = np.convolve(g, k)[:L]
        out = conv_packed_2d(g, k, M1, M2)
        assert np.allclose(out, ref), (L, M1, M2, np.abs(out - ref).max())
    print("proto ok")
