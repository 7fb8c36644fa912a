"""Drop-in for the reference's `src/ops/fftconv.py` (call surface: fftconv_ref, fftconv_h3_ref,
FFTConvFunc, fftconv_func) whose compiled backend `fftconv` (fftconv_fwd / fftconv_bwd, imported at
/root/reference/src/ops/fftconv.py:8) is absent from the reference tree.

`fftconv_func` runs the hand-written sm_100a kernels of libhyena_b200.so through the C-ABI
(include/hyena_b200.h); it takes CUDA tensors only and raises otherwise — there is no CPU fallback.
`fftconv_ref` / `fftconv_h3_ref` / `fftconv_heads_ref` keep the names `src/models/sequence/hyena.py:12-17`
imports from this module, with the reference's argument order, and run on the same kernels: the package
contains no torch.fft path at all (the eager restatement used for checking lives in oracle/).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import kernels as K
from ._lib import IN_PLAIN, IN_PREGATE, OUT_PLAIN, OUT_POSTGATE

__all__ = ["fftconv_ref", "fftconv_h3_ref", "fftconv_heads_ref", "FFTConvFunc", "fftconv_func", "circular_conv"]


def fftconv_ref(u, k, D, dropout_mask=None, gelu=True, k_rev=None, bidirectional=False):
    """Same name and argument list as the reference's eager definition (src/ops/fftconv.py:15-34,
    src/models/sequence/hyena.py:60-92) — `hyena.py:12-17` imports it from this module — but executed by the
    sm_100a kernels: it is `fftconv_func` with the reference's positional order.  (The reference's own torch.fft
    restatement lives in oracle/, which the package never imports.)"""
    return fftconv_func(u, k, D, dropout_mask, gelu, k_rev=k_rev, bidirectional=bidirectional)


def fftconv_h3_ref(k, ssm_kernel, D, q, v, head_dim=1, ssm_kernel_rev=None):
    """H3 form (src/ops/fftconv.py:38-55): out = sum_d1 (conv(ssm_kernel, k (x) v) + D k (x) v) * q, on the kernels."""
    return fftconv_func(k, ssm_kernel, D, None, False, v=v, head_dim=head_dim, q=q, k_rev=ssm_kernel_rev)


def fftconv_heads_ref(u, k, D, dropout_mask=None, gelu=True, k_rev=None, bidirectional=False):
    """`hyena.py:13` imports this name; the reference tree never defined it.  Same contraction as
    fftconv_ref for the 5-D "b h v z l" layout HyenaOperator uses."""
    return fftconv_ref(u, k, D, dropout_mask, gelu, k_rev, bidirectional)


def _rows(t):
    """[B, H, L] view with unit stride along L (copy only if the caller handed a strided tensor)."""
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t


class FFTConvFunc(torch.autograd.Function):
    """Same argument list as the reference's FFTConvFunc (src/ops/fftconv.py:58-103).

    Differences that are implementation, not interface: the FFT length is
    `hy_fft_len(L) * 2 >= 2L` (a power of two as at :64, or 3 / 5 times one for L > 4096), the spectrum of k is computed by our own
    kernel instead of torch.fft.rfft (:65), D is folded into that spectrum, and the backward
    recomputes spectra instead of saving `k_f`.
    """

    @staticmethod
    def forward(ctx, u, k, D, dropout_mask=None, gelu=True, force_fp16_output=False, output_hbl_layout=False,
                v=None, head_dim=1, q=None, fftfp16=False, k_rev=None, bidirectional=False):
        if bidirectional or head_dim != 1 or k_rev is not None or output_hbl_layout:
            raise RuntimeError("FFTConvFunc is the causal, head_dim = 1 core; call fftconv_func for the other keywords")
        K._check_dev(u)  # CUDA tensors only: hyena-b200 has no CPU fallback
        in_dtype = u.dtype
        cdt = torch.bfloat16 if in_dtype == torch.bfloat16 else torch.float32
        L = u.shape[-1]
        shape = u.shape
        kf32 = k.detach().to(torch.float32).reshape(-1, k.shape[-1])     # [H, L] (H = all channel axes of u)
        H = kf32.shape[0]
        u3 = _rows(u.reshape(-1, H, L).to(cdt))
        if kf32.stride(-1) != 1:
            kf32 = kf32.contiguous()
        D32 = D.detach().to(torch.float32).reshape(-1).contiguous()
        Kf = K.filter_spectrum(kf32, D32, L)
        v3 = q3 = None
        if v is not None:
            v3 = v.reshape(u3.shape).to(cdt)
            if v3.stride() != u3.stride():
                v3 = v3.contiguous()
                u3 = u3.contiguous()
        if q is not None:
            q3 = _rows(q.reshape(u3.shape).to(cdt))
        in_mode = IN_PREGATE if v3 is not None else IN_PLAIN
        out_mode = OUT_POSTGATE if q3 is not None else OUT_PLAIN
        # long sequences: keep the spectrum of the (gated) input so the backward transforms dout only
        gs = K.conv_gsave_alloc(u3.shape[0], u3.shape[1], L, u3.device) if any(ctx.needs_input_grad) else None
        y, ys = K.conv_fwd(u3, Kf, L, in_mode=in_mode, out_mode=out_mode, pre=v3, post=q3, save_y=q3 is not None, gsave=gs)
        ctx.modes = (in_mode, out_mode)
        ctx.L = L
        ctx.gelu = gelu
        ctx.shapes = (shape, k.shape, D.shape, in_dtype)
        ctx.has_mask = dropout_mask is not None
        out = y
        pre_act = None
        if gelu or dropout_mask is not None:
            # post-ops of the reference (src/ops/fftconv.py:29-32); off in every HyenaDNA config
            # (hyena.py:260 passes gelu=False, dropout_mask=None), kept as plain device-side torch ops.
            pre_act = y
            o = y.float()
            if gelu:
                o = F.gelu(o)
            if dropout_mask is not None:
                o = o * dropout_mask.reshape(o.shape[0], o.shape[1], 1).to(o.dtype)
            out = o.to(cdt)
        ctx.save_for_backward(u3, Kf, v3, q3, ys, pre_act, dropout_mask, gs)
        out = out.reshape(shape)
        if force_fp16_output and in_dtype == torch.float32:
            out = out.to(torch.float16)
        elif out.dtype != in_dtype and not force_fp16_output:
            out = out.to(in_dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        u3, Kf, v3, q3, ys, pre_act, dropout_mask, gs = ctx.saved_tensors
        shape, kshape, Dshape, in_dtype = ctx.shapes
        L = ctx.L
        d = dout.reshape(u3.shape).to(u3.dtype)
        if pre_act is not None:
            with torch.enable_grad():
                p = pre_act.detach().float().requires_grad_(True)
                o = F.gelu(p) if ctx.gelu else p
                if dropout_mask is not None:
                    o = o * dropout_mask.reshape(o.shape[0], o.shape[1], 1).to(o.dtype)
                (d,) = torch.autograd.grad(o, p, d.float())
            d = d.to(u3.dtype)
        d = _rows(d)
        in_mode, out_mode = ctx.modes
        du, dv, dq, dKacc, dD = K.conv_bwd(d, u3, Kf, L, in_mode=in_mode, out_mode=out_mode, pre=v3, post=q3, ysave=ys,
                                           gsave=gs)
        dk = K.conv_dk(dKacc, L)
        if dD is None:
            dD = dk[:, 0]              # saved-spectrum backward: the skip weight is the lag-0 tap
        dk = dk.reshape(kshape)
        du = du.reshape(shape).to(in_dtype)
        dv = dv.reshape(shape).to(in_dtype) if dv is not None else None
        dq = dq.reshape(shape).to(in_dtype) if dq is not None else None
        return du, dk, dD.reshape(Dshape), None, None, None, None, dv, None, dq, None, None, None


def _post_ops(y, gelu, dropout_mask, out_dtype):
    """GELU / dropout-mask epilogue of the reference (src/ops/fftconv.py:29-34) for the composed variants."""
    if gelu or dropout_mask is not None:
        o = y.float()
        if gelu:
            o = F.gelu(o)
        if dropout_mask is not None:
            o = o * dropout_mask.reshape(o.shape[0], o.shape[1], *([1] * (o.dim() - 2))).to(o.dtype)
        y = o
    return y.to(out_dtype)


def _anticausal(g, k_rev):
    """sum_s k_rev[s] g[t + s] — what adding conj(rfft(k_rev)) to the filter spectrum does to a causal zero-padded
    convolution (src/ops/fftconv.py:19-21, hyena.py:64-66): the causal kernel on the time-reversed sequence."""
    zero = torch.zeros(k_rev.shape[:-1], dtype=torch.float32, device=k_rev.device)
    return FFTConvFunc.apply(g.flip(-1), k_rev, zero, None, False).flip(-1)


def circular_conv(g, kk, L_out):
    """y[t] = sum_s kk[s] g[(t - s) mod N] for t < L_out, N = kk.shape[-1] >= g.shape[-1]: a circular convolution of
    period N is the linear convolution folded once, y[t] = c[t] + c[t + N]; c comes from the causal kernels run at
    length N + L_out."""
    N = kk.shape[-1]
    P = g.shape[-1]
    T = N + L_out
    gp = F.pad(g, (0, T - P))
    kp = F.pad(kk.float(), (0, T - N))
    zero = torch.zeros(kp.shape[:-1], dtype=torch.float32, device=kp.device)
    c = FFTConvFunc.apply(gp, kp, zero, None, False)
    return c[..., :L_out].float() + c[..., N:N + L_out].float()


def _circular_2l(u, k, k_rev):
    """The fork's `bidirectional` long convolution (hyena.py:68-74): u is zero-padded by L/2 on both sides and convolved
    CIRCULARLY (period N = 2L) with k (+ the time-reversed k_rev)."""
    L = u.shape[-1]
    N = 2 * L
    padded_length = L + 2 * (L // 2)
    pad_before = padded_length // 2 - (L // 2)
    up = F.pad(u, (pad_before, padded_length - L - pad_before))                 # padded_u
    kk = F.pad(k.float(), (0, N - L))                                           # [..., N]
    if k_rev is not None:
        kr = F.pad(k_rev.float(), (0, N - L))
        kk = kk + torch.roll(kr.flip(-1), 1, dims=-1)                           # kr[(N - n) mod N]
    return circular_conv(up, kk, L)


def fftconv_func(u, k, D, dropout_mask=None, gelu=True, force_fp16_output=False, output_hbl_layout=False, v=None,
                 head_dim=1, q=None, fftfp16=False, k_rev=None, bidirectional=False):
    """Reference signature: src/ops/fftconv.py:105-108 (+ `bidirectional`, the fork's flag of hyena.py:60).

    The causal head_dim = 1 call — every HyenaDNA configuration — is ONE fused kernel family (FFTConvFunc).  The other
    keywords are compositions of that core with elementwise torch ops, differentiated by autograd:
      k_rev                adds the anticausal correlation with k_rev (src/ops/fftconv.py:19-21);
      bidirectional        period-2L circular convolution of the centred input (hyena.py:68-74);
      head_dim > 1         H3 multi-head outer product (src/ops/fftconv.py:38-55): the (d1, d2) pairs become channels of
                           the gated core, then the d1 axis is summed;
      output_hbl_layout    the result's memory is laid out [H, B, L] (src/ops/fftconv.py:91-94)."""
    in_dtype = u.dtype
    if head_dim == 1 and k_rev is None and not bidirectional:
        out = FFTConvFunc.apply(u, k, D, dropout_mask, gelu, force_fp16_output, False, v, 1, q, fftfp16, None, False)
    elif head_dim > 1:
        if v is None or q is None:
            raise ValueError("fftconv_func: head_dim > 1 needs v and q (H3 form)")
        K._check_dev(u)
        B, Hd, L = u.shape
        h = Hd // head_dim
        # (h, d1, d2) channels: k broadcast over d2, v over d1, q over d2; one filter / skip weight per head h
        ku = u.reshape(B, h, head_dim, 1, L).expand(B, h, head_dim, head_dim, L).reshape(B, -1, L)
        vu = v.reshape(B, h, 1, head_dim, L).expand(B, h, head_dim, head_dim, L).reshape(B, -1, L)
        qu = q.reshape(B, h, head_dim, 1, L).expand(B, h, head_dim, head_dim, L).reshape(B, -1, L)
        kf = k.reshape(h, 1, -1).expand(h, head_dim * head_dim, k.shape[-1]).reshape(-1, k.shape[-1])
        Df = D.reshape(h, 1).expand(h, head_dim * head_dim).reshape(-1)
        krf = None if k_rev is None else k_rev.reshape(h, 1, -1).expand(h, head_dim * head_dim, k_rev.shape[-1]).reshape(-1, k_rev.shape[-1])
        full = fftconv_func(ku, kf, Df, None, False, False, False, vu, 1, qu, fftfp16, krf, bidirectional)
        out = full.reshape(B, h, head_dim, head_dim, L).float().sum(dim=2).reshape(B, Hd, L)      # sum over d1 -> (h d2)
        out = _post_ops(out, gelu, dropout_mask, in_dtype)
    else:
        K._check_dev(u)
        g = u if v is None else u * v
        Dv = D.unsqueeze(-1)                                             # src/ops/fftconv.py:28
        if bidirectional:
            y = _circular_2l(g, k, k_rev)
        else:
            y = FFTConvFunc.apply(g, k, torch.zeros_like(D, dtype=torch.float32), None, False).float() + _anticausal(g, k_rev).float()
        y = y + g.float() * Dv.float()
        if q is not None:
            y = y * q.float()
        out = _post_ops(y, gelu, dropout_mask, in_dtype)
    if force_fp16_output and in_dtype == torch.float32 and out.dtype == torch.float32:
        out = out.to(torch.float16)
    if output_hbl_layout and out.dim() == 3:
        out = out.transpose(0, 1).contiguous().transpose(0, 1)        # [B, H, L] view of [H, B, L] memory
    return out
