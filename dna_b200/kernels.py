"""Thin tensor -> pointer layer over the C-ABI (include/hyena_b200.h).

Every function here allocates its outputs/workspaces with torch (the caller's caching allocator owns
all memory, as at the reference's extension boundary, src/ops/fftconv.py:58-103), passes raw device
pointers + sizes + the current CUDA stream to libhyena_b200.so, and raises on any error. No function
in this module computes anything in Python/torch: if the CUDA library is unavailable they raise.
"""
from __future__ import annotations

import ctypes as C
import functools
import itertools
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (HY_BF16, HY_F32, IN_PLAIN, IN_PREGATE, IN_SHORTCONV, OUT_PLAIN, OUT_POSTGATE, OUT_SHORTCONV,
                   ConvBwdArgs, ConvFwdArgs, FilterArgs)


def _device_guarded(fn):
    """Run a C-ABI wrapper on the device of its tensor arguments: the library, its per-device tables and the stream it
    launches on are all looked up through the CUDA *current* device, so a model living on cuda:1 while the current
    device is cuda:0 must switch for the duration of the call (native torch ops do this implicitly).  All tensor
    arguments have to share one device."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if _lib.is_emulation():
            return fn(*args, **kwargs)
        dev = None
        for a in itertools.chain(args, kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise _lib.HyenaB200Error(f"hyena-b200 kernels: tensor arguments on different devices ({dev} and {a.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


# ---- optional device-side timing of the long-conv entry points (bench.py roofline leg) -----------
_TIMING = None   # list of (tag, start_event, end_event) when enabled


def enable_timing(on: bool = True):
    """Record a CUDA event pair around every long-conv C-ABI call on the current stream."""
    global _TIMING
    _TIMING = [] if on else None


def drain_timing():
    """Synchronise and return {tag: (calls, total_ms)} for the calls recorded since the last drain."""
    global _TIMING
    out = {}
    if _TIMING is None:
        return out
    torch.cuda.synchronize()
    for tag, e0, e1 in _TIMING:
        c, t = out.get(tag, (0, 0.0))
        out[tag] = (c + 1, t + e0.elapsed_time(e1))
    _TIMING = []
    return out


class _timed:
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if _TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _TIMING is not None:
            self.e1.record()
            _TIMING.append((self.tag, self.e0, self.e1))
        return False


@_device_guarded
def clock_probe(out: torch.Tensor):
    """Enqueue the SM-clock probe; `out` is an int64 [2] device tensor receiving (cycles, nanoseconds)."""
    lib = _lib.lib()
    assert out.dtype == torch.int64 and out.numel() >= 2 and out.is_contiguous()
    _lib.check(lib.hy_clock_probe(_p(out), _lib.current_stream_ptr()))


def launch_count() -> int:
    return int(_lib.load_library().hy_launch_count())


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return HY_F32
    if t.dtype == torch.bfloat16:
        return HY_BF16
    raise TypeError(f"hyena-b200 kernels take float32 or bfloat16 activations, got {t.dtype}")


def _check_dev(*ts):
    emu = _lib.is_emulation()
    for t in ts:
        if t is None:
            continue
        if not emu and not t.is_cuda:
            raise _lib.HyenaB200Error("hyena-b200 kernels need CUDA tensors (no CPU fallback)")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _rows3(t: torch.Tensor) -> Tuple[int, int]:
    """(batch stride, row stride) of a [B, H, L] tensor whose last dim is contiguous."""
    assert t.dim() == 3 and t.stride(2) == 1, "expected [B, H, L] with unit stride along L"
    return t.stride(0), t.stride(1)


def fft_len(L: int) -> int:
    m = _lib.load_library().hy_fft_len(int(L))
    if m < 0:
        raise _lib.HyenaB200Error(f"unsupported sequence length {L}")
    return m


def _workspace(B, H, L, nseq, device):
    n = _lib.load_library().hy_conv_workspace_bytes(B, H, L, nseq)
    if n == 0:
        return None, 0
    ws = torch.empty(n, dtype=torch.uint8, device=device)
    return ws, n


@_device_guarded
def filter_spectrum(k: torch.Tensor, D: Optional[torch.Tensor], L: int) -> torch.Tensor:
    """Kf [H, M] complex64 (internal order) of (k + D*delta)/M. k: fp32 [H, >=L], D: fp32 [H] or None."""
    lib = _lib.lib()
    _check_dev(k, D)
    assert k.dtype == torch.float32 and k.dim() == 2 and k.stride(1) == 1 and k.shape[1] >= L
    H = k.shape[0]
    M = fft_len(L)
    Kf = torch.empty((H, M, 2), dtype=torch.float32, device=k.device)
    ws, n = _workspace(1, H, L, 1, k.device)
    if D is not None:
        D = D.detach().to(torch.float32).contiguous()
        assert D.numel() == H
    with _timed("spectrum"):
        _lib.check(lib.hy_filter_spectrum(_p(k), k.stride(0), _p(D), _p(Kf), H, L, _p(ws), n, _lib.current_stream_ptr()))
    return Kf


def conv_gsave_alloc(B: int, H: int, L: int, device, max_bytes: Optional[int] = None):
    """Buffer for the saved spectrum of g (conv_fwd writes it, conv_bwd reads it and skips one of its two forward
    transforms), or None when this length has no use for it (single-kernel regime) or it would exceed `max_bytes`
    (default: HYENA_B200_GSAVE_MAX_MB, 4096 MB per call)."""
    n = int(_lib.lib().hy_conv_gsave_bytes(B, H, L))
    if max_bytes is None:
        max_bytes = int(os.environ.get("HYENA_B200_GSAVE_MAX_MB", "4096")) << 20
    if n == 0 or n > max_bytes:
        return None
    return torch.empty(n // 4, dtype=torch.float32, device=device)


@_device_guarded
def conv_fwd(u, Kf, L, *, in_mode=IN_PLAIN, out_mode=OUT_PLAIN, pre=None, post=None, sw=None, sb=None, pb=None,
             H=None, save_y=False, gsave=None):
    """Fused long conv forward. Returns (out [B,H,ldo->L view], ysave or None). `gsave` (from conv_gsave_alloc) is
    filled with the spectrum of g for conv_bwd."""
    lib = _lib.lib()
    _check_dev(u, Kf, pre, post, sw, sb, pb)
    B = u.shape[0]
    if H is None:
        H = u.shape[1] // 3 if in_mode == IN_SHORTCONV else u.shape[1]
    u_bs, ldu = _rows3(u)
    ldo = (L + 7) // 8 * 8
    out_full = torch.empty((B, H, ldo), dtype=u.dtype, device=u.device)
    ys_full = torch.empty((B, H, ldo), dtype=u.dtype, device=u.device) if save_y else None
    a = ConvFwdArgs()
    a.dtype = _dtype_code(u)
    a.B, a.H, a.L = B, H, L
    a.in_mode, a.out_mode = in_mode, out_mode
    a.u, a.u_bs, a.ldu = u.data_ptr(), u_bs, ldu
    if pre is not None:
        assert pre.dtype == u.dtype and _rows3(pre) == (u_bs, ldu)
        a.pre = pre.data_ptr()
    if post is not None:
        assert post.dtype == u.dtype
        a.post = post.data_ptr()
        a.post_bs, a.ldpost = _rows3(post)
    for name, t in (("sw", sw), ("sb", sb), ("pb", pb)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous()
            setattr(a, name, t.data_ptr())
    a.Kf = Kf.data_ptr()
    a.out = out_full.data_ptr()
    a.ysave = ys_full.data_ptr() if save_y else None
    a.out_bs, a.ldo = H * ldo, ldo
    ws, n = _workspace(B, H, L, 1, u.device)
    a.ws, a.ws_bytes = (ws.data_ptr() if ws is not None else None), n
    if gsave is not None:
        assert gsave.is_contiguous() and gsave.numel() * gsave.element_size() == lib.hy_conv_gsave_bytes(B, H, L)
        a.gsave = gsave.data_ptr()
    with _timed("conv_fwd"):
        _lib.check(lib.hy_conv_fwd(C.byref(a), _lib.current_stream_ptr()))
    return out_full[:, :, :L], (ys_full[:, :, :L] if save_y else None)


@_device_guarded
def conv_bwd(dout, u, Kf, L, *, in_mode=IN_PLAIN, out_mode=OUT_PLAIN, pre=None, post=None, sw=None, sb=None, pb=None,
             ysave=None, H=None, nslot=None, gsave=None, defer_dx0=False):
    """Fused long conv backward.

    Returns (du, dpre, dpost, dKacc, dD) where du has the layout of u (for SHORTCONV it is
    dX = (dx0|dx1|dv) in uT layout), dKacc is the [nslot, H, M] spectrum product consumed by
    conv_dk and dD [H] fp32 -- or None when `gsave` (the buffer conv_fwd filled) is given: dD is then
    conv_dk(dKacc)[:, 0].  defer_dx0 (SHORTCONV): the x0 group of dX is left unwritten; pass dout and ysave to
    shortconv_bwd, which forms dx0 = dout * y while it streams those rows.
    """
    lib = _lib.lib()
    _check_dev(dout, u, Kf, pre, post, sw, sb, pb, ysave)
    B = u.shape[0]
    if H is None:
        H = u.shape[1] // 3 if in_mode == IN_SHORTCONV else u.shape[1]
    M = fft_len(L)
    u_bs, ldu = _rows3(u)
    assert dout.dtype == u.dtype
    out_bs, ldo = _rows3(dout)
    if nslot is None:
        nslot = min(B, 8)
    a = ConvBwdArgs()
    a.dtype = _dtype_code(u)
    a.B, a.H, a.L = B, H, L
    a.in_mode, a.out_mode = in_mode, out_mode
    a.u, a.u_bs, a.ldu = u.data_ptr(), u_bs, ldu
    du = torch.empty_strided(u.shape, u.stride(), dtype=u.dtype, device=u.device)
    a.du = du.data_ptr()
    dpre = dpost = None
    if pre is not None:
        assert pre.dtype == u.dtype and _rows3(pre) == (u_bs, ldu)
        a.pre = pre.data_ptr()
        dpre = torch.empty_strided(u.shape, u.stride(), dtype=u.dtype, device=u.device)
        a.dpre = dpre.data_ptr()
    if post is not None:
        a.post = post.data_ptr()
        a.post_bs, a.ldpost = _rows3(post)
        dpost = torch.empty_strided(post.shape, post.stride(), dtype=post.dtype, device=post.device)
        a.dpost = dpost.data_ptr()
    for name, t in (("sw", sw), ("sb", sb), ("pb", pb)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous()
            setattr(a, name, t.data_ptr())
    a.Kf = Kf.data_ptr()
    a.dout = dout.data_ptr()
    if ysave is not None:
        assert ysave.dtype == u.dtype
        a.ysave = ysave.data_ptr()
        a.ys_bs, a.ldys = _rows3(ysave)
    a.out_bs, a.ldo = out_bs, ldo
    dKacc = torch.empty((nslot, H, M, 2), dtype=torch.float32, device=u.device)
    a.dKacc, a.nslot = dKacc.data_ptr(), nslot
    a.defer_dx0 = int(bool(defer_dx0))
    dDpart = None
    if gsave is not None:
        assert gsave.is_contiguous() and gsave.numel() * gsave.element_size() == lib.hy_conv_gsave_bytes(B, H, L)
        a.gsave = gsave.data_ptr()
    else:
        ndpart = lib.hy_conv_ndpart(L)
        dDpart = torch.zeros((B, H, ndpart), dtype=torch.float32, device=u.device)
        a.dDpart = dDpart.data_ptr()
    ws, n = _workspace(B, H, L, 1 if gsave is not None else 2, u.device)
    a.ws, a.ws_bytes = (ws.data_ptr() if ws is not None else None), n
    with _timed("conv_bwd"):
        _lib.check(lib.hy_conv_bwd(C.byref(a), _lib.current_stream_ptr()))
    dD = dDpart.sum(dim=(0, 2)) if dDpart is not None else None
    return du, dpre, dpost, dKacc, dD


@_device_guarded
def conv_dk(dKacc: torch.Tensor, L: int) -> torch.Tensor:
    """dk [H, L] fp32 from the spectrum products of conv_bwd."""
    lib = _lib.lib()
    nslot, H, M, _ = dKacc.shape
    ld = (L + 7) // 8 * 8
    dk = torch.empty((H, ld), dtype=torch.float32, device=dKacc.device)
    ws, n = _workspace(1, H, L, 1, dKacc.device)
    with _timed("conv_dk"):
        _lib.check(lib.hy_conv_dk(_p(dKacc), nslot, _p(dk), ld, H, L, _p(ws), n, _lib.current_stream_ptr()))
    return dk[:, :L]


@_device_guarded
def shortconv_fwd(uT, sw, sb, pb, L):
    lib = _lib.lib()
    _check_dev(uT, sw, sb, pb)
    B, H3, _ = uT.shape
    bs, ld = _rows3(uT)
    xc = torch.empty_strided(uT.shape, uT.stride(), dtype=uT.dtype, device=uT.device)
    _lib.check(lib.hy_shortconv_fwd(_dtype_code(uT), _p(uT), _p(xc), bs, ld, _p(sw), _p(sb), _p(pb), B, H3, L,
                                    _lib.current_stream_ptr()))
    return xc[:, :, :L]


def _rows16(t):
    bs, ld = _rows3(t)
    return t.data_ptr() % 16 == 0 and bs % 8 == 0 and ld % 8 == 0


def shortconv_gate_supported(uT, dout, ysave) -> bool:
    """Can shortconv_bwd form dx0 = dout * ysave itself (conv_bwd(defer_dx0=True))? Needs 16-byte aligned rows with
    strides that are multiples of 8 elements."""
    return all(t is not None and t.stride(-1) == 1 and t.dim() == 3 and _rows16(t) for t in (uT, dout, ysave))


@_device_guarded
def shortconv_bwd(uT, dX, sw, pb, L, dout=None, ysave=None):
    """Returns (duT, dsw [3H,3], dsb [3H], dpb [3H]). dout/ysave: the x0 group of dX is taken as dout * ysave
    (partner of conv_bwd(defer_dx0=True))."""
    lib = _lib.lib()
    _check_dev(uT, dX, sw, pb, dout, ysave)
    B, H3, _ = uT.shape
    bs, ld = _rows3(uT)
    assert _rows3(dX) == (bs, ld) and dX.dtype == uT.dtype
    duT = torch.empty_strided(uT.shape, uT.stride(), dtype=uT.dtype, device=uT.device)
    nchunk = lib.hy_shortconv_nchunk(B, L)
    dwpart = torch.empty((nchunk, H3, 4), dtype=torch.float32, device=uT.device)
    dpbpart = torch.empty((nchunk, H3), dtype=torch.float32, device=uT.device)
    if dout is not None:
        assert ysave is not None and dout.dtype == uT.dtype and ysave.dtype == uT.dtype
        z_bs, z_ld = _rows3(dout)
        y_bs, y_ld = _rows3(ysave)
        with _timed("shortconv_bwd"):
            _lib.check(lib.hy_shortconv_bwd_gate(_dtype_code(uT), _p(uT), _p(dX), _p(duT), bs, ld, _p(sw), _p(pb), _p(dwpart),
                                                 _p(dpbpart), B, H3, L, _p(dout), z_bs, z_ld, _p(ysave), y_bs, y_ld,
                                                 _lib.current_stream_ptr()))
        dw = dwpart.sum(0)
        return duT, dw[:, :3].contiguous(), dw[:, 3].contiguous(), dpbpart.sum(0)
    with _timed("shortconv_bwd"):
        _lib.check(lib.hy_shortconv_bwd(_dtype_code(uT), _p(uT), _p(dX), _p(duT), bs, ld, _p(sw), _p(pb), _p(dwpart),
                                        _p(dpbpart), B, H3, L, _lib.current_stream_ptr()))
    dw = dwpart.sum(0)
    return duT, dw[:, :3].contiguous(), dw[:, 3].contiguous(), dpbpart.sum(0)


@_device_guarded
def filter_fwd(z, t, w_in, b_in, w_h, b_h, w_out, freq, deltas, shift, modulate, L, save_h=False, save_trunk=False):
    """k [D, L] fp32 (channel-major, padded row stride). save_h: also return the last hidden activation h_last
    [L, order] (needs filter_trunk_bwd_supported(order, emb, n_inner)) -> (k, h_last). save_trunk (with save_h): also
    the trunk's pre-activations a_save [1 + n_inner, 64, lda] for filter_trunk_bwd(a_save=...) -> (k, h_last, a_save)."""
    lib = _lib.lib()
    _check_dev(z, t, w_in, b_in, w_h, b_h, w_out, freq, deltas)
    D, order = w_out.shape
    emb = w_in.shape[1]
    n_inner = 0 if w_h is None else w_h.shape[0]
    a = FilterArgs()
    a.L, a.D, a.order, a.emb_dim, a.n_inner = L, D, order, emb, n_inner
    z2 = z.reshape(-1, z.shape[-1])
    assert z2.stride(1) == 1 and z2.shape[0] >= L and z2.dtype == torch.float32
    t1 = t.reshape(-1)
    assert t1.is_contiguous() and t1.shape[0] >= L and t1.dtype == torch.float32
    a.z, a.ldz, a.t = z2.data_ptr(), z2.stride(0), t1.data_ptr()
    keep = [z2, t1]
    for name, ten in (("w_in", w_in), ("b_in", b_in), ("w_h", w_h), ("b_h", b_h), ("w_out", w_out), ("freq", freq),
                      ("deltas", deltas)):
        if ten is not None:
            ten = ten.detach().to(torch.float32).contiguous()
            keep.append(ten)
            setattr(a, name, ten.data_ptr())
    a.shift, a.modulate = float(shift), int(bool(modulate))
    ld = (L + 7) // 8 * 8
    k = torch.empty((D, ld), dtype=torch.float32, device=w_out.device)
    if save_h and save_trunk:
        h_last = torch.empty((L, order), dtype=torch.float32, device=w_out.device)
        lda, elems = C.c_int(0), C.c_longlong(0)
        _lib.check(lib.hy_filter_trunk_save_layout(C.byref(a), C.byref(lda), C.byref(elems)))
        a_save = torch.empty((1 + n_inner, elems.value // ((1 + n_inner) * lda.value), lda.value), dtype=torch.float32,
                             device=w_out.device)
        with _timed("filter_fwd"):
            _lib.check(lib.hy_filter_fwd_save_trunk(C.byref(a), _p(k), ld, _p(h_last), order, _p(a_save), lda.value,
                                                    _lib.current_stream_ptr()))
        return k[:, :L], h_last, a_save
    if save_h:
        h_last = torch.empty((L, order), dtype=torch.float32, device=w_out.device)
        with _timed("filter_fwd"):
            _lib.check(lib.hy_filter_fwd_save(C.byref(a), _p(k), ld, _p(h_last), order, _lib.current_stream_ptr()))
        return k[:, :L], h_last
    with _timed("filter_fwd"):
        _lib.check(lib.hy_filter_fwd(C.byref(a), _p(k), ld, _lib.current_stream_ptr()))
    return k[:, :L]


@_device_guarded
def filter_modulate_bwd(dk, t, deltas, shift, modulate, L):
    """dh [L, D] fp32 = dk[D, L]^T * (exp(-t|deltas|) + shift): gradient wrt the MLP's last Linear output."""
    lib = _lib.lib()
    _check_dev(dk, t, deltas)
    assert dk.dtype == torch.float32 and dk.dim() == 2 and dk.stride(1) == 1
    D = dk.shape[0]
    t1 = t.detach().reshape(-1)
    assert t1.is_contiguous() and t1.dtype == torch.float32 and t1.numel() >= L
    dl = deltas.detach().to(torch.float32).reshape(-1).contiguous() if deltas is not None else None
    dh = torch.empty((L, D), dtype=torch.float32, device=dk.device)
    with _timed("filter_bwd"):
        _lib.check(lib.hy_filter_modulate_bwd(_p(dk), dk.stride(0), _p(t1), _p(dl), float(shift), int(bool(modulate)),
                                              _p(dh), D, L, D, _lib.current_stream_ptr()))
    return dh


def filter_out_bwd_supported(D, order) -> bool:
    return bool(_lib.lib().hy_filter_out_bwd_supported(int(D), int(order)))


@_device_guarded
def filter_out_bwd(dk, t, deltas, shift, modulate, w_out, h_last, L):
    """Backward of the filter MLP's last Linear fused with the modulation backward (tensor cores, 3xTF32).
    dk [D, L] fp32 channel-major, h_last [L, order] fp32 -> (dh_last [L, order], dW_out [D, order])."""
    lib = _lib.lib()
    _check_dev(dk, t, deltas, w_out, h_last)
    assert dk.dtype == torch.float32 and dk.dim() == 2 and dk.stride(1) == 1
    D, order = w_out.shape
    if dk.stride(0) % 4 or dk.data_ptr() % 16:
        ld = (L + 3) // 4 * 4
        buf = torch.empty((D, ld), dtype=torch.float32, device=dk.device)
        buf[:, :L].copy_(dk[:, :L])
        dk = buf
    t1 = t.detach().reshape(-1)
    assert t1.is_contiguous() and t1.dtype == torch.float32 and t1.numel() >= L
    dl = deltas.detach().to(torch.float32).reshape(-1).contiguous() if deltas is not None else None
    w = w_out.detach().to(torch.float32).contiguous()
    assert h_last.dtype == torch.float32 and h_last.dim() == 2 and h_last.stride(1) == 1 and h_last.shape[0] >= L
    dh_last = torch.empty((L, order), dtype=torch.float32, device=dk.device)
    dW = torch.empty((D, order), dtype=torch.float32, device=dk.device)
    nws = int(lib.hy_filter_out_bwd_workspace_bytes(L))
    ws = torch.empty((nws,), dtype=torch.uint8, device=dk.device)
    with _timed("filter_bwd"):
        _lib.check(lib.hy_filter_out_bwd(_p(dk), dk.stride(0), _p(t1), _p(dl), float(shift), int(bool(modulate)), _p(w),
                                         _p(h_last), h_last.stride(0), _p(dh_last), order, _p(dW), D, order, L,
                                         _p(ws), nws, _lib.current_stream_ptr()))
    return dh_last, dW


def _filter_args(z, t, w_in, b_in, w_h, b_h, w_out, freq, deltas, shift, modulate, L, keep):
    D, order = w_out.shape
    emb = w_in.shape[1]
    n_inner = 0 if w_h is None else w_h.shape[0]
    a = FilterArgs()
    a.L, a.D, a.order, a.emb_dim, a.n_inner = L, D, order, emb, n_inner
    z2 = z.reshape(-1, z.shape[-1])
    assert z2.stride(1) == 1 and z2.shape[0] >= L and z2.dtype == torch.float32
    t1 = t.reshape(-1)
    assert t1.is_contiguous() and t1.shape[0] >= L and t1.dtype == torch.float32
    a.z, a.ldz, a.t = z2.data_ptr(), z2.stride(0), t1.data_ptr()
    keep += [z2, t1]
    for name, ten in (("w_in", w_in), ("b_in", b_in), ("w_h", w_h), ("b_h", b_h), ("w_out", w_out), ("freq", freq),
                      ("deltas", deltas)):
        if ten is not None:
            ten = ten.detach().to(torch.float32).contiguous()
            keep.append(ten)
            setattr(a, name, ten.data_ptr())
    a.shift, a.modulate = float(shift), int(bool(modulate))
    return a


def filter_trunk_bwd_supported(order, emb, n_inner):
    return order <= 64 and emb <= 8 and n_inner <= 2


def filter_trunk_save_bytes(order, emb, n_inner, L) -> int:
    """Size of the pre-activation buffer filter_fwd(save_trunk=True) keeps for the backward."""
    return 4 * (1 + n_inner) * 64 * ((L + 63) // 64 * 64)


@_device_guarded
def filter_trunk_bwd(dh_last, z, t, w_in, b_in, w_h, b_h, w_out, freq, L, a_save=None):
    """Fused backward of the MLP trunk. dh_last: [L, order] fp32. a_save: the pre-activations kept by
    filter_fwd(save_trunk=True) (no recompute of the trunk). Returns (dW_in, db_in, dW_h, db_h, dfreq)."""
    lib = _lib.lib()
    _check_dev(dh_last, z, t, w_in, b_in, w_h, b_h, freq)
    keep = []
    a = _filter_args(z, t, w_in, b_in, w_h, b_h, w_out, freq, None, 0.0, False, L, keep)
    O, E, NI = a.order, a.emb_dim, a.n_inner
    n_cta, stride = C.c_int(0), C.c_int(0)
    _lib.check(lib.hy_filter_trunk_bwd_layout(C.byref(a), C.byref(n_cta), C.byref(stride)))
    assert dh_last.dtype == torch.float32 and dh_last.dim() == 2 and dh_last.stride(1) == 1 and dh_last.shape[0] >= L
    part = torch.zeros((n_cta.value, stride.value), dtype=torch.float32, device=dh_last.device)
    with _timed("filter_bwd"):
        if a_save is not None:
            _check_dev(a_save)
            assert a_save.dtype == torch.float32 and a_save.is_contiguous() and a_save.dim() == 3
            _lib.check(lib.hy_filter_trunk_bwd_saved(C.byref(a), _p(dh_last), dh_last.stride(0), _p(a_save), a_save.shape[2],
                                                     _p(part), _lib.current_stream_ptr()))
        else:
            _lib.check(lib.hy_filter_trunk_bwd(C.byref(a), _p(dh_last), dh_last.stride(0), _p(part),
                                               _lib.current_stream_ptr()))
    tot = part.sum(0)
    off = 0
    dW_in = tot[off:off + O * E].reshape(O, E); off += O * E
    db_in = tot[off:off + O]; off += O
    dW_h, db_h = [], []
    for _ in range(NI):
        dW_h.append(tot[off:off + O * O].reshape(O, O)); off += O * O
        db_h.append(tot[off:off + O]); off += O
    dfreq = tot[off:off + O]
    return dW_in, db_in, dW_h, db_h, dfreq


@_device_guarded
def tokenize(seqs: torch.Tensor, lens: Optional[torch.Tensor], max_length: int, flags: int) -> torch.Tensor:
    """seqs: uint8 [B, max_chars]; lens: int32 [B] or None -> ids int64 [B, max_length]."""
    lib = _lib.lib()
    _check_dev(seqs, lens)
    assert seqs.dtype == torch.uint8 and seqs.dim() == 2 and seqs.stride(1) == 1
    B, max_chars = seqs.shape
    if lens is not None:
        assert lens.dtype == torch.int32 and lens.is_contiguous() and lens.numel() == B
    ids = torch.empty((B, max_length), dtype=torch.int64, device=seqs.device)
    _lib.check(lib.hy_tokenize(_p(seqs), seqs.stride(0), _p(lens), max_chars, _p(ids), B, max_length, flags,
                               _lib.current_stream_ptr()))
    return ids


@_device_guarded
def reverse_complement(seqs: torch.Tensor, lens: Optional[torch.Tensor] = None,
                       apply: Optional[torch.Tensor] = None) -> torch.Tensor:
    """seqs: uint8 [B, max_chars]; lens int32 [B] or None; apply: uint8/bool [B] or None (all rows) -> uint8 [B, max_chars]."""
    lib = _lib.lib()
    _check_dev(seqs, lens, apply)
    assert seqs.dtype == torch.uint8 and seqs.dim() == 2 and seqs.stride(1) == 1
    B, max_chars = seqs.shape
    if lens is not None:
        assert lens.dtype == torch.int32 and lens.is_contiguous() and lens.numel() == B
    if apply is not None:
        apply = apply.to(torch.uint8).contiguous()
        assert apply.numel() == B
    out = torch.empty((B, max_chars), dtype=torch.uint8, device=seqs.device)
    if B > 0 and max_chars > 0:
        _lib.check(lib.hy_reverse_complement(_p(seqs), seqs.stride(0), _p(lens), _p(apply), _p(out), out.stride(0), B,
                                             max_chars, _lib.current_stream_ptr()))
    return out


@_device_guarded
def fetch_intervals(chrom: torch.Tensor, starts: torch.Tensor, ends: torch.Tensor, max_length: int, *, rc=None,
                    pad_interval: bool = False, width: Optional[int] = None):
    """FastaInterval.__call__ (hg38_dataset.py:72-124) for a batch of BED intervals of ONE chromosome held in device
    memory: chrom uint8 [chrom_len]; starts / ends int64 [B]; rc: bool/uint8 [B] or None.
    -> (bytes uint8 [B, width], lens int32 [B]); width defaults to max_length."""
    lib = _lib.lib()
    _check_dev(chrom)
    assert chrom.dtype == torch.uint8 and chrom.dim() == 1 and chrom.is_contiguous()
    # the interval table is a few integers per row: host tensors / lists are accepted and copied to the chromosome's device
    starts = torch.as_tensor(starts).to(chrom.device, torch.int64).contiguous()
    ends = torch.as_tensor(ends).to(chrom.device, torch.int64).contiguous()
    B = starts.numel()
    assert ends.numel() == B
    if rc is not None:
        rc = torch.as_tensor(rc).to(chrom.device, torch.uint8).contiguous()
        assert rc.numel() == B
    width = int(width or max_length)
    ld = (width + 15) // 16 * 16
    out = torch.empty((B, ld), dtype=torch.uint8, device=chrom.device)
    lens = torch.empty((B,), dtype=torch.int32, device=chrom.device)
    if B > 0:
        _lib.check(lib.hy_fetch_intervals(_p(chrom), chrom.numel(), _p(starts), _p(ends), _p(rc), B, int(max_length),
                                          int(bool(pad_interval)), _p(out), out.stride(0), _p(lens), width,
                                          _lib.current_stream_ptr()))
    return out[:, :width], lens


@_device_guarded
def bert_mask(seq: torch.Tensor, r_mask: torch.Tensor, r_kind: torch.Tensor, rand_tok: torch.Tensor, mask_token_id: int,
              pad_token_id: int, mask_prob: float = 0.15, random_token_prob: float = 0.1, unchanged_token_prob: float = 0.1):
    """bert_mask (hg38_dataset.py:238-286) for given random draws -> (masked seq int64, mask bool, labels int64)."""
    lib = _lib.lib()
    _check_dev(seq, r_mask, r_kind, rand_tok)
    assert seq.dtype == torch.int64 and rand_tok.dtype == torch.int64 and r_mask.dtype == torch.float32 and r_kind.dtype == torch.float32
    seq, r_mask, r_kind, rand_tok = (t.contiguous() for t in (seq, r_mask, r_kind, rand_tok))
    n = seq.numel()
    assert r_mask.numel() == n and r_kind.numel() == n and rand_tok.numel() == n
    out = torch.empty_like(seq)
    mask = torch.empty(seq.shape, dtype=torch.uint8, device=seq.device)
    labels = torch.empty_like(seq)
    _lib.check(lib.hy_bert_mask(_p(seq), _p(r_mask), _p(r_kind), _p(rand_tok), n, int(mask_token_id), int(pad_token_id),
                                float(mask_prob), float(random_token_prob), float(unchanged_token_prob), _p(out), _p(mask),
                                _p(labels), _lib.current_stream_ptr()))
    return out, mask.bool(), labels


# ---- Block glue: residual add + LayerNorm ------------------------------------------------------------
def add_ln_supported(D: int) -> bool:
    return bool(_lib.load_library().hy_add_ln_supported(int(D)))


def _torch_dtype(code: int):
    return torch.bfloat16 if code == HY_BF16 else torch.float32


@_device_guarded
def add_ln_fwd(x, res_in, gamma, beta, eps, y_dtype, res_dtype, write_res, keep=None, keep_scale=1.0):
    """x, res_in: [..., D] (either may be None). Returns (y, res_out or None, mean, rstd).
    keep: uint8 / bool mask of x's shape (dropout applied to x before the add), keep_scale = 1 / (1 - p)."""
    lib = _lib.lib()
    _check_dev(x, res_in, gamma, beta, keep)
    if keep is not None:
        assert x is not None and keep.shape == x.shape and keep.is_contiguous() and keep.element_size() == 1
    ref = x if x is not None else res_in
    D = ref.shape[-1]
    rows = ref.numel() // D
    for t in (x, res_in):
        assert t is None or (t.is_contiguous() and t.shape == ref.shape)
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32 and gamma.is_contiguous() and beta.is_contiguous()
    if res_in is not None:
        assert res_in.dtype == res_dtype
    y = torch.empty(ref.shape, dtype=y_dtype, device=ref.device)
    res_out = torch.empty(ref.shape, dtype=res_dtype, device=ref.device) if write_res else None
    mean = torch.empty(rows, dtype=torch.float32, device=ref.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=ref.device)
    xdt = _dtype_code(x) if x is not None else HY_F32
    with _timed("add_ln_fwd"):
        _lib.check(lib.hy_add_ln_dropout_fwd(_p(x), xdt, _p(keep), float(keep_scale), _p(res_in),
                                             _dtype_code(torch.empty(0, dtype=res_dtype)), _p(gamma), _p(beta), float(eps),
                                             _p(y), _dtype_code(y), _p(res_out), _p(mean), _p(rstd), rows, D,
                                             _lib.current_stream_ptr()))
    return y, res_out, mean, rstd


@_device_guarded
def add_ln_bwd(dy, dres_out, r, mean, rstd, gamma, x_dtype, want_dx, want_dres, keep=None, keep_scale=1.0):
    """Returns (dx or None, dres_in or None, dgamma, dbeta). keep / keep_scale: the forward's dropout mask of x."""
    lib = _lib.lib()
    _check_dev(dy, dres_out, r, mean, rstd, gamma, keep)
    D = r.shape[-1]
    rows = r.numel() // D
    assert dy.is_contiguous() and r.is_contiguous() and dy.shape == r.shape
    assert dres_out is None or (dres_out.is_contiguous() and dres_out.dtype == r.dtype and dres_out.shape == r.shape)
    dx = torch.empty(r.shape, dtype=x_dtype, device=r.device) if want_dx else None
    dres = torch.empty(r.shape, dtype=r.dtype, device=r.device) if want_dres else None
    nparts = int(lib.hy_add_ln_bwd_parts(rows, D))
    part = torch.empty((nparts, 2, D), dtype=torch.float32, device=r.device)
    dgamma = torch.empty(D, dtype=torch.float32, device=r.device)
    dbeta = torch.empty(D, dtype=torch.float32, device=r.device)
    xdt = HY_BF16 if x_dtype == torch.bfloat16 else HY_F32
    with _timed("add_ln_bwd"):
        _lib.check(lib.hy_add_ln_dropout_bwd(_p(dy), _dtype_code(dy), _p(dres_out), _dtype_code(r), _p(r), _p(mean), _p(rstd),
                                             _p(gamma), _p(dx), xdt, _p(keep), float(keep_scale), _p(dres), _p(part),
                                             _p(dgamma), _p(dbeta), rows, D, _lib.current_stream_ptr()))
    return dx, dres, dgamma, dbeta
