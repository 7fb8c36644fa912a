"""Parity checks shared by the CPU-emulation suite (tests/test_emu_*.py, device='cpu') and the
GPU suite (tests/test_gpu_*.py, device='cuda').  Every check compares the kernels — reached through the
C-ABI via dna_b200.kernels / dna_b200.hyena — against oracle/hyena_oracle.py on the same seeded inputs.

Tolerances (stated once, used everywhere):
  fp32 activations : max-abs error <= 2e-5 * max|oracle|   (the oracle's own fp32-vs-fp64 error is 2-3e-7
                     relative per SURVEY §8c; ours measures 2-7e-7 — the bound leaves headroom for sums
                     over 1e6 terms such as dbias)
  bf16 activations : err(ours, fp64 truth) <= 2 * err(oracle_bf16, fp64 truth) + 2^-8 * max|truth|
                     i.e. we may not be further from the exact answer than the reference's own bf16 path
                     (whose intermediates round to bf16) plus one output ulp.
  tokenizer        : bit-exact.
"""
from __future__ import annotations

import numpy as np
import torch

from dna_b200 import kernels as K
from dna_b200._lib import IN_PLAIN, IN_PREGATE, IN_SHORTCONV, OUT_PLAIN, OUT_POSTGATE, OUT_SHORTCONV
from oracle import hyena_oracle as O

FP32_TOL = 2e-5


def relerr(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def decaying_filter(H, L, gen):
    return torch.randn(H, L, generator=gen) * torch.exp(-torch.arange(L) / (L / 4.0))[None]


def check_fp32(name, got, ref, tol=FP32_TOL):
    e = relerr(got, ref)
    assert e <= tol, f"{name}: rel err {e:.3e} > {tol:.1e}"
    return e


def conv_case(B, H, L, mode, device, seed=0, dtype=torch.float32, gsave=False, defer=False, nslot=None):
    """One fused long-conv forward+backward case; returns dict of relative errors vs the oracle.
    gsave: the forward keeps the spectrum of g and the backward reads it back (four-step lengths only; dD is then
    dk[:, 0])."""
    gs = K.conv_gsave_alloc(B, H, L, device) if gsave else None
    if gsave:
        assert gs is not None, "gsave requested for a length of the single-kernel regime"
        gs.fill_(float("nan"))          # every entry the backward reads must have been written by the forward
    gen = torch.Generator().manual_seed(seed)
    k = decaying_filter(H, L, gen).requires_grad_(True)
    D = torch.randn(H, generator=gen).requires_grad_(True)
    w = torch.randn(B, H, L, generator=gen)
    dev = lambda t: None if t is None else t.detach().to(device)
    errs = {}
    if mode == "plain":
        u = torch.randn(B, H, L, generator=gen).to(dtype).requires_grad_(True)
        ref = O.fftconv_ref(u, k, D, None, gelu=False)
        (ref.float() * w).sum().backward()
        Kf = K.filter_spectrum(dev(k), dev(D), L)
        out, _ = K.conv_fwd(dev(u), Kf, L, gsave=gs)
        du, _, _, dKacc, dD = K.conv_bwd(dev(w).to(dtype), dev(u), Kf, L, gsave=gs, nslot=nslot)
        dk = K.conv_dk(dKacc, L)
        dD = dk[:, 0] if gsave else dD
        errs = dict(out=(out, ref), du=(du, u.grad), dk=(dk, k.grad), dD=(dD, D.grad))
    elif mode == "gated":
        u, pre, q = (torch.randn(B, H, L, generator=gen).to(dtype).requires_grad_(True) for _ in range(3))
        ref = O.fftconv_h3_ref(u, k, D, q, pre, head_dim=1)
        (ref.float() * w).sum().backward()
        Kf = K.filter_spectrum(dev(k), dev(D), L)
        out, ys = K.conv_fwd(dev(u), Kf, L, in_mode=IN_PREGATE, out_mode=OUT_POSTGATE, pre=dev(pre), post=dev(q), save_y=True,
                             gsave=gs)
        du, dpre, dq, dKacc, dD = K.conv_bwd(dev(w).to(dtype), dev(u), Kf, L, in_mode=IN_PREGATE, out_mode=OUT_POSTGATE,
                                             pre=dev(pre), post=dev(q), ysave=ys, gsave=gs)
        dk = K.conv_dk(dKacc, L)
        dD = dk[:, 0] if gsave else dD
        errs = dict(out=(out, ref), du=(du, u.grad), dpre=(dpre, pre.grad), dq=(dq, q.grad), dk=(dk, k.grad), dD=(dD, D.grad))
    elif mode == "shortconv":
        uT = torch.randn(B, 3 * H, L, generator=gen).to(dtype).requires_grad_(True)
        sw = (torch.randn(3 * H, 1, 3, generator=gen) * 0.5).requires_grad_(True)
        sb = torch.randn(3 * H, generator=gen).requires_grad_(True)
        pb = torch.randn(3 * H, generator=gen).requires_grad_(True)
        x = uT + pb[None, :, None].to(dtype)
        uc = O.short_filter(x, sw.to(dtype), sb.to(dtype), L)
        x0, x1, v = uc.split(H, dim=1)
        y = O.fftconv_ref(v * x1, k, D, None, gelu=False)
        z = y * x0
        (z.float() * w).sum().backward()
        Kf = K.filter_spectrum(dev(k), dev(D), L)
        swc = dev(sw).reshape(3 * H, 3).contiguous()
        out, ys = K.conv_fwd(dev(uT), Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=swc, sb=dev(sb), pb=dev(pb), save_y=True,
                             gsave=gs)
        dzd, uTd = dev(w).to(dtype), dev(uT)
        if defer:       # dx0 = dout * y formed by the short-filter backward (needs rows aligned for 16-byte access)
            ldp = (L + 7) // 8 * 8
            dzd = torch.nn.functional.pad(dzd, (0, ldp - L))[:, :, :L]
            uTd = torch.nn.functional.pad(uTd, (0, ldp - L))[:, :, :L]
            assert K.shortconv_gate_supported(uTd, dzd, ys)
        dX, _, _, dKacc, dD = K.conv_bwd(dzd, uTd, Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=swc,
                                         sb=dev(sb), pb=dev(pb), ysave=ys, gsave=gs, defer_dx0=defer, nslot=nslot)
        dk = K.conv_dk(dKacc, L)
        dD = dk[:, 0] if gsave else dD
        if defer:
            dX[:, :H].fill_(float("nan"))       # the x0 group of dX must not be read
        duT, dsw, dsb, dpb = K.shortconv_bwd(uTd, dX, swc, dev(pb), L, dout=dzd if defer else None, ysave=ys if defer else None)
        xc = K.shortconv_fwd(dev(uT), swc, dev(sb), dev(pb), L)
        errs = dict(xc=(xc, uc), out=(out, z), y=(ys, y), duT=(duT, uT.grad), dsw=(dsw, sw.grad.reshape(3 * H, 3)),
                    dsb=(dsb, sb.grad), dpb=(dpb, pb.grad), dk=(dk, k.grad), dD=(dD, D.grad))
    else:
        raise ValueError(mode)
    return {n: relerr(a, b) for n, (a, b) in errs.items()}


def bf16_forward_case(B, H, L, device, seed=0):
    """bf16 I/O, Hyena gating: ours vs the oracle's own bf16 path, both measured against fp64 truth
    computed from the same bf16-rounded inputs."""
    gen = torch.Generator().manual_seed(seed)
    k = decaying_filter(H, L, gen)
    D = torch.randn(H, generator=gen)
    uT = torch.randn(B, 3 * H, L, generator=gen).to(torch.bfloat16)
    sw = torch.randn(3 * H, 1, 3, generator=gen) * 0.5
    sb = torch.randn(3 * H, generator=gen)

    def chain(dt):
        uc = O.short_filter(uT.to(dt), sw.to(dt), sb.to(dt), L)
        x0, x1, v = uc.split(H, dim=1)
        y = O.fftconv_ref(v * x1, k.to(torch.float64 if dt == torch.float64 else torch.float32),
                          D.to(torch.float64 if dt == torch.float64 else torch.float32), None, gelu=False)
        return (y.to(dt) * x0)

    truth = chain(torch.float64)
    ref_bf16 = chain(torch.bfloat16)
    Kf = K.filter_spectrum(k.to(device), D.to(device), L)
    out, _ = K.conv_fwd(uT.to(device), Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV,
                        sw=sw.reshape(3 * H, 3).contiguous().to(device), sb=sb.to(device), pb=None, save_y=True)
    scale = truth.abs().max().item()
    e_ours = (out.detach().cpu().double() - truth).abs().max().item()
    e_ref = (ref_bf16.double() - truth).abs().max().item()
    return e_ours, e_ref, scale


def filter_case(D, order, emb, n_inner, L, lmax, device, seed=1, shift=0.05):
    gen = torch.Generator().manual_seed(seed)
    z, t = O.positional_tables(emb, lmax)
    w_in = torch.randn(order, emb, generator=gen) * 0.5
    b_in = torch.randn(order, generator=gen) * 0.1
    w_h = torch.randn(n_inner, order, order, generator=gen) * 0.1
    b_h = torch.randn(n_inner, order, generator=gen) * 0.1
    w_out = torch.randn(D, order, generator=gen) * 0.1
    freq = torch.full((order,), 10.0) + torch.randn(order, generator=gen)
    deltas = O.modulation_deltas(D)
    p = {"pos_emb.z": z, "pos_emb.t": t, "implicit_filter.1.freq": freq[None], "modulation.deltas": deltas,
         "implicit_filter.0.weight": w_in, "implicit_filter.0.bias": b_in}
    for i in range(n_inner):
        p[f"implicit_filter.{2 * (i + 1)}.weight"] = w_h[i]
        p[f"implicit_filter.{2 * (i + 1)}.bias"] = b_h[i]
    p[f"implicit_filter.{2 * (n_inner + 1)}.weight"] = w_out
    ref32 = O.hyena_filter(p, L, shift=shift)[0].transpose(0, 1)
    ref64 = O.hyena_filter({kk: vv.double() for kk, vv in p.items()}, L, shift=shift)[0].transpose(0, 1)
    d = lambda x: x.to(device)
    k = K.filter_fwd(d(z[0]), d(t[0]), d(w_in), d(b_in), d(w_h) if n_inner else None, d(b_h) if n_inner else None, d(w_out),
                     d(freq), d(deltas.reshape(-1)), shift, True, L)
    return relerr(k, ref64), relerr(ref32, ref64)


def filter_trunk_saved_case(order, emb, n_inner, L, lmax, device, seed=7):
    """Saved-trunk backward (forward keeps the pre-activations, hy_filter_fwd_save_trunk / hy_filter_trunk_bwd_saved)
    against the recomputing one (hy_filter_trunk_bwd) on the same dh_last: same gradients up to fp32 rounding, and the
    forward's k / h_last are unchanged by the extra stores."""
    gen = torch.Generator().manual_seed(seed)
    z, t = O.positional_tables(emb, lmax)
    w_in = torch.randn(order, emb, generator=gen) * 0.5
    b_in = torch.randn(order, generator=gen) * 0.1
    w_h = torch.randn(n_inner, order, order, generator=gen) * 0.1
    b_h = torch.randn(n_inner, order, generator=gen) * 0.1
    D = 16
    w_out = torch.randn(D, order, generator=gen) * 0.1
    freq = torch.full((order,), 10.0) + torch.randn(order, generator=gen)
    deltas = O.modulation_deltas(D).reshape(-1)
    dh_last = torch.randn(L, order, generator=gen)
    d = lambda x: x.to(device)
    args = (d(z[0]), d(t[0]), d(w_in), d(b_in), d(w_h) if n_inner else None, d(b_h) if n_inner else None, d(w_out), d(freq),
            d(deltas), 0.05, True, L)
    k1, h1 = K.filter_fwd(*args, save_h=True)
    k2, h2, a_save = K.filter_fwd(*args, save_h=True, save_trunk=True)
    assert torch.equal(k1, k2) and torch.equal(h1, h2)
    targs = (d(dh_last), d(z[0]), d(t[0]), d(w_in), d(b_in), d(w_h) if n_inner else None, d(b_h) if n_inner else None,
             d(w_out), d(freq), L)
    g_re = K.filter_trunk_bwd(*targs)
    g_sv = K.filter_trunk_bwd(*targs, a_save=a_save)
    flat = lambda g: torch.cat([x.reshape(-1) for x in (g[0], g[1], *g[2], *g[3], g[4])])
    return relerr(flat(g_sv), flat(g_re))


def filter_out_bwd_case(D, L, device, seed=3, shift=0.05, modulate=True, ragged=False):
    """hy_filter_out_bwd (tensor cores, 3xTF32) against the same two contractions in fp64 (the backward of
    hyena.py:219 + the modulation, hyena.py:156-159); the fp32 torch result gives the error scale."""
    order = 64
    gen = torch.Generator().manual_seed(seed)
    ld = L if not ragged else (L + 3) // 4 * 4 + 4
    dk = torch.randn(D, ld, generator=gen)[:, :L]
    h_last = torch.sin(3.0 * torch.randn(L, order, generator=gen))
    w_out = torch.randn(D, order, generator=gen) * 0.1
    t = torch.linspace(0, 1, L)
    deltas = O.modulation_deltas(D).reshape(-1)

    def ref(dt):
        m = (torch.exp(-t.to(dt)[:, None] * deltas.to(dt).abs()[None]) + shift) if modulate else 1.0
        dh = dk.to(dt).t() * m
        return dh @ w_out.to(dt), dh.t() @ h_last.to(dt)

    r64, r32 = ref(torch.float64), ref(torch.float32)
    d = lambda x: x.to(device)
    dkd = torch.randn(D, ld, device=device)[:, :L]   # ragged: padded row stride, garbage beyond L
    dkd.copy_(d(dk))
    got = K.filter_out_bwd(dkd, d(t), d(deltas), shift, modulate, d(w_out), d(h_last), L)
    return [(relerr(g, a), relerr(b, a)) for g, a, b in zip(got, r64, r32)]


def channel_slab_case(Dm, L, world, device, dtype=torch.float32, seed=5):
    """SURVEY 8e "Channels, B = 1": the fused operator on the channel slabs of `world` ranks (dp.channel_slab),
    concatenated, must equal the operator on all channels — outputs and every gradient, bit for bit (rows never
    interact; this is what makes the channel partition collective-free)."""
    from dna_b200.dp import channel_slab
    from dna_b200.fftconv import fftconv_func
    gen = torch.Generator().manual_seed(seed)
    mk = lambda *s: torch.randn(*s, generator=gen)
    x0, x1, v, dz = (mk(1, Dm, L).to(dtype).to(device) for _ in range(4))
    k = (decaying_filter(Dm, L, gen)).to(device)
    Dskip = mk(Dm).to(device)

    def run(lo, hi):
        ins = [t[:, lo:hi].contiguous().requires_grad_(True) for t in (x0, x1, v)]
        kk, dd = k[lo:hi].contiguous().requires_grad_(True), Dskip[lo:hi].contiguous().requires_grad_(True)
        out = fftconv_func(ins[1], kk, dd, dropout_mask=None, gelu=False, v=ins[2], q=ins[0])
        out.backward(dz[:, lo:hi].contiguous())
        return [out.detach()] + [t.grad for t in ins], [kk.grad, dd.grad]

    full_a, full_p = run(0, Dm)
    parts = [run(*channel_slab(Dm, r, world)) for r in range(world)]
    for i, f in enumerate(full_a):
        assert torch.equal(torch.cat([p_[0][i] for p_ in parts], dim=1), f), ("activation", i)
    for i, f in enumerate(full_p):
        assert torch.equal(torch.cat([p_[1][i] for p_ in parts], dim=0), f), ("parameter", i)
    return True


def tokenizer_case(B, maxchars, max_length, flags, device, seed=0):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTNacgtn.X", dtype=np.uint8)
    lens = rng.integers(0, maxchars + 1, size=B).astype(np.int32)
    arr = alphabet[rng.integers(0, len(alphabet), size=(B, maxchars))].astype(np.uint8)
    ids = K.tokenize(torch.from_numpy(arr).to(device), torch.from_numpy(lens).to(device), max_length, flags).cpu()
    for i in range(B):
        s = bytes(arr[i, :lens[i]]).decode()
        r = torch.tensor(O.tokenize_ref(s, max_length, add_special_tokens=bool(flags & 3), cls_token=bool(flags & 2)))
        if flags & 4:
            r[r == 11] = 4
        if flags & 8:
            r = r - 7
            r[(r >= 4) | (r < 0)] = 4
        assert torch.equal(ids[i], r), (i, s, ids[i].tolist(), r.tolist())
    return True


def revcomp_case(B, maxchars, device, seed=0, with_apply=True, golden=None):
    """Reverse-complement kernel, bit-exact against the oracle (and the reference's golden strings when given)."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTNacgtn.XRY-", dtype=np.uint8)
    lens = rng.integers(0, maxchars + 1, size=B).astype(np.int32)
    if B > 0:
        lens[0] = maxchars
    arr = alphabet[rng.integers(0, len(alphabet), size=(B, maxchars))].astype(np.uint8)
    apply = (rng.integers(0, 2, size=B).astype(np.uint8) if with_apply else None)
    out = K.reverse_complement(torch.from_numpy(arr).to(device), torch.from_numpy(lens).to(device),
                               None if apply is None else torch.from_numpy(apply).to(device)).cpu().numpy()
    for i in range(B):
        s = bytes(arr[i, :lens[i]]).decode()
        want = O.reverse_complement_ref(s) if (apply is None or apply[i]) else s
        assert bytes(out[i, :lens[i]]).decode() == want, (i, s)
        assert np.array_equal(out[i, lens[i]:], arr[i, lens[i]:])       # tail copied
    if golden is not None:
        n = len([k for k in golden.files if k.startswith("in")])
        for i in range(n):
            x = np.asarray(golden[f"in{i}"])
            if x.size == 0:
                continue
            got = K.reverse_complement(torch.from_numpy(x.copy())[None].to(device)).cpu().numpy()[0]
            assert np.array_equal(got, np.asarray(golden[f"out{i}"])), i
    return True


def filter_bwd_case(D, L, device, order=64, emb=5, n_inner=2, seed=3, init="default", shift=0.05):
    """HyenaFilter backward at production length: every parameter gradient of k = filter(L) (hyena.py:203-242) for a
    random dk, ours (hy_filter_out_bwd / hy_filter_trunk_bwd through dna_b200.hyena.HyenaFilter) against the oracle's
    autograd in fp64; the oracle's own fp32 autograd gives the error scale.  Returns {name: (e_ours, e_ref32)}.
    init "default": nn.Linear's own init (large weights, sin(10 x) far into its oscillating range);
    init "backbone": the N(0, 0.02) / zero-bias re-draw of `_init_weights` (standalone_hyenadna.py:612-641)."""
    from dna_b200.hyena import HyenaFilter
    torch.manual_seed(seed)
    f = HyenaFilter(D, emb_dim=emb, order=order, seq_len=L + 2, w=10, lr_pos_emb=0.0, num_inner_mlps=n_inner, shift=shift)
    if init == "backbone":
        for m in f.implicit_filter:
            if isinstance(m, torch.nn.Linear):
                torch.nn.init.normal_(m.weight, std=0.02)
                if m.bias is not None:
                    torch.nn.init.zeros_(m.bias)
    sd = {k: v.detach().clone() for k, v in f.state_dict().items()}
    gen = torch.Generator().manual_seed(seed + 1)
    dk = torch.randn(D, L, generator=gen)

    def ref(dt):
        p = {k: v.to(dt).requires_grad_(k.startswith("implicit_filter")) for k, v in sd.items()}
        h = O.hyena_filter(p, L, shift=shift)[0]                       # [L, D]
        names = [k for k in p if p[k].requires_grad and not (k.endswith(".freq") and not k.endswith("1.freq"))]
        gr = torch.autograd.grad(h, [p[k] for k in names], dk.t().to(dt))
        return dict(zip(names, gr))

    r64, r32 = ref(torch.float64), ref(torch.float32)
    f = f.to(device)
    k = f.filter_cm(L)
    k.backward(dk.to(device))
    ours = {n: p_.grad for n, p_ in f.named_parameters() if n.startswith("implicit_filter")}
    out = {}
    for n in r64:
        assert ours[n] is not None, n
        out[n] = (relerr(ours[n], r64[n]), relerr(r32[n], r64[n]))
    return out


def bf16_conv_truth_case(B, H, L, device, seed=0, gsave=True):
    """bf16 activations, Hyena gating, forward AND backward: ours vs the oracle's own bf16 path, both measured against
    fp64 truth computed from the same bf16-rounded inputs.  Returns {name: (e_ours, e_ref_bf16, scale)} (absolute
    max errors): we may not be further from the exact answer than the reference's own bf16 path (+ one output ulp)."""
    gen = torch.Generator().manual_seed(seed)
    k = decaying_filter(H, L, gen)
    D = torch.randn(H, generator=gen)
    uT = torch.randn(B, 3 * H, L, generator=gen).to(torch.bfloat16)
    sw = torch.randn(3 * H, 1, 3, generator=gen) * 0.5
    sb = torch.randn(3 * H, generator=gen)
    w = torch.randn(B, H, L, generator=gen).to(torch.bfloat16)

    def chain(dt):
        kd = torch.float64 if dt == torch.float64 else torch.float32
        ins = dict(uT=uT.to(dt).requires_grad_(True), sw=sw.to(dt).requires_grad_(True), sb=sb.to(dt).requires_grad_(True),
                   k=k.to(kd).requires_grad_(True), D=D.to(kd).requires_grad_(True))
        uc = O.short_filter(ins["uT"], ins["sw"], ins["sb"], L)
        x0, x1, v = uc.split(H, dim=1)
        y = O.fftconv_ref(v * x1, ins["k"], ins["D"], None, gelu=False)
        z = y.to(dt) * x0
        gr = torch.autograd.grad(z, list(ins.values()), w.to(dt))
        res = dict(out=z.detach(), y=y.detach())
        res.update({"d" + n: g for n, g in zip(ins, gr)})
        return res

    truth, ref = chain(torch.float64), chain(torch.bfloat16)
    gs = K.conv_gsave_alloc(B, H, L, device) if gsave else None
    dev = lambda t: t.to(device)
    swc = dev(sw).reshape(3 * H, 3).contiguous()
    Kf = K.filter_spectrum(dev(k), dev(D), L)
    out, ys = K.conv_fwd(dev(uT), Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=swc, sb=dev(sb), pb=None, save_y=True,
                         gsave=gs)
    dX, _, _, dKacc, dD = K.conv_bwd(dev(w), dev(uT), Kf, L, in_mode=IN_SHORTCONV, out_mode=OUT_SHORTCONV, sw=swc, sb=dev(sb),
                                     pb=None, ysave=ys, gsave=gs)
    dk = K.conv_dk(dKacc, L)
    dD = dk[:, 0] if gs is not None else dD
    duT, dsw, dsb, _ = K.shortconv_bwd(dev(uT), dX, swc, None, L)
    ours = dict(out=out, y=ys, duT=duT, dsw=dsw.reshape(3 * H, 1, 3), dsb=dsb, dk=dk, dD=dD)
    res = {}
    for n, got in ours.items():
        t = truth[n].double()
        res[n] = ((got.detach().cpu().double() - t).abs().max().item(), (ref[n].double() - t).abs().max().item(),
                  t.abs().max().item())
    return res


def model_step_case(n_layer, d_model, L, B, device, autocast=True, seed=0):
    """One model-level training step (BASELINE config shapes): next-token loss and gradients of
    dna_b200.standalone.HyenaDNAModel (+ tied head) against oracle.hyena_model_oracle.lm_loss with the same state_dict.
    Under bf16 autocast the comparison is against the oracle's fp32 answer (pin bf16 against fp32 truth).
    Returns (loss_ours, loss_ref, {param name: relative error of its gradient})."""
    import torch.nn.functional as F
    from dna_b200.standalone import HyenaDNAModel
    from oracle import hyena_model_oracle as MO
    torch.manual_seed(seed)
    model = HyenaDNAModel(d_model=d_model, n_layer=n_layer, d_inner=4 * d_model, vocab_size=12, embed_dropout=0.0,
                          pad_vocab_size_multiple=8,
                          layer=dict(l_max=L + 2, emb_dim=5, filter_order=64, short_filter_order=3, modulate=True, w=10,
                                     lr=6e-4, wd=0.0, lr_pos_emb=0.0))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    gen = torch.Generator().manual_seed(seed + 1)
    ids = torch.randint(7, 11, (B, L + 1), generator=gen)
    data, target = ids[:, :-1], ids[:, 1:]
    psd = {k: v.requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    loss_ref = MO.lm_loss(data, target, psd, n_layer=n_layer, l_max=L + 2, shift=0.05)
    loss_ref.backward()
    model = model.to(device)
    with torch.autocast(device, dtype=torch.bfloat16, enabled=autocast and device != "cpu"):
        h = model(data.to(device))
        logits = F.linear(h, model.backbone.embeddings.word_embeddings.weight.to(h.dtype))
    loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), target.reshape(-1).to(device))
    loss.backward()
    errs = {}
    for n, p_ in model.named_parameters():
        if p_.grad is not None and psd[n].grad is not None:
            errs[n] = relerr(p_.grad, psd[n].grad)
    return loss.item(), loss_ref.item(), errs


# ---- SURVEY section 8(f) rows built in round 2: keyword surface of fftconv_func, operator options, LongConv --------------
def _T(a):
    import numpy as np
    return torch.from_numpy(np.asarray(a))


def fftconv_variant_case(g, tag, variant, device):
    """fftconv_func(k_rev=, bidirectional=) against the reference's own outputs and autograd gradients (features.npz)."""
    from dna_b200.fftconv import fftconv_func
    u, k, kr, D = (_T(g[f"{tag}_{n}"]).to(device).requires_grad_(True) for n in ("u", "k", "krev", "D"))
    kw = dict(bidirectional="bidir" in variant, k_rev=kr if "krev" in variant else None)
    y = fftconv_func(u, k, D, None, False, **kw)
    errs = {"y": relerr(y, _T(g[f"{tag}_{variant}_y"]))}
    ins = [u, k, D] + ([kr] if "krev" in variant else [])
    for name, gr in zip(["du", "dk", "dD", "dkrev"], torch.autograd.grad((y * _T(g[f"{tag}_w"]).to(device)).sum(), ins)):
        errs[name] = relerr(gr, _T(g[f"{tag}_{variant}_{name}"]))
    return errs


def h3_heads_case(g, hd, device):
    """fftconv_func(head_dim > 1) — the H3 multi-head form (src/ops/fftconv.py:38-55) — against the reference."""
    from dna_b200.fftconv import fftconv_func
    pre = f"h3_hd{hd}_"
    k, v, q, ssm, D = (_T(g[pre + n]).to(device).requires_grad_(True) for n in ("k", "v", "q", "ssm", "D"))
    y = fftconv_func(k, ssm, D, None, False, False, False, v, hd, q)
    errs = {"y": relerr(y, _T(g[pre + "y"]))}
    gr = torch.autograd.grad((y * _T(g[pre + "w"]).to(device)).sum(), [k, ssm, D, q, v])
    for name, t in zip(["dk", "dssm", "dD", "dq", "dv"], gr):
        errs[name] = relerr(t, _T(g[pre + name]))
    return errs


OPTION_KW = {
    "blocks2": dict(d_model=8, l_max=64, kw=dict(num_blocks=2, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
    "blocks4_o3": dict(d_model=4, l_max=96, kw=dict(num_blocks=4, order=3, emb_dim=3, filter_order=16, w=2, lr_pos_emb=0)),
    "outer": dict(d_model=6, l_max=50, kw=dict(outer_mixing=True, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
    "ffn_o3": dict(d_model=6, l_max=48, kw=dict(post_order_ffn=True, order=3, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
    "short5": dict(d_model=8, l_max=70, kw=dict(short_filter_order=5, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
    "drop": dict(d_model=8, l_max=64, train=True, kw=dict(dropout=0.25, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
    "gelu_act": dict(d_model=8, l_max=40, kw=dict(activation="gelu", emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
    "bidir_src": dict(d_model=8, l_max=72, kw=dict(bidirectional=True, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
}
LONGCONV_KW = {
    "lc_causal": dict(d_model=8, l_max=64, channels=1, lam=0.001),
    "lc_bidir": dict(d_model=6, l_max=50, channels=1, bidirectional=True, lam=0.001),
    "lc_ch2_bld": dict(d_model=8, l_max=64, channels=2, transposed=False, lam=0.0005, postact="glu", activation="gelu"),
    "lc_bidir_short": dict(d_model=4, l_max=64, channels=2, bidirectional=True, lam=0.001, postact=None),
}


def module_vs_golden(g, tag, module, device, train=False, seed=None):
    """load the reference's state_dict strictly, run forward + backward, compare with its outputs and gradients"""
    pre = f"{tag}/sd/"
    sd = {k[len(pre):]: _T(g[k]) for k in g.files if k.startswith(pre)}
    assert set(module.state_dict().keys()) == set(sd.keys()), (sorted(module.state_dict()), sorted(sd))
    module.load_state_dict(sd, strict=True)
    module = module.to(device)
    module.train(train)
    u = _T(g[f"{tag}/u"]).to(device).requires_grad_(True)
    if seed is not None:
        torch.manual_seed(seed)
    y = module(u)
    y = y[0] if isinstance(y, tuple) else y
    assert y.shape == tuple(g[f"{tag}/y"].shape)
    (y * _T(g[f"{tag}/w"]).to(device)).sum().backward()
    errs = {"y": relerr(y, _T(g[f"{tag}/y"])), "du": relerr(u.grad, _T(g[f"{tag}/du"]))}
    params = dict(module.named_parameters())
    gp = f"{tag}/grad/"
    for key in g.files:
        if key.startswith(gp):
            errs[key[len(gp):]] = relerr(params[key[len(gp):]].grad, _T(g[key]))
    return errs


def operator_option_case(g, tag, device):
    from dna_b200.hyena import HyenaOperator
    c = OPTION_KW[tag]
    op = HyenaOperator(d_model=c["d_model"], l_max=c["l_max"], layer_idx=0, device=None, dtype=None, **c["kw"])
    return module_vs_golden(g, tag, op, device, train=bool(c.get("train")), seed=77)


def long_conv_case(g, tag, device):
    from dna_b200.long_conv import LongConv
    return module_vs_golden(g, tag, LongConv(**LONGCONV_KW[tag]), device)


# ---- data ingest: interval fetch + BERT masking kernels, bit-exact ------------------------
def fetch_intervals_case(g, device):
    """K.fetch_intervals against the reference's FastaInterval outputs (ingest.npz) and, for random intervals with
    reverse complement, against the oracle."""
    import numpy as np
    chrom_np = np.asarray(g["chrom"])
    chrom = torch.from_numpy(chrom_np.copy()).to(device)
    cases = g["fetch/cases"].tolist()
    n = 0
    for pad in (0, 1):
        by_len = {}
        for i, (s0, e0, ml) in enumerate(cases):
            by_len.setdefault(ml, []).append((i, s0, e0))
        for ml, rows in by_len.items():
            out, lens = K.fetch_intervals(chrom, torch.tensor([r[1] for r in rows]), torch.tensor([r[2] for r in rows]), ml,
                                          pad_interval=bool(pad))
            for (i, _, _), row, ln in zip(rows, out.cpu(), lens.cpu().tolist()):
                want = bytes(g[f"fetch/pad{pad}/{i}"])
                assert ln == min(len(want), ml) and bytes(row[:ln].numpy()) == want[:ln], (pad, i)
                assert (row[ln:] == ord(".")).all()
                n += 1
    # random intervals + reverse complement vs the oracle
    rng = np.random.default_rng(0)
    chrom_s = bytes(chrom_np).decode()
    B, ml = 64, 777
    starts = rng.integers(-0, len(chrom_s) - 10, size=B)
    ends = starts + rng.integers(1, 1500, size=B)
    ends = np.minimum(ends, len(chrom_s) + 0)
    rc = rng.random(B) > 0.5
    out, lens = K.fetch_intervals(chrom, torch.from_numpy(starts), torch.from_numpy(ends), ml, rc=torch.from_numpy(rc), pad_interval=True)
    for b in range(B):
        want = O.fetch_interval_ref(chrom_s, int(starts[b]), int(ends[b]), ml, pad_interval=True, reverse_complement=bool(rc[b])).encode()
        ln = int(lens[b])
        assert ln == min(len(want), ml) and bytes(out[b, :ln].cpu().numpy()) == want[:ln], b
        n += 1
    return n


def bert_mask_case(g, i, device):
    """K.bert_mask fed the oracle's own draws: bit-exact with the reference's outputs (ingest.npz)."""
    seq = _T(g[f"bert/{i}/seq"])
    torch.manual_seed(500 + i)
    (o, m, l), (r_mask, r_kind, rtok) = O.bert_mask_ref(seq, 3, 4, int(g[f"bert/{i}/vocab"]), special_token_ids=g[f"bert/{i}/special"].tolist())
    out, mask, labels = K.bert_mask(seq.to(device), r_mask.to(device), r_kind.to(device), rtok.to(device), 3, 4)
    assert torch.equal(out.cpu(), _T(g[f"bert/{i}/out"])) and torch.equal(mask.cpu(), _T(g[f"bert/{i}/mask"]))
    assert torch.equal(labels.cpu(), _T(g[f"bert/{i}/labels"]))
    return True


def add_ln_dropout_case(device, xdt=torch.float32, rdt=torch.float32, with_res=True, D=256, rows=37, p=0.1, seed=0):
    """fused dropout -> add -> LayerNorm against the oracle restatement of standalone_hyenadna.py:521-525 for a GIVEN
    keep mask (values, stream, and every gradient)."""
    from dna_b200 import block_ops
    from oracle.hyena_model_oracle import block_add_norm
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, D, generator=g).to(xdt)
    res = (torch.randn(rows, D, generator=g) * 3).to(rdt) if with_res else None
    w, b = torch.randn(D, generator=g), torch.randn(D, generator=g)
    keep = torch.rand(rows, D, generator=g) >= p
    norm = torch.nn.LayerNorm(D, eps=1e-5)
    with torch.no_grad():
        norm.weight.copy_(w); norm.bias.copy_(b)
    norm = norm.to(device)
    xr = x.clone().to(device).requires_grad_(True)
    rr = None if res is None else res.clone().to(device).requires_grad_(True)
    y, ro = block_ops.add_layer_norm(xr, rr, norm, dropout_p=p, keep_mask=keep.to(device))
    xo = x.clone().requires_grad_(True)
    ro_in = None if res is None else res.clone().requires_grad_(True)
    wo, bo = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y_ref, r_ref = block_add_norm(xo, ro_in, wo, bo, 1e-5, keep_mask=keep, dropout_p=p)
    errs = {"y": relerr(y, y_ref), "stream_equal": float(not torch.equal(ro.detach().cpu(), r_ref.detach()))}
    gy, gr = torch.randn(y_ref.shape, generator=g), torch.randn(r_ref.shape, generator=g)
    ((y.float() * gy.to(device)).sum() + (ro.float() * gr.to(device)).sum()).backward()
    ((y_ref.float() * gy).sum() + (r_ref.float() * gr).sum()).backward()
    errs["dx"] = relerr(xr.grad, xo.grad)
    if res is not None:
        errs["dres"] = relerr(rr.grad, ro_in.grad)
    errs["dgamma"] = relerr(norm.weight.grad, wo.grad)
    errs["dbeta"] = relerr(norm.bias.grad, bo.grad)
    # the dropped positions carry no gradient
    assert (xr.grad.cpu()[~keep] == 0).all()
    return errs
