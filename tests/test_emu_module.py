"""No-GPU suite, module level: dna_b200.hyena.HyenaOperator / HyenaFilter / fftconv_func / the
HyenaDNA harness (running the kernel sources under the CPU emulator) against OUTPUTS OF THE REFERENCE
ITSELF (tests/golden/*.npz) with the reference's state_dict loaded strictly — the drop-in contract of
SURVEY §8b."""
import os

import numpy as np
import pytest
import torch

import parity_cases as P

CFGS = {
    "src": dict(mod="src", d_model=16, l_max=70, kw=dict(emb_dim=5, filter_order=64, w=10, lr=6e-4, wd=0, lr_pos_emb=0, modulate=True)),
    "src_e3": dict(mod="src", d_model=8, l_max=40, kw=dict(emb_dim=3, filter_order=16, w=1, lr_pos_emb=0)),
    "sa": dict(mod="sa", d_model=16, l_max=130, kw=dict(emb_dim=5, filter_order=64, w=10, lr=6e-4, wd=0, lr_pos_emb=0)),
    "sa_trunc": dict(mod="sa", d_model=8, l_max=50, kw=dict(emb_dim=5, filter_order=64, w=10, lr_pos_emb=0)),
}
# round-2 fixtures (tests/golden/features.npz): order > 2 in both filter-channel layouts
FEATS = {
    "o3_src": dict(mod="src", d_model=8, l_max=64, kw=dict(order=3, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
    "o3_sa": dict(mod="sa", d_model=8, l_max=64, kw=dict(order=3, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
    "o4_src": dict(mod="src", d_model=4, l_max=40, kw=dict(order=4, emb_dim=3, filter_order=16, w=2, lr_pos_emb=0)),
}
T = lambda a: torch.from_numpy(np.asarray(a))


def build_operator(tag, device="cpu"):
    from dna_b200.hyena import HyenaOperator, standalone_hyena_operator
    c = CFGS[tag] if tag in CFGS else FEATS[tag]
    if c["mod"] == "src":
        op = HyenaOperator(d_model=c["d_model"], l_max=c["l_max"], layer_idx=0, device=None, dtype=None, **c["kw"])
    else:
        op = standalone_hyena_operator(d_model=c["d_model"], l_max=c["l_max"], **c["kw"])
    return op.to(device)


def operator_vs_golden(tag, g, device):
    op = build_operator(tag, device)
    pre = f"{tag}/sd/"
    sd = {k[len(pre):]: T(g[k]) for k in g.files if k.startswith(pre)}
    assert set(op.state_dict().keys()) == set(sd.keys())           # incl. aliased .1/.3/.5.freq and buffers
    op.load_state_dict(sd, strict=True)
    u = T(g[f"{tag}/u"]).to(device).requires_grad_(True)
    y = op(u)
    assert y.shape == tuple(g[f"{tag}/y"].shape)
    (y * T(g[f"{tag}/w"]).to(device)).sum().backward()
    errs = {"y": P.relerr(y, T(g[f"{tag}/y"])), "du": P.relerr(u.grad, T(g[f"{tag}/du"]))}
    params = dict(op.named_parameters())
    gp = f"{tag}/grad/"
    for key in g.files:
        if key.startswith(gp):
            errs[key[len(gp):]] = P.relerr(params[key[len(gp):]].grad, T(g[key]))
    return errs, op


@pytest.mark.parametrize("tag", list(CFGS))
def test_operator_matches_reference(emu_lib, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "operator.npz"))
    errs, op = operator_vs_golden(tag, g, "cpu")
    for name, e in errs.items():
        assert e <= 5e-5, (tag, name, e)       # sin(10x) filter: the reference's own fp32 noise is ~1e-5
    # optimizer hyper-parameter tags the training script groups on (train.py:468-487)
    for name, p in op.filter_fn.implicit_filter.named_parameters():
        assert p._optim == {"weight_decay": CFGS[tag]["kw"].get("wd", 0), "lr": CFGS[tag]["kw"].get("lr", 1e-3)}, name
    assert isinstance(op.in_proj, torch.nn.Linear) and isinstance(op.out_proj, torch.nn.Linear)
    assert op.d_output == CFGS[tag]["d_model"]


def test_filter_api_matches_reference(emu_lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "operator.npz"))
    for tag in CFGS:
        op = build_operator(tag)
        pre = f"{tag}/sd/"
        op.load_state_dict({k[len(pre):]: T(g[k]) for k in g.files if k.startswith(pre)}, strict=True)
        ref = T(g[f"{tag}/filter"])
        k = op.filter_fn.filter(ref.shape[1])
        assert k.shape == ref.shape
        assert P.relerr(k, ref) <= 5e-5


def test_fftconv_func_matches_reference(emu_lib, golden_dir):
    from dna_b200.fftconv import fftconv_func, fftconv_ref
    g = np.load(os.path.join(golden_dir, "fftconv.npz"))
    for tag in "abc":
        u = T(g[f"{tag}_u"]).requires_grad_(True)
        k = T(g[f"{tag}_k"]).requires_grad_(True)
        D = T(g[f"{tag}_D"]).requires_grad_(True)
        out = fftconv_func(u, k, D, None, False)
        assert P.relerr(out, T(g[f"{tag}_hy_nogelu"])) <= P.FP32_TOL
        (out * T(g[f"{tag}_w"])).sum().backward()
        assert P.relerr(u.grad, T(g[f"{tag}_du"])) <= P.FP32_TOL
        assert P.relerr(k.grad, T(g[f"{tag}_dk"])) <= P.FP32_TOL
        assert P.relerr(D.grad, T(g[f"{tag}_dD"])) <= P.FP32_TOL
        assert P.relerr(fftconv_func(u.detach(), k.detach(), D.detach(), None, True), T(g[f"{tag}_hy_gelu"])) <= P.FP32_TOL
        # 5-D call shape of HyenaOperator (hyena.py:447-453, 484)
        out5 = fftconv_func(u.detach()[:, None, :, None, :], k.detach(), D.detach()[None, :, None], None, False)
        assert P.relerr(out5, T(g[f"{tag}_hy_5d"])) <= P.FP32_TOL
        # the module keeps the name `fftconv_ref` (hyena.py:12-17 imports it) on the same kernels
        assert P.relerr(fftconv_ref(u.detach(), k.detach(), D.detach(), None, gelu=False), T(g[f"{tag}_hy_nogelu"])) <= P.FP32_TOL
    kk, vv, qq, ssm, Dh = (T(g[n]) for n in ("h3_k", "h3_v", "h3_q", "h3_ssm", "h3_D"))
    out = fftconv_func(kk, ssm, Dh, None, False, False, False, vv, 1, qq)
    assert P.relerr(out, T(g["h3_out_hd1"])) <= P.FP32_TOL


def test_tiny_model_matches_reference(emu_lib, golden_dir):
    """BASELINE config C1 path (standalone tiny HyenaDNA) with the reference's weights."""
    from dna_b200.standalone import HyenaDNAModel
    g = np.load(os.path.join(golden_dir, "model_tiny.npz"))
    model = HyenaDNAModel(d_model=32, n_layer=2, d_inner=128, vocab_size=12, embed_dropout=0.0,
                          layer=dict(l_max=258, emb_dim=5, filter_order=64, short_filter_order=3, modulate=True, w=10,
                                     lr=6e-4, wd=0.0, lr_pos_emb=0.0))
    model.load_state_dict({k[3:]: T(g[k]) for k in g.files if k.startswith("sd/")}, strict=True)
    model.eval()
    h = model(T(g["ids"]))
    assert P.relerr(h, T(g["hidden"])) <= 1e-5
    loss = h.float().pow(2).mean()
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-6
    assert P.relerr(model.backbone.layers[0].mixer.in_proj.weight.grad, T(g["grad/backbone.layers.0.mixer.in_proj.weight"])) <= 5e-5
    assert P.relerr(model.backbone.layers[1].mixer.filter_fn.bias.grad, T(g["grad/backbone.layers.1.mixer.filter_fn.bias"])) <= 5e-5
    assert P.relerr(model.backbone.embeddings.word_embeddings.weight.grad, T(g["grad/backbone.embeddings.word_embeddings.weight"])) <= 5e-5


@pytest.mark.parametrize("tag", list(FEATS))
def test_order_gt2_matches_reference(emu_lib, golden_dir, tag):
    """order > 2 recurrence (hyena.py:475-484, standalone :286-288) against the reference's own outputs and gradients
    with its state_dict loaded strictly (tests/golden/features.npz); GPU twin in test_gpu_parity.py."""
    g = np.load(os.path.join(golden_dir, "features.npz"))
    errs, _ = operator_vs_golden(tag, g, "cpu")
    for name, e in errs.items():
        assert e <= 5e-5, (tag, name, e)


def test_unsupported_options_raise():
    """num_heads > 1 and inner_factor != 1 do not run in the reference either (its own forward raises)."""
    from dna_b200.hyena import HyenaOperator
    for kw in (dict(num_heads=2), dict(inner_factor=2), dict(jit_filter=True), dict(fused_bias_fc=True)):
        with pytest.raises(NotImplementedError):
            HyenaOperator(d_model=8, l_max=16, **kw)


@pytest.mark.parametrize("tag", ["bi_a", "bi_b", "bi_c"])
@pytest.mark.parametrize("variant", ["bidir", "krev", "bidir_krev"])
def test_fftconv_func_krev_bidirectional(emu_lib, golden_dir, tag, variant):
    g = np.load(os.path.join(golden_dir, "features.npz"))
    for name, e in P.fftconv_variant_case(g, tag, variant, "cpu").items():
        assert e <= 5e-5, (tag, variant, name, e)


@pytest.mark.parametrize("hd", [2, 8])
def test_fftconv_func_h3_heads(emu_lib, golden_dir, hd):
    g = np.load(os.path.join(golden_dir, "features.npz"))
    for name, e in P.h3_heads_case(g, hd, "cpu").items():
        assert e <= 5e-5, (hd, name, e)


@pytest.mark.parametrize("tag", [t for t in P.OPTION_KW if t != "bidir_src"])
def test_operator_options_match_reference(emu_lib, golden_dir, tag):
    """num_blocks / outer_mixing / post_order_ffn / short_filter_order / dropout (same CPU RNG stream as the reference) /
    activation: the reference's outputs and gradients with its state_dict loaded strictly (options.npz)."""
    g = np.load(os.path.join(golden_dir, "options.npz"))
    for name, e in P.operator_option_case(g, tag, "cpu").items():
        assert e <= 5e-5, (tag, name, e)


def test_bidirectional_operator_matches_reference(emu_lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "features.npz"))
    for name, e in P.operator_option_case(g, "bidir_src", "cpu").items():
        assert e <= 5e-5, (name, e)


@pytest.mark.parametrize("tag", list(P.LONGCONV_KW))
def test_long_conv_matches_reference(emu_lib, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "options.npz"))
    for name, e in P.long_conv_case(g, tag, "cpu").items():
        assert e <= 5e-5, (tag, name, e)


def test_output_hbl_layout(emu_lib):
    from dna_b200.fftconv import fftconv_func
    gen = torch.Generator().manual_seed(3)
    u, k, D = torch.randn(3, 4, 70, generator=gen), torch.randn(4, 70, generator=gen), torch.randn(4, generator=gen)
    a = fftconv_func(u, k, D, None, False)
    b = fftconv_func(u, k, D, None, False, False, True)
    assert torch.equal(a, b) and b.transpose(0, 1).is_contiguous() and b.shape == a.shape


@pytest.mark.parametrize("gated", [False, True])
def test_fftconv_func_four_step_saved_spectrum(emu_lib, gated):
    """fftconv_func at a four-step length: the autograd Function keeps the spectrum of the (gated) input for its
    backward; values and every gradient against the oracle's autograd."""
    from dna_b200.fftconv import fftconv_func
    from oracle import hyena_oracle as O
    gen = torch.Generator().manual_seed(5)
    B, H, L = 2, 2, 4500
    u, v, q = (torch.randn(B, H, L, generator=gen).requires_grad_(True) for _ in range(3))
    k = P.decaying_filter(H, L, gen).requires_grad_(True)
    D = torch.randn(H, generator=gen).requires_grad_(True)
    w = torch.randn(B, H, L, generator=gen)
    ins = [u, k, D] + ([v, q] if gated else [])
    ref = O.fftconv_h3_ref(u, k, D, q, v, head_dim=1) if gated else O.fftconv_ref(u, k, D, None, gelu=False)
    gref = torch.autograd.grad((ref * w).sum(), ins)
    out = fftconv_func(u, k, D, None, False, False, False, v if gated else None, 1, q if gated else None)
    gout = torch.autograd.grad((out * w).sum(), ins)
    assert P.relerr(out, ref) <= P.FP32_TOL
    for name, a, b in zip(["du", "dk", "dD", "dv", "dq"], gout, gref):
        assert P.relerr(a, b) <= P.FP32_TOL, (name, P.relerr(a, b))


def inference_filter_cache_case(device):
    """Under no_grad the operator reuses the filter spectrum; any in-place parameter update invalidates it."""
    from dna_b200 import kernels as K
    op = build_operator("sa", device)
    u = torch.randn(2, 100, 16, generator=torch.Generator().manual_seed(4)).to(device)
    y_train = op(u)                                   # grad mode: no cache involved
    with torch.no_grad():
        n0 = K.launch_count()
        y1 = op(u)
        n1 = K.launch_count()
        y2 = op(u)
        n2 = K.launch_count()
    assert op._kf_cache is not None
    assert (n2 - n1) < (n1 - n0), "second no_grad call must skip the filter + spectrum kernels"
    assert torch.equal(y1, y2) and P.relerr(y1, y_train) <= 1e-6
    with torch.no_grad():
        op.filter_fn.bias.add_(0.5)                   # optimizer-style in-place update
        y3 = op(u)
        op.cache_filter_spectrum = False
        y4 = op(u)
    assert not torch.equal(y3, y1) and torch.equal(y3, y4)
    op.cache_filter_spectrum = True
    assert P.relerr(op(u), y3) <= 1e-6                # training-mode forward agrees with the refreshed cache
    # state that does not live in a tensor version counter: `.data` writes, python attributes, explicit invalidation
    with torch.no_grad():
        y5 = op(u)
        op.filter_fn.bias.data.mul_(2.0)              # .data bypasses the version counter: call invalidate
        op.invalidate_filter_cache()
        y6 = op(u)
        assert not torch.equal(y5, y6)
        op.filter_fn.modulation.shift = 0.5           # python-level state is part of the key
        y7 = op(u)
        assert not torch.equal(y6, y7)
    op.train()
    assert op._kf_cache is None                       # train() drops the cache


def test_inference_filter_cache(emu_lib, golden_dir):
    inference_filter_cache_case("cpu")


def test_filter_reuse_across_micro_batches(emu_lib):
    """filter_reuse: two gradient-accumulation micro-batches share ONE filter generation / spectrum and ONE filter
    backward (run by flush_filter_grads) — same gradients as the reference-style regeneration per forward."""
    from dna_b200 import kernels as K
    from dna_b200.hyena import flush_filter_grads
    g = torch.Generator().manual_seed(6)
    us = [torch.randn(2, 100, 16, generator=g) for _ in range(2)]
    ws = [torch.randn(2, 100, 16, generator=g) for _ in range(2)]

    def run(reuse):
        torch.manual_seed(0)
        op = build_operator("sa")
        op.filter_reuse = reuse
        n0 = K.launch_count()
        for u, w in zip(us, ws):
            (op(u) * w).sum().backward()
        flush_filter_grads(op)
        return {n: p.grad.clone() for n, p in op.named_parameters() if p.grad is not None}, K.launch_count() - n0

    ref, n_ref = run(False)
    got, n_got = run(True)
    assert n_got < n_ref, (n_got, n_ref)
    assert set(ref) == set(got)
    for n in ref:
        assert P.relerr(got[n], ref[n]) <= 2e-5, (n, P.relerr(got[n], ref[n]))
