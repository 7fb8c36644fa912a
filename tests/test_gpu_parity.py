"""GPU suite (-m gpu): the shipped sm_100a library, called through the C-ABI, against the oracle on
the same seeded inputs; at BASELINE.json's full sizes through direct oracle comparison on a reduced
channel count plus size-independent properties (delta kernel, shift, linearity, causality)."""
import ctypes
import os

import numpy as np
import pytest
import torch

import parity_cases as P

pytestmark = pytest.mark.gpu
DEV = "cuda"
T = lambda a: torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module", autouse=True)
def real_library():
    from dna_b200 import _lib
    assert not _lib.is_emulation()
    lib = _lib.lib()
    assert b"sm_100a" in lib.hy_version()
    yield lib


@pytest.mark.parametrize("mode", ["plain", "gated", "shortconv"])
@pytest.mark.parametrize("shape", [(2, 3, 100), (3, 2, 257), (1, 2, 1000), (9, 1, 300), (2, 4, 1024), (1, 3, 4096), (2, 1, 2047), (1, 2, 1)])
def test_fused_regime_fp32(mode, shape):
    for name, e in P.conv_case(*shape, mode=mode, device=DEV).items():
        assert e <= P.FP32_TOL, (mode, shape, name, e)


@pytest.mark.parametrize("shape,mode", [((1, 2, 5000), "plain"), ((2, 2, 8192), "shortconv"), ((1, 2, 20000), "shortconv"),
                                        ((2, 1, 4097), "gated"), ((1, 3, 32768), "shortconv"), ((1, 1, 70000), "shortconv"),
                                        ((1, 2, 160000), "shortconv"), ((1, 1, 262144), "plain"), ((1, 1, 300000), "shortconv")])
def test_four_step_regime_fp32(shape, mode):
    """M1 = 2 .. 128 with the production 4096-point rows (C2: L=32768; C3: L=160000)."""
    for name, e in P.conv_case(*shape, mode=mode, device=DEV).items():
        assert e <= 5e-5, (mode, shape, name, e)


@pytest.mark.parametrize("shape,mode", [((1, 2, 5000), "plain"), ((2, 2, 8192), "shortconv"), ((2, 1, 4097), "gated"),
                                        ((1, 3, 32768), "shortconv"), ((1, 2, 160000), "shortconv"), ((1, 1, 1_000_000), "shortconv")])
def test_four_step_saved_spectrum_backward(shape, mode):
    """Forward keeps the spectrum of g, backward transforms dy only (hy_conv_*_args.gsave); dD = dk[:, 0]."""
    for name, e in P.conv_case(*shape, mode=mode, device=DEV, gsave=True).items():
        assert e <= 5e-5, (mode, shape, name, e)


@pytest.mark.parametrize("L,M1", [(40000, 10), (45000, 12), (70000, 20), (90000, 24), (160000, 40), (190000, 48), (300000, 80),
                                  (390000, 96), (600000, 160), (780000, 192), (1_200_000, 320), (1_500_000, 384),
                                  (100000, 32), (250000, 64), (500000, 128)])
def test_four_step_transform_lengths(real_library, L, M1):
    """Every column length with an instance at the production row length: the 5 * 2^a and 3 * 2^a families (the reference
    transforms exactly 2L points, hyena.py:61-62; C3's L = 160 000 runs 40 x 4096 points, not 2^18) and the power-of-two
    lengths the older cases no longer reach; recompute and saved-spectrum backward."""
    real_library.hy_fft_len.restype = ctypes.c_int
    assert real_library.hy_fft_len(L) == M1 * 4096
    for gsave in (False, True):
        for name, e in P.conv_case(1, 1, L, mode="shortconv", device=DEV, seed=L % 1000, gsave=gsave).items():
            assert e <= 1e-4, (L, gsave, name, e)


def test_result_independent_of_transform_length(real_library):
    """The linear convolution does not depend on the transform length: the 40 x 4096 and the 2^18 transforms of the same
    L = 160 000 rows agree to fp32 rounding."""
    from dna_b200 import kernels as K
    g = torch.Generator().manual_seed(5)
    B, H, L = 1, 4, 160000
    uT = torch.randn(B, 3 * H, L, generator=g).to(DEV)
    k = (torch.randn(H, L, generator=g) * torch.exp(-torch.arange(L) / 2000.0)).to(DEV)
    Dp = torch.randn(H, generator=g).to(DEV)
    sw = torch.randn(3 * H, 3, generator=g).to(DEV)
    sb = torch.randn(3 * H, generator=g).to(DEV)
    real_library.hy_debug_set_odd_lengths.restype = ctypes.c_int
    outs = []
    for on in (1, 0):
        real_library.hy_debug_set_odd_lengths(on)
        try:
            Kf = K.filter_spectrum(k, Dp, L)
            outs.append(K.conv_fwd(uT, Kf, L, in_mode=K.IN_SHORTCONV, out_mode=K.OUT_SHORTCONV, sw=sw, sb=sb)[0][..., :L].float())
        finally:
            real_library.hy_debug_set_odd_lengths(1)
    scale = outs[1].abs().max().item()
    assert (outs[0] - outs[1]).abs().max().item() <= 2e-5 * scale


def test_saved_spectrum_bf16_matches_recompute():
    a = P.conv_case(1, 2, 40000, mode="shortconv", device=DEV, dtype=torch.bfloat16)
    b = P.conv_case(1, 2, 40000, mode="shortconv", device=DEV, dtype=torch.bfloat16, gsave=True)
    for name in a:
        assert b[name] <= 6e-2 and abs(a[name] - b[name]) <= 2e-3, (name, a[name], b[name])


@pytest.mark.parametrize("L", [1_000_000, 1_048_576])
def test_one_million_direct_parity(L):
    """C4 sequence length, 2 channels, straight against the CPU oracle (M1 = 256)."""
    for name, e in P.conv_case(1, 2, L, mode="shortconv", device=DEV, seed=7).items():
        assert e <= 1e-4, (L, name, e)        # dsb / dD sum 1e6 fp32 terms on both sides


def test_two_million_supported():
    for name, e in P.conv_case(1, 1, 1 << 21, mode="plain", device=DEV, seed=9).items():
        assert e <= 1e-4, (name, e)


def test_one_million_properties():
    """Size-independent properties at the full C4 row length, all 256 channels' worth of rows kept small
    by using H = 8: delta filter = identity (+D), shifted delta = delay, linearity, causality."""
    from dna_b200 import kernels as K
    L, H, B = 1_000_000, 8, 1
    g = torch.Generator().manual_seed(3)
    u = torch.randn(B, H, L, generator=g).to(DEV)
    D = torch.randn(H, generator=g).to(DEV)
    k = torch.zeros(H, L, device=DEV)
    k[:, 0] = 1.0
    out, _ = K.conv_fwd(u, K.filter_spectrum(k, D, L), L)
    ref = u * (1.0 + D)[None, :, None]
    assert (out - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    s = 123_457
    k.zero_()
    k[:, s] = 1.0
    out, _ = K.conv_fwd(u, K.filter_spectrum(k, None, L), L)
    assert out[..., :s].abs().max().item() <= 2e-5
    assert (out[..., s:] - u[..., :L - s]).abs().max().item() <= 2e-5 * u.abs().max().item()
    kk = P.decaying_filter(H, L, g).to(DEV)
    Kf = K.filter_spectrum(kk, D, L)
    u2 = torch.randn(B, H, L, generator=g).to(DEV)
    a, _ = K.conv_fwd(u, Kf, L)
    b, _ = K.conv_fwd(u2, Kf, L)
    c, _ = K.conv_fwd(u + 0.5 * u2, Kf, L)
    assert (c - (a + 0.5 * b)).abs().max().item() <= 5e-5 * c.abs().max().item()
    u3 = u.clone()
    u3[..., 600_000:] = 0
    d, _ = K.conv_fwd(u3, Kf, L)
    # causality: zeroing the future must not change the past (beyond FFT rounding)
    assert (d[..., :600_000] - a[..., :600_000]).abs().max().item() <= 5e-5 * a.abs().max().item()


@pytest.mark.parametrize("shape", [(2, 3, 100), (1, 2, 1000), (2, 2, 3000), (1, 2, 32768), (1, 1, 200000)])
def test_bf16_forward_not_worse_than_reference_bf16(shape):
    e_ours, e_ref, scale = P.bf16_forward_case(*shape, device=DEV)
    assert e_ours <= 2 * e_ref + scale * 2 ** -8, (e_ours, e_ref, scale)


def test_bf16_four_step_backward_close_to_reference_bf16():
    errs = P.conv_case(1, 2, 20000, mode="shortconv", device=DEV, dtype=torch.bfloat16)
    for name, e in errs.items():
        assert e <= 6e-2, (name, e)


@pytest.mark.parametrize("mode", ["plain", "gated", "shortconv"])
def test_bf16_backward_close_to_reference_bf16(mode):
    errs = P.conv_case(2, 2, 3000, mode=mode, device=DEV, dtype=torch.bfloat16)
    for name, e in errs.items():
        assert e <= 6e-2, (mode, name, e)
    assert errs["out"] <= 2e-2


@pytest.mark.parametrize("cfg", [(16, 64, 5, 2, 100, 130), (256, 64, 5, 2, 300, 400), (8, 16, 3, 2, 33, 40), (70, 64, 5, 1, 64, 64),
                                 (256, 64, 5, 2, 32768, 32770), (128, 64, 5, 2, 1_000_000, 1_000_002), (600, 64, 5, 2, 5000, 5000),
                                 (5, 32, 7, 3, 200, 200), (4, 16, 9, 1, 50, 50)])
def test_filter_kernel(cfg):
    e_ours, e_ref32 = P.filter_case(*cfg, device=DEV)
    assert e_ours <= 4 * e_ref32 + 1e-6, (cfg, e_ours, e_ref32)


@pytest.mark.parametrize("cfg", [(256, 64, True, False), (256, 1000, True, False), (128, 4097, True, True), (32, 130, False, False),
                                 (256, 300_001, True, True), (64, 1, True, False)])
def test_filter_out_bwd_tensor_core_kernel(cfg):
    """3xTF32 keeps fp32-class accuracy: no worse than 4x the fp32 torch GEMMs against fp64 (+ a floor)."""
    from dna_b200 import kernels as K
    assert K.filter_out_bwd_supported(cfg[0], 64) and not K.filter_out_bwd_supported(cfg[0], 32)
    for e_ours, e_ref32 in P.filter_out_bwd_case(cfg[0], cfg[1], DEV, modulate=cfg[2], ragged=cfg[3]):
        assert e_ours <= 4 * e_ref32 + 1e-6, (cfg, e_ours, e_ref32)


@pytest.mark.parametrize("cfg", [(64, 5, 2, 300, 320), (64, 5, 2, 64, 64), (16, 3, 1, 130, 130), (32, 7, 0, 100, 128),
                                 (64, 5, 2, 40_001, 40_002), (64, 5, 2, 1_000_000, 1_000_002)])
def test_filter_saved_trunk_backward_matches_recompute(cfg):
    # the saved pre-activations come from the tensor-core forward (3xTF32, hy_filter_tc05.cu), the recompute runs on
    # FFMA: the two agree to the 3xTF32 level times the sin(10 x) gain; each is gated against the fp64 oracle in
    # test_filter_backward_at_production_length_vs_oracle
    assert P.filter_trunk_saved_case(*cfg, device=DEV) <= 5e-5


@pytest.mark.parametrize("cfg", [(3, 50, 40, 1), (2, 20, 32, 0), (4, 100, 64, 3), (2, 37, 37, 5), (3, 64, 65, 9),
                                 (2, 1_000_000, 1_000_001, 1), (1, 1_048_576, 1_000_001, 1), (2, 5000, 8192, 3), (5, 70, 41, 3),
                                 (6, 3000, 2049, 1), (5, 300_000, 300_001, 1)])
def test_tokenizer_bit_exact(cfg):
    assert P.tokenizer_case(*cfg, device=DEV)


def test_operator_and_model_match_reference_outputs(golden_dir):
    """The reference's own outputs/gradients (tests/golden, produced from /root/reference) with the
    reference's state_dict loaded strictly into our modules on the GPU."""
    from test_emu_module import CFGS, operator_vs_golden
    g = np.load(os.path.join(golden_dir, "operator.npz"))
    for tag in CFGS:
        errs, _ = operator_vs_golden(tag, g, DEV)
        for name, e in errs.items():
            assert e <= 5e-5, (tag, name, e)
    from dna_b200.standalone import HyenaDNAModel
    g = np.load(os.path.join(golden_dir, "model_tiny.npz"))
    model = HyenaDNAModel(d_model=32, n_layer=2, d_inner=128, vocab_size=12, embed_dropout=0.0,
                          layer=dict(l_max=258, emb_dim=5, filter_order=64, short_filter_order=3, modulate=True, w=10,
                                     lr=6e-4, wd=0.0, lr_pos_emb=0.0))
    model.load_state_dict({k[3:]: T(g[k]) for k in g.files if k.startswith("sd/")}, strict=True)
    model = model.to(DEV).eval()
    h = model(T(g["ids"]).to(DEV))
    assert P.relerr(h, T(g["hidden"])) <= 1e-5
    h.float().pow(2).mean().backward()
    assert P.relerr(model.backbone.layers[0].mixer.in_proj.weight.grad, T(g["grad/backbone.layers.0.mixer.in_proj.weight"])) <= 5e-5


def test_operator_c1_config_vs_oracle_fp32_and_autocast():
    """BASELINE C1 shape (d_model=128, L=1024, B=8): fp32 vs the oracle, and bf16 autocast vs the
    oracle's fp32 answer within the bf16 budget (pin bf16 against fp32 truth)."""
    from dna_b200.hyena import HyenaOperator
    from oracle import hyena_oracle as O
    torch.manual_seed(0)
    D, L, B = 128, 1024, 8
    op = HyenaOperator(d_model=D, l_max=L + 2, emb_dim=5, filter_order=64, w=10, lr_pos_emb=0.0, shift=0.05)
    sd = {k: v.clone() for k, v in op.state_dict().items()}
    u = torch.randn(B, L, D)
    ref = O.hyena_operator(u, sd, l_max=L + 2, shift=0.05)
    op = op.to(DEV)
    y = op(u.to(DEV))
    assert P.relerr(y, ref) <= 5e-5
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yb = op(u.to(DEV))
    assert yb.dtype == torch.bfloat16
    assert P.relerr(yb.float(), ref) <= 3e-2


def test_no_silent_fallback_on_cpu_tensors():
    from dna_b200 import _lib
    from dna_b200.fftconv import fftconv_func
    with pytest.raises(_lib.HyenaB200Error):
        fftconv_func(torch.randn(1, 2, 64), torch.randn(2, 64), torch.randn(2), None, False)


def test_dp_allreduce_single_rank_noop_and_launch_counter():
    from dna_b200 import kernels as K
    n0 = K.launch_count()
    P.conv_case(1, 1, 300, mode="plain", device=DEV)
    assert K.launch_count() > n0


@pytest.mark.parametrize("D,rows", [(128, 1000), (256, 40_001), (512, 777), (1024, 300)])
def test_add_layer_norm_vs_oracle(D, rows):
    """Block glue (hy_addln.cu) against the oracle's restatement of standalone_hyenadna.py:521-525: fp32 values and
    gradients, and the autocast pattern (bf16 hidden + fp32 stream -> bf16 normed)."""
    from dna_b200 import block_ops
    from oracle.hyena_model_oracle import block_add_norm
    g = torch.Generator().manual_seed(D)
    x = torch.randn(rows, D, generator=g)
    r = torch.randn(rows, D, generator=g) * 2
    norm = torch.nn.LayerNorm(D).to(DEV)
    with torch.no_grad():
        norm.weight.copy_(torch.randn(D, generator=g)); norm.bias.copy_(torch.randn(D, generator=g))
    gy, gr = torch.randn(rows, D, generator=g), torch.randn(rows, D, generator=g)
    # fp32
    xd, rd = x.to(DEV).requires_grad_(True), r.to(DEV).requires_grad_(True)
    y, ro = block_ops.add_layer_norm(xd, rd, norm)
    ((y * gy.to(DEV)).sum() + (ro * gr.to(DEV)).sum()).backward()
    xo, ro_in = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
    wo, bo = norm.weight.detach().cpu().requires_grad_(True), norm.bias.detach().cpu().requires_grad_(True)
    y_ref, r_ref = block_add_norm(xo, ro_in, wo, bo, norm.eps)
    ((y_ref * gy).sum() + (r_ref * gr).sum()).backward()
    assert torch.equal(ro.detach().cpu(), r_ref.detach())
    for name, a, b in [("y", y, y_ref), ("dx", xd.grad, xo.grad), ("dres", rd.grad, ro_in.grad),
                       ("dgamma", norm.weight.grad, wo.grad), ("dbeta", norm.bias.grad, bo.grad)]:
        assert P.relerr(a.detach().cpu(), b.detach()) <= 1e-5, (name, D, rows)
    # autocast pattern
    xb = x.to(torch.bfloat16)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yb, rb = block_ops.add_layer_norm(xb.to(DEV), r.to(DEV), norm)
    y_ref, r_ref = block_add_norm(xb, r, wo.detach(), bo.detach(), norm.eps)
    assert yb.dtype == torch.bfloat16 and rb.dtype == torch.float32 and torch.equal(rb.cpu(), r_ref)
    assert (yb.float().cpu() - y_ref.to(torch.bfloat16).float()).abs().max() <= 2 ** -7 * y_ref.abs().max()


@pytest.mark.parametrize("shape,dtype,gsave", [((2, 2, 300), torch.float32, False), ((1, 2, 1001), torch.float32, False),
                                               ((2, 1, 5000), torch.float32, True), ((1, 2, 160000), torch.bfloat16, True),
                                               ((1, 1, 1_000_000), torch.float32, True)])
def test_deferred_dx0_in_short_filter_backward(shape, dtype, gsave):
    """hy_conv_bwd_args.defer_dx0 + hy_shortconv_bwd_gate against the oracle, and equal to the two-step form."""
    a = P.conv_case(*shape, mode="shortconv", device=DEV, dtype=dtype, gsave=gsave)
    b = P.conv_case(*shape, mode="shortconv", device=DEV, dtype=dtype, gsave=gsave, defer=True)
    tol = 5e-5 if dtype == torch.float32 else 6e-2
    for name in a:
        assert b[name] <= tol and abs(a[name] - b[name]) <= (1e-6 if dtype == torch.float32 else 2e-3), (name, a[name], b[name])


@pytest.mark.parametrize("cfg", [(3, 50, True), (1, 1, False), (4, 4097, True), (2, 1_000_000, True)])
def test_reverse_complement_bit_exact(golden_dir, cfg):
    B, maxchars, with_apply = cfg
    g = np.load(os.path.join(golden_dir, "revcomp.npz")) if B == 3 else None
    assert P.revcomp_case(B, maxchars, DEV, seed=maxchars, with_apply=with_apply, golden=g)
    # round trip at full length: rc(rc(x)) == x
    x = torch.from_numpy(np.frombuffer(b"ACGTNacgtn.", dtype=np.uint8)[np.random.default_rng(1).integers(0, 11, size=(2, 100_003))].copy()).to(DEV)
    from dna_b200 import kernels as K
    assert torch.equal(K.reverse_complement(K.reverse_complement(x)), x)


@pytest.mark.parametrize("cfg", [(8, 300, 2, torch.float32), (8, 5000, 4, torch.float32), (16, 70000, 8, torch.bfloat16),
                                 (256, 1_000_000, 8, torch.bfloat16)])
def test_channel_slabs_equal_the_full_operator(cfg):
    """Channel partition of a single sequence over the ranks (BASELINE configs[3]): collective-free and bit-identical."""
    Dm, L, world, dt = cfg
    assert P.channel_slab_case(Dm, L, world, DEV, dtype=dt)


def test_group_pipeline_and_scratch_budget_do_not_change_results():
    """hy_set_pipeline (row groups on internal streams) and hy_set_scratch_budget (rows per group) are scheduling knobs:
    the result must be bit-identical for any setting, and the call must stay ordered on the caller's stream."""
    import ctypes
    from dna_b200 import _lib, kernels as K
    lib = _lib.lib()
    g = torch.Generator().manual_seed(11)
    H, L = 24, 40000
    u = torch.randn(1, H, L, generator=g).to(DEV)
    k = P.decaying_filter(H, L, g).to(DEV)
    D = torch.randn(H, generator=g).to(DEV)
    w = torch.randn(1, H, L, generator=g).to(DEV)

    def run():
        Kf = K.filter_spectrum(k, D, L)
        out, _ = K.conv_fwd(u, Kf, L)
        du, _, _, dKacc, dD = K.conv_bwd(w, u, Kf, L)
        return out.clone(), du.clone(), K.conv_dk(dKacc, L).clone(), dD.clone()

    base = run()
    try:
        for nstream, budget in ((2, 8 << 20), (4, 3 << 20), (1, 1 << 20)):
            assert lib.hy_set_pipeline(nstream, ctypes.c_size_t(budget)) == 0
            for a, b in zip(base, run()):
                assert torch.equal(a, b), (nstream, budget)
    finally:
        lib.hy_set_pipeline(1, ctypes.c_size_t(0))
        lib.hy_set_scratch_budget(ctypes.c_size_t(0))


def test_clock_probe_reports_a_plausible_sm_clock():
    from dna_b200 import kernels as K
    out = torch.zeros(2, dtype=torch.int64, device=DEV)
    K.clock_probe(out)
    torch.cuda.synchronize()
    cycles, ns = out.tolist()
    assert ns > 10_000 and 500 <= cycles * 1000 / ns <= 3000      # MHz


# ---- round 2: the configurations VERDICT r01 listed as untested -------------------------------------------------------
def test_bf16_headline_length_forward_and_backward_vs_oracle():
    """The bench's own dtype x length: bf16 activations at L = 1 000 000 (saved-spectrum backward), values and every
    gradient.  Ours and the oracle's own bf16 path are both measured against fp64 truth from the same bf16 inputs."""
    res = P.bf16_conv_truth_case(1, 2, 1_000_000, DEV, seed=5, gsave=True)
    for name, (e_ours, e_ref, scale) in res.items():
        assert e_ours <= 2 * e_ref + scale * 2 ** -8, (name, e_ours, e_ref, scale)


@pytest.mark.parametrize("shape,gsave", [((2, 2, 3000), False), ((1, 2, 20000), True), ((1, 2, 160000), True)])
def test_bf16_backward_not_worse_than_reference_bf16(shape, gsave):
    """Same budget form as the forward test for every gradient (replaces a flat 6e-2 gate as the binding check)."""
    res = P.bf16_conv_truth_case(*shape, DEV, seed=1, gsave=gsave)
    for name, (e_ours, e_ref, scale) in res.items():
        assert e_ours <= 2 * e_ref + scale * 2 ** -8, (shape, name, e_ours, e_ref, scale)


@pytest.mark.parametrize("D,L,init", [(256, 32768, "default"), (256, 32768, "backbone"), (256, 1_000_000, "backbone"),
                                      (128, 300_001, "default"), (64, 4097, "default")])
def test_filter_backward_at_production_length_vs_oracle(D, L, init):
    """Every MLP parameter gradient (hy_filter_out_bwd at D = 256 / 128 / 64, hy_filter_trunk_bwd_saved) against the
    oracle's fp64 autograd of hyena.py:203-242; gate of the same form as test_filter_kernel: no worse than 4x the
    oracle's own fp32 error (+ a floor for sums over 1e6 positions)."""
    for name, (e_ours, e_ref32) in P.filter_bwd_case(D, L, DEV, init=init).items():
        assert e_ours <= 4 * e_ref32 + 2e-6, (D, L, init, name, e_ours, e_ref32)


@pytest.mark.parametrize("tag", ["o3_src", "o3_sa", "o4_src"])
def test_order_gt2_matches_reference_on_gpu(golden_dir, tag):
    from test_emu_module import operator_vs_golden
    g = np.load(os.path.join(golden_dir, "features.npz"))
    errs, _ = operator_vs_golden(tag, g, DEV)
    for name, e in errs.items():
        assert e <= 5e-5, (tag, name, e)


def test_order3_four_step_length_vs_oracle():
    """order = 3 at a four-step length (the PREGATE / POSTGATE kernels with saved spectra) against O.hyena_operator."""
    from dna_b200.hyena import HyenaOperator
    from oracle import hyena_oracle as O
    torch.manual_seed(1)
    D, L, B = 16, 9000, 2
    op = HyenaOperator(d_model=D, l_max=L, order=3, emb_dim=5, filter_order=64, w=10, lr_pos_emb=0.0, shift=0.05)
    sd = {k: v.detach().clone() for k, v in op.state_dict().items()}
    u = torch.randn(B, L, D)
    w = torch.randn(B, L, D)
    psd = {k: v.requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    ur = u.clone().requires_grad_(True)
    (O.hyena_operator(ur, psd, l_max=L, shift=0.05, order=3) * w).sum().backward()
    op = op.to(DEV)
    ud = u.to(DEV).requires_grad_(True)
    y = op(ud)
    (y * w.to(DEV)).sum().backward()
    assert P.relerr(ud.grad, ur.grad) <= 5e-5
    for n, p_ in op.named_parameters():
        assert P.relerr(p_.grad, psd[n].grad) <= 1e-4, n


def test_inference_filter_cache_on_gpu():
    from test_emu_module import inference_filter_cache_case
    inference_filter_cache_case(DEV)


@pytest.mark.parametrize("cfg", [(2, 128, 1024, 2, False), (4, 256, 32768, 1, False), (4, 256, 32768, 1, True),
                                 (2, 64, 40000, 1, False), (2, 64, 160000, 1, True)])   # the last two: 10 x 4096 and 40 x 4096 point transforms
def test_model_level_step_vs_oracle(cfg):
    """BASELINE C1- and C2-shaped models (C2: 4 layers / d_model 256 / 32 k): loss and every parameter gradient of one
    training step against oracle.hyena_model_oracle.lm_loss.  fp32 tight; bf16 autocast against the oracle's fp32
    answer within a bf16 budget."""
    n_layer, d_model, L, B, autocast = cfg
    loss, loss_ref, errs = P.model_step_case(n_layer, d_model, L, B, DEV, autocast=autocast)
    if autocast:
        assert abs(loss - loss_ref) <= 2e-2 * abs(loss_ref), (loss, loss_ref)
        bad = {n: e for n, e in errs.items() if e > 0.15}
    else:
        assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref), (loss, loss_ref)
        bad = {n: e for n, e in errs.items() if e > 5e-4}
    assert not bad, bad


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_operator_on_non_current_device():
    """tensors on cuda:1 while the current device is cuda:0 — the wrappers must switch device and stream."""
    from dna_b200.hyena import HyenaOperator
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    op = HyenaOperator(d_model=32, l_max=5002, emb_dim=5, filter_order=64, w=10, lr_pos_emb=0.0, shift=0.05)
    u = torch.randn(2, 5000, 32)
    op0 = op.to("cuda:0")
    y0 = op0(u.to("cuda:0"))
    y0.sum().backward()
    g0 = op0.in_proj.weight.grad.clone()
    op.zero_grad()
    op1 = op.to("cuda:1")
    assert torch.cuda.current_device() == 0
    y1 = op1(u.to("cuda:1"))
    y1.sum().backward()
    assert y1.device.index == 1 and torch.equal(y1.cpu(), y0.cpu())
    assert torch.equal(op1.in_proj.weight.grad.cpu(), g0.cpu())
    with pytest.raises(Exception):
        from dna_b200 import kernels as K
        K.conv_fwd(torch.randn(1, 2, 300, device="cuda:0"), torch.randn(2, 512, 2, device="cuda:1"), 300)


@pytest.mark.parametrize("cfg", [((1, 6, 1_000_000), torch.bfloat16), ((2, 5, 250_000), torch.float32)])
def test_persistent_pipeline_equals_per_phase_launches_gpu(cfg):
    """hy_conv_pipe.cuh (opt-in, HYENA_B200_CONV_PIPE=1): the persistent A / B / C pipeline over a ring of row buffers
    must give the same bits as the default per-phase launches — also with more rows than ring buffers at L = 1 M."""
    from dna_b200 import _lib
    lib = _lib.lib()
    shape, dt = cfg
    res = []
    try:
        for pipe in (1, 0):
            lib.hy_debug_set_conv_pipe(pipe)
            res.append(P.conv_case(*shape, mode="shortconv", device=DEV, dtype=dt, gsave=True, seed=5))
    finally:
        lib.hy_debug_set_conv_pipe(0)
    assert res[0] == res[1], (res[0], res[1])



# ---- keyword surface of fftconv_func, operator options, LongConv — GPU twins of test_emu_module.py ---
@pytest.mark.parametrize("tag", ["bi_a", "bi_b", "bi_c"])
@pytest.mark.parametrize("variant", ["bidir", "krev", "bidir_krev"])
def test_fftconv_func_krev_bidirectional_gpu(golden_dir, tag, variant):
    g = np.load(os.path.join(golden_dir, "features.npz"))
    for name, e in P.fftconv_variant_case(g, tag, variant, DEV).items():
        assert e <= 5e-5, (tag, variant, name, e)


@pytest.mark.parametrize("hd", [2, 8])
def test_fftconv_func_h3_heads_gpu(golden_dir, hd):
    g = np.load(os.path.join(golden_dir, "features.npz"))
    for name, e in P.h3_heads_case(g, hd, DEV).items():
        assert e <= 5e-5, (hd, name, e)


@pytest.mark.parametrize("tag", [t for t in P.OPTION_KW if t not in ("bidir_src", "drop")])
def test_operator_options_match_reference_gpu(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "options.npz"))
    for name, e in P.operator_option_case(g, tag, DEV).items():
        assert e <= 5e-5, (tag, name, e)


def test_bidirectional_operator_matches_reference_gpu(golden_dir):
    g = np.load(os.path.join(golden_dir, "features.npz"))
    for name, e in P.operator_option_case(g, "bidir_src", DEV).items():
        assert e <= 5e-5, (name, e)


@pytest.mark.parametrize("tag", list(P.LONGCONV_KW))
def test_long_conv_matches_reference_gpu(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "options.npz"))
    for name, e in P.long_conv_case(g, tag, DEV).items():
        assert e <= 5e-5, (tag, name, e)


def test_operator_dropout_on_gpu():
    """dropout > 0 (hyena.py:481): eval mode equals p = 0; training mode zeroes ~p of the gated products and rescales."""
    from dna_b200.hyena import HyenaOperator
    torch.manual_seed(0)
    op = HyenaOperator(d_model=16, l_max=512, dropout=0.5, emb_dim=5, filter_order=16, lr_pos_emb=0.0).to(DEV)
    op0 = HyenaOperator(d_model=16, l_max=512, dropout=0.0, emb_dim=5, filter_order=16, lr_pos_emb=0.0).to(DEV)
    op0.load_state_dict(op.state_dict())
    u = torch.randn(2, 512, 16, device=DEV)
    op.eval(), op0.eval()
    assert P.relerr(op(u), op0(u)) <= 1e-6
    op.train()
    y1, y2 = op(u), op(u)
    assert not torch.equal(y1, y2) and torch.isfinite(y1).all()


def test_bidirectional_long_sequence_vs_oracle():
    """bidirectional long convolution at a four-step length against the oracle (hyena.py:68-74)."""
    from dna_b200.fftconv import fftconv_func
    from oracle import hyena_oracle as O
    gen = torch.Generator().manual_seed(2)
    B, H, L = 2, 3, 20_001
    u = torch.randn(B, H, L, generator=gen).requires_grad_(True)
    k = P.decaying_filter(H, L, gen).requires_grad_(True)
    kr = P.decaying_filter(H, L, gen).requires_grad_(True)
    D = torch.randn(H, generator=gen).requires_grad_(True)
    w = torch.randn(B, H, L, generator=gen)
    ref = O.fftconv_ref(u, k, D, None, gelu=False, k_rev=kr, bidirectional=True)
    gref = torch.autograd.grad((ref * w).sum(), [u, k, kr, D])
    ud, kd, krd, Dd = (t.detach().to(DEV).requires_grad_(True) for t in (u, k, kr, D))
    out = fftconv_func(ud, kd, Dd, None, False, k_rev=krd, bidirectional=True)
    gout = torch.autograd.grad((out * w.to(DEV)).sum(), [ud, kd, krd, Dd])
    assert P.relerr(out, ref) <= P.FP32_TOL
    for name, a, b in zip(["du", "dk", "dkrev", "dD"], gout, gref):
        assert P.relerr(a, b) <= 5e-5, (name, P.relerr(a, b))


def test_fetch_intervals_bit_exact_gpu(golden_dir):
    assert P.fetch_intervals_case(np.load(os.path.join(golden_dir, "ingest.npz")), DEV) > 20


@pytest.mark.parametrize("i", [0, 1, 2])
def test_bert_mask_bit_exact_gpu(golden_dir, i):
    assert P.bert_mask_case(np.load(os.path.join(golden_dir, "ingest.npz")), i, DEV)


def test_device_ingest_pipeline_gpu():
    """DeviceFastaInterval -> tokenizer -> bert_mask_cuda end to end on the device: shapes, padding, masking rate."""
    from dna_b200.tokenizer import CharacterTokenizer, DeviceFastaInterval, bert_mask_cuda
    rng = np.random.default_rng(1)
    chrom = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=200_000)])
    fi = DeviceFastaInterval({"chr1": chrom}, rc_aug=True, pad_interval=True, shift_augs=(-3, 3), device=torch.device(DEV))
    starts = torch.arange(0, 190_000, 10_000)
    out, lens = fi("chr1", starts, starts + 8192, 8192, generator=torch.Generator().manual_seed(0))
    assert out.shape == (19, 8192) and (lens == 8192).all()
    tok = CharacterTokenizer(["A", "C", "G", "T", "N"], model_max_length=8193)
    ids = tok.encode_bytes_cuda(out, lens, 8193, add_special_tokens=True)
    seq, mask, labels = bert_mask_cuda(ids, 3, 4, 12, special_token_ids=list(range(7)))
    rate = mask.float().mean().item()
    assert 0.13 < rate < 0.17 and (labels[~mask] == -100).all() and (seq[~mask] == ids[~mask]).all()
    assert ((seq[mask] == 3).float().mean().item()) > 0.7 and (seq[mask] >= 3).all()


def test_filter_reuse_across_micro_batches_gpu():
    """filter_reuse at a four-step length on the GPU: two micro-batches, one filter generation + spectrum + backward."""
    from dna_b200 import kernels as K
    from dna_b200.hyena import HyenaOperator, flush_filter_grads
    g = torch.Generator().manual_seed(6)
    D, L = 32, 9000
    us = [torch.randn(1, L, D, generator=g).to(DEV) for _ in range(2)]
    ws = [torch.randn(1, L, D, generator=g).to(DEV) for _ in range(2)]

    def run(reuse):
        torch.manual_seed(0)
        op = HyenaOperator(d_model=D, l_max=L, emb_dim=5, filter_order=64, w=10, lr_pos_emb=0.0, shift=0.05).to(DEV)
        op.filter_reuse = reuse
        n0 = K.launch_count()
        for u, w in zip(us, ws):
            (op(u) * w).sum().backward()
        flush_filter_grads(op)
        return {n: p.grad.clone() for n, p in op.named_parameters() if p.grad is not None}, K.launch_count() - n0

    ref, n_ref = run(False)
    got, n_got = run(True)
    assert n_got < n_ref and set(ref) == set(got)
    for n in ref:
        assert P.relerr(got[n], ref[n]) <= 5e-5, (n, P.relerr(got[n], ref[n]))


@pytest.mark.parametrize("cfg", [(torch.float32, torch.float32, True), (torch.bfloat16, torch.float32, True),
                                 (torch.bfloat16, torch.bfloat16, False)])
def test_add_ln_with_dropout_mask_gpu(cfg):
    xdt, rdt, with_res = cfg
    errs = P.add_ln_dropout_case(DEV, xdt, rdt, with_res, rows=4099)
    tol = 5e-6 if xdt == torch.float32 else 2e-2
    assert errs.pop("stream_equal") == 0.0
    for n, e in errs.items():
        assert e <= tol, (cfg, n, e)


@pytest.mark.parametrize("cfg", [((3, 3, 1700), "shortconv", torch.float32), ((2, 5, 2048), "plain", torch.float32),
                                 ((4, 16, 4096), "gated", torch.float32), ((8, 32, 1100), "shortconv", torch.bfloat16),
                                 ((2, 6, 4095), "shortconv", torch.bfloat16)])
def test_single_kernel_regime_backward_with_saved_spectrum_gpu(cfg):
    shape, mode, dt = cfg
    for name, e in P.conv_case(*shape, mode=mode, device=DEV, dtype=dt, gsave=True, seed=4).items():
        assert e <= (5e-5 if dt == torch.float32 else 6e-2), (cfg, name, e)
