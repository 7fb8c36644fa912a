"""Pins oracle/hyena_oracle.py against outputs of the reference itself (tests/golden/*.npz, made by
tests/golden/make_golden.py from /root/reference).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import hyena_oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def g_fft(golden_dir):
    return np.load(os.path.join(golden_dir, "fftconv.npz"))


@pytest.fixture(scope="module")
def g_op(golden_dir):
    return np.load(os.path.join(golden_dir, "operator.npz"))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_fftconv_ref_matches_reference(g_fft, tag):
    u, k, D = T(g_fft[f"{tag}_u"]), T(g_fft[f"{tag}_k"]), T(g_fft[f"{tag}_D"])
    # bit-exact: same torch primitives in the same order (hyena.py:60-92; standalone :45-60)
    assert torch.equal(O.fftconv_ref(u, k, D, None, gelu=False), T(g_fft[f"{tag}_hy_nogelu"]))
    assert torch.equal(O.fftconv_ref(u, k, D, None, gelu=True), T(g_fft[f"{tag}_hy_gelu"]))
    assert torch.equal(O.fftconv_ref(u, k, D, None, gelu=False), T(g_fft[f"{tag}_standalone"]))
    out_bf = O.fftconv_ref(u.to(torch.bfloat16), k, D, None, gelu=False)
    assert out_bf.dtype == torch.bfloat16
    assert torch.equal(out_bf.float(), T(g_fft[f"{tag}_hy_bf16"]))
    out5 = O.fftconv_ref(u[:, None, :, None, :], k, D[None, :, None], None, gelu=False)
    assert torch.equal(out5, T(g_fft[f"{tag}_hy_5d"]))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_fftconv_ref_gradients_match_reference(g_fft, tag):
    u = T(g_fft[f"{tag}_u"]).requires_grad_(True)
    k = T(g_fft[f"{tag}_k"]).requires_grad_(True)
    D = T(g_fft[f"{tag}_D"]).requires_grad_(True)
    (O.fftconv_ref(u, k, D, None, gelu=False) * T(g_fft[f"{tag}_w"])).sum().backward()
    assert torch.equal(u.grad, T(g_fft[f"{tag}_du"]))
    assert torch.equal(k.grad, T(g_fft[f"{tag}_dk"]))
    assert torch.equal(D.grad, T(g_fft[f"{tag}_dD"]))


def test_h3_ref_matches_reference(g_fft):
    k, v, q, ssm, D = (T(g_fft[n]) for n in ("h3_k", "h3_v", "h3_q", "h3_ssm", "h3_D"))
    assert torch.equal(O.fftconv_h3_ref(k, ssm, D, q, v, head_dim=1), T(g_fft["h3_out_hd1"]))
    assert torch.equal(O.fftconv_ref(k, ssm, D, None, gelu=False), T(g_fft["ops_ref_nogelu"]))


def _sd(g_op, tag):
    pre = f"{tag}/sd/"
    return {key[len(pre):]: T(g_op[key]) for key in g_op.files if key.startswith(pre)}


@pytest.mark.parametrize("tag", ["src", "src_e3", "sa", "sa_trunc"])
def test_filter_matches_reference(g_op, tag):
    sd = _sd(g_op, tag)
    fp = {key[len("filter_fn."):]: val for key, val in sd.items() if key.startswith("filter_fn.")}
    ref = T(g_op[f"{tag}/filter"])
    h = O.hyena_filter(fp, ref.shape[1], shift=float(g_op[f"{tag}/shift"]))
    assert torch.equal(h, ref)


def test_positional_tables_match_reference(g_op):
    for tag, emb in (("src", 5), ("src_e3", 3), ("sa", 5)):
        sd = _sd(g_op, tag)
        l_max = int(g_op[f"{tag}/l_max"])
        z, t = O.positional_tables(emb, l_max)
        assert torch.equal(z, sd["filter_fn.pos_emb.z"]) and torch.equal(t, sd["filter_fn.pos_emb.t"])
        D = sd["filter_fn.modulation.deltas"].shape[-1]
        assert torch.equal(O.modulation_deltas(D), sd["filter_fn.modulation.deltas"])


@pytest.mark.parametrize("tag", ["src", "src_e3", "sa", "sa_trunc"])
def test_operator_forward_backward_matches_reference(g_op, tag):
    sd = {key: val.clone().requires_grad_(val.dtype.is_floating_point) for key, val in _sd(g_op, tag).items()}
    u = T(g_op[f"{tag}/u"]).requires_grad_(True)
    y = O.hyena_operator(u, sd, l_max=int(g_op[f"{tag}/l_max"]), shift=float(g_op[f"{tag}/shift"]))
    ref = T(g_op[f"{tag}/y"])
    assert y.shape == ref.shape
    assert torch.allclose(y, ref, rtol=0, atol=1e-6 * ref.abs().max().item())
    (y * T(g_op[f"{tag}/w"])).sum().backward()
    du = T(g_op[f"{tag}/du"])
    assert torch.allclose(u.grad, du, rtol=0, atol=2e-6 * du.abs().max().item())
    pre = f"{tag}/grad/"
    for key in g_op.files:
        if key.startswith(pre):
            name = key[len(pre):]
            gref = T(g_op[key])
            got = sd[name].grad
            assert got is not None, name
            assert torch.allclose(got, gref, rtol=0, atol=5e-6 * max(gref.abs().max().item(), 1e-3)), name


def test_tokenizer_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "tokenizer.npz"))
    for tag, cls in (("src", False), ("sa", True)):
        text = bytes(g[f"{tag}/text"]).decode()
        assert dict(zip(g[f"{tag}/vocab_keys"].tolist(), g[f"{tag}/vocab_vals"].tolist())) == O.VOCAB
        n = len(text) + 1 + int(cls)
        assert O.tokenize_ref(text, n, add_special_tokens=True, cls_token=cls) == g[f"{tag}/with_special"].tolist()
        assert O.tokenize_ref(text, len(text), add_special_tokens=False) == g[f"{tag}/ids"].tolist()
    # docstring example of the reference (hg38_char_tokenizer.py:21-30 id table) + HF 4.28 semantics (restated)
    assert O.tokenize_ref("ACGT", 8) == [4, 4, 4, 7, 8, 9, 10, 1]
    assert O.tokenize_ref("ACGTACGT", 5) == [7, 8, 9, 10, 1]
    assert O.tokenize_ref("", 3) == [4, 4, 1]
    d, t = O.dataset_item_ref("ACGNT", 8, replace_N_token=True)
    assert d.tolist() == [4, 4, 7, 8, 9, 4, 10] and t.tolist() == [4, 7, 8, 9, 4, 10, 1]
    d, _ = O.dataset_item_ref("ACGNTx", 7, nucleotide_encode=True)
    assert d.tolist() == [0, 1, 2, 4, 3, 4]


def test_reverse_complement_oracle_matches_reference(golden_dir):
    """oracle.reverse_complement_ref against outputs of the reference's string_reverse_complement
    (hg38_dataset.py:28-38), tests/golden/revcomp.npz (generated by make_golden.py --only-revcomp)."""
    import os
    import numpy as np
    from oracle import hyena_oracle as O
    g = np.load(os.path.join(golden_dir, "revcomp.npz"))
    n = len([k for k in g.files if k.startswith("in")])
    assert n >= 5
    for i in range(n):
        s = bytes(g[f"in{i}"]).decode()
        assert O.reverse_complement_ref(s) == bytes(g[f"out{i}"]).decode()


# ---- round-2 fixtures (tests/golden/features.npz, make_golden.py --only-features) ------------------------------------
@pytest.fixture(scope="module")
def g_feat(golden_dir):
    return np.load(os.path.join(golden_dir, "features.npz"))


@pytest.mark.parametrize("tag", ["bi_a", "bi_b", "bi_c"])
@pytest.mark.parametrize("variant", ["bidir", "krev", "bidir_krev"])
def test_fftconv_variants_match_reference(g_feat, tag, variant):
    """bidirectional (hyena.py:68-74) and k_rev (:64-66): values bit-exact, autograd gradients bit-exact."""
    u, k, kr, D = (T(g_feat[f"{tag}_{n}"]).requires_grad_(True) for n in ("u", "k", "krev", "D"))
    kw = dict(bidirectional="bidir" in variant, k_rev=kr if "krev" in variant else None)
    y = O.fftconv_ref(u, k, D, None, gelu=False, **kw)
    assert torch.equal(y, T(g_feat[f"{tag}_{variant}_y"]))
    ins = [u, k, D] + ([kr] if "krev" in variant else [])
    for name, g in zip(["du", "dk", "dD", "dkrev"], torch.autograd.grad((y * T(g_feat[f"{tag}_w"])).sum(), ins)):
        assert torch.equal(g, T(g_feat[f"{tag}_{variant}_{name}"])), name


@pytest.mark.parametrize("hd", [2, 8])
def test_h3_multihead_matches_reference(g_feat, hd):
    pre = f"h3_hd{hd}_"
    k, v, q, ssm, D = (T(g_feat[pre + n]).requires_grad_(True) for n in ("k", "v", "q", "ssm", "D"))
    y = O.fftconv_h3_ref(k, ssm, D, q, v, head_dim=hd)
    assert torch.equal(y, T(g_feat[pre + "y"]))
    gr = torch.autograd.grad((y * T(g_feat[pre + "w"])).sum(), [k, ssm, D, q, v])
    for name, g in zip(["dk", "dssm", "dD", "dq", "dv"], gr):
        assert torch.equal(g, T(g_feat[pre + name])), name


@pytest.mark.parametrize("tag,kw", [("o3_src", dict(order=3)), ("o3_sa", dict(order=3, channel_order="standalone")),
                                    ("o4_src", dict(order=4)), ("bidir_src", dict(bidirectional=True))])
def test_operator_variants_match_reference(g_feat, tag, kw):
    """order > 2 recurrence (hyena.py:475-484; standalone :286-288) and the bidirectional operator."""
    pre = f"{tag}/sd/"
    sd = {key[len(pre):]: T(g_feat[key]).clone() for key in g_feat.files if key.startswith(pre)}
    sd = {key: val.requires_grad_(val.dtype.is_floating_point) for key, val in sd.items()}
    u = T(g_feat[f"{tag}/u"]).requires_grad_(True)
    l_max = sd["filter_fn.pos_emb.t"].shape[1]
    y = O.hyena_operator(u, sd, l_max=l_max, shift=float(g_feat[f"{tag}/shift"]), **kw)
    ref = T(g_feat[f"{tag}/y"])
    assert y.shape == ref.shape
    assert torch.allclose(y, ref, rtol=0, atol=1e-6 * ref.abs().max().item())
    (y * T(g_feat[f"{tag}/w"])).sum().backward()
    du = T(g_feat[f"{tag}/du"])
    assert torch.allclose(u.grad, du, rtol=0, atol=2e-6 * du.abs().max().item())
    gp = f"{tag}/grad/"
    for key in g_feat.files:
        if key.startswith(gp):
            gref = T(g_feat[key])
            got = sd[key[len(gp):]].grad
            assert got is not None, key
            assert torch.allclose(got, gref, rtol=0, atol=5e-6 * max(gref.abs().max().item(), 1e-3)), key


# ---- round-2 second batch (tests/golden/options.npz, make_golden.py --only-options) ----------------------------------
OPTION_CASES = {
    "blocks2": dict(num_blocks=2), "blocks4_o3": dict(num_blocks=4, order=3), "outer": dict(outer_mixing=True),
    "ffn_o3": dict(post_order_ffn=True, order=3), "short5": dict(short_filter_order=5),
    "drop": dict(dropout_p=0.25, training=True), "gelu_act": dict(activation=torch.nn.functional.gelu),
}
LONGCONV_CASES = {
    "lc_causal": dict(channels=1, lam=0.001), "lc_bidir": dict(channels=1, bidirectional=True, lam=0.001),
    "lc_ch2_bld": dict(channels=2, transposed=False, lam=0.0005),
    "lc_bidir_short": dict(channels=2, bidirectional=True, lam=0.001, postact=None),
}


@pytest.fixture(scope="module")
def g_opt(golden_dir):
    return np.load(os.path.join(golden_dir, "options.npz"))


def _load_sd(g, tag):
    pre = f"{tag}/sd/"
    sd = {key[len(pre):]: T(g[key]).clone() for key in g.files if key.startswith(pre)}
    return {key: val.requires_grad_(val.dtype.is_floating_point) for key, val in sd.items()}


def _check_grads(g, tag, sd, u, tol=2e-6):
    du = T(g[f"{tag}/du"])
    assert torch.allclose(u.grad, du, rtol=0, atol=tol * du.abs().max().item())
    gp = f"{tag}/grad/"
    n = 0
    for key in g.files:
        if key.startswith(gp):
            gref, got = T(g[key]), sd[key[len(gp):]].grad
            assert got is not None, key
            assert torch.allclose(got, gref, rtol=0, atol=tol * max(gref.abs().max().item(), 1e-30)), key
            n += 1
    assert n > 0


@pytest.mark.parametrize("tag", list(OPTION_CASES))
def test_operator_options_match_reference(g_opt, tag):
    """num_blocks / outer_mixing / post_order_ffn / short_filter_order / dropout / activation (hyena.py:447-496)."""
    sd = _load_sd(g_opt, tag)
    u = T(g_opt[f"{tag}/u"]).requires_grad_(True)
    torch.manual_seed(77)
    y = O.hyena_operator_options(u, sd, l_max=sd["filter_fn.pos_emb.t"].shape[1], shift=0.0, **OPTION_CASES[tag])
    ref = T(g_opt[f"{tag}/y"])
    assert y.shape == ref.shape and torch.allclose(y, ref, rtol=0, atol=1e-6 * ref.abs().max().item())
    (y * T(g_opt[f"{tag}/w"])).sum().backward()
    _check_grads(g_opt, tag, sd, u)


@pytest.mark.parametrize("tag", list(LONGCONV_CASES))
def test_long_conv_matches_reference(g_opt, tag):
    """LongConv + LongConvKernel (long_conv.py:107-165, long_conv_kernel.py:68-81)."""
    sd = _load_sd(g_opt, tag)
    u = T(g_opt[f"{tag}/u"]).requires_grad_(True)
    y = O.long_conv_ref(u, sd, **LONGCONV_CASES[tag])
    ref = T(g_opt[f"{tag}/y"])
    assert y.shape == ref.shape and torch.allclose(y, ref, rtol=0, atol=1e-6 * ref.abs().max().item())
    (y * T(g_opt[f"{tag}/w"])).sum().backward()
    _check_grads(g_opt, tag, sd, u)


# ---- data ingest (tests/golden/ingest.npz, make_golden.py --only-ingest) ---------------------------------------------
@pytest.fixture(scope="module")
def g_ing(golden_dir):
    return np.load(os.path.join(golden_dir, "ingest.npz"))


def test_fetch_interval_matches_reference(g_ing):
    """FastaInterval.__call__ (hg38_dataset.py:72-124): every case, with and without '.' padding, byte-exact."""
    chrom = bytes(g_ing["chrom"]).decode()
    for pad in (0, 1):
        for i, (s0, e0, ml) in enumerate(g_ing["fetch/cases"].tolist()):
            got = O.fetch_interval_ref(chrom, s0, e0, ml, pad_interval=bool(pad))
            assert got.encode() == bytes(g_ing[f"fetch/pad{pad}/{i}"]), (pad, i)


@pytest.mark.parametrize("i", [0, 1, 2])
def test_bert_mask_matches_reference(g_ing, i):
    """bert_mask (hg38_dataset.py:238-286) under the same CPU generator state: bit-exact."""
    seq = T(g_ing[f"bert/{i}/seq"])
    torch.manual_seed(500 + i)
    (o, m, l), _ = O.bert_mask_ref(seq, 3, 4, int(g_ing[f"bert/{i}/vocab"]), special_token_ids=g_ing[f"bert/{i}/special"].tolist())
    assert torch.equal(o, T(g_ing[f"bert/{i}/out"])) and torch.equal(m, T(g_ing[f"bert/{i}/mask"])) and torch.equal(l, T(g_ing[f"bert/{i}/labels"]))
