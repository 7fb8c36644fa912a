"""Generate golden vectors by RUNNING THE REFERENCE ITSELF (CPU) in the build container.

    python tests/golden/make_golden.py         # needs /root/reference; writes tests/golden/*.npz

The reference (open-genome/dna) has no tests or fixtures for the Hyena hot path, so
these files are what pins oracle/hyena_oracle.py: outputs of the reference's own functions/modules
on seeded inputs.  /root/reference does not exist on the GPU box, hence the committed .npz files.

Imports used from the reference:
  src/models/sequence/hyena.py      fftconv_ref, HyenaFilter, HyenaOperator   (needs 4 stub modules)
  src/ops/fftconv.py                not importable (needs the absent CUDA ext `fftconv`) -> its pure
                                    torch functions fftconv_ref / fftconv_h3_ref are exec'd from source
  standalone_hyenadna.py            fftconv, HyenaOperator, HyenaDNAModel, CharacterTokenizer
  src/dataloaders/datasets/hg38_char_tokenizer.py   CharacterTokenizer (methods only, see below)
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True


def install_stubs():
    """hydra / omegaconf / pytorch_lightning / opt_einsum are absent from this image."""
    def dotted(path):
        mod, _, name = path.rpartition(".")
        return getattr(importlib.import_module(mod), name)

    hydra = types.ModuleType("hydra")
    hydra.utils = types.ModuleType("hydra.utils")
    hydra.utils.get_method = dotted
    hydra.utils.get_class = dotted
    omegaconf = types.ModuleType("omegaconf")
    omegaconf.DictConfig = dict
    omegaconf.ListConfig = list
    omegaconf.OmegaConf = type("OmegaConf", (), {})
    pl = types.ModuleType("pytorch_lightning")
    pl.utilities = types.ModuleType("pytorch_lightning.utilities")
    pl.utilities.rank_zero_only = lambda f: f
    oe = types.ModuleType("opt_einsum")
    oe.contract = torch.einsum
    for name, mod in (("hydra", hydra), ("hydra.utils", hydra.utils), ("omegaconf", omegaconf),
                      ("pytorch_lightning", pl), ("pytorch_lightning.utilities", pl.utilities), ("opt_einsum", oe)):
        sys.modules.setdefault(name, mod)


def np_(t):
    t = t.detach()
    if t.dtype == torch.bfloat16:
        return t.float().numpy()
    return t.numpy()


def gen_fftconv(hy, sa):
    out = {}
    g = torch.Generator().manual_seed(11)
    for tag, (B, H, L) in {"a": (2, 3, 64), "b": (1, 4, 100), "c": (2, 2, 257)}.items():
        u = torch.randn(B, H, L, generator=g)
        k = torch.randn(H, L, generator=g) * torch.exp(-torch.arange(L) / (L / 3.0))
        D = torch.randn(H, generator=g)
        out[f"{tag}_u"], out[f"{tag}_k"], out[f"{tag}_D"] = np_(u), np_(k), np_(D)
        out[f"{tag}_hy_nogelu"] = np_(hy.fftconv_ref(u, k, D, None, gelu=False))
        out[f"{tag}_hy_gelu"] = np_(hy.fftconv_ref(u, k, D, None, gelu=True))
        out[f"{tag}_standalone"] = np_(sa.fftconv(u, k, D))
        ub = u.to(torch.bfloat16)
        out[f"{tag}_hy_bf16"] = np_(hy.fftconv_ref(ub, k, D, None, gelu=False))
        # 5-D call shape used by HyenaOperator (hyena.py:447-453, 484): u [B,1,H,1,L], bias [1,H,1]
        out[f"{tag}_hy_5d"] = np_(hy.fftconv_ref(u[:, None, :, None, :], k, D[None, :, None], None, gelu=False))
        # gradients through the reference (autograd)
        u2 = u.clone().requires_grad_(True)
        k2 = k.clone().requires_grad_(True)
        D2 = D.clone().requires_grad_(True)
        w = torch.randn(B, H, L, generator=g)
        out[f"{tag}_w"] = np_(w)
        (hy.fftconv_ref(u2, k2, D2, None, gelu=False) * w).sum().backward()
        out[f"{tag}_du"], out[f"{tag}_dk"], out[f"{tag}_dD"] = np_(u2.grad), np_(k2.grad), np_(D2.grad)
    # pure-torch functions of src/ops/fftconv.py (module import fails on `from fftconv import ...`)
    src = open(os.path.join(REF, "src/ops/fftconv.py")).read().replace("from fftconv import fftconv_fwd, fftconv_bwd", "")
    src = src.replace("@torch.jit.script", "")   # TorchScript needs a real module; semantics unchanged
    ns = {"__name__": "ref_ops_fftconv"}
    exec(compile(src, "src/ops/fftconv.py", "exec"), ns)
    B, H, L = 2, 4, 96
    kk, vv, qq = (torch.randn(B, H, L, generator=g) for _ in range(3))
    ssm = torch.randn(H, L, generator=g) * torch.exp(-torch.arange(L) / 20.0)
    D = torch.randn(H, generator=g)
    out["h3_k"], out["h3_v"], out["h3_q"], out["h3_ssm"], out["h3_D"] = map(np_, (kk, vv, qq, ssm, D))
    out["h3_out_hd1"] = np_(ns["fftconv_h3_ref"](kk, ssm, D, qq, vv, head_dim=1))
    out["ops_ref_nogelu"] = np_(ns["fftconv_ref"](kk, ssm, D, None, gelu=False))
    np.savez_compressed(os.path.join(OUT, "fftconv.npz"), **out)


def gen_filter_and_operator(hy, sa):
    out = {}
    cfgs = {
        "src": dict(mod="src", d_model=16, l_max=70, L=64, B=2, kw=dict(emb_dim=5, filter_order=64, w=10, lr=6e-4, wd=0, lr_pos_emb=0, modulate=True)),
        "src_e3": dict(mod="src", d_model=8, l_max=40, L=33, B=1, kw=dict(emb_dim=3, filter_order=16, w=1, lr_pos_emb=0)),
        "sa": dict(mod="sa", d_model=16, l_max=130, L=128, B=2, kw=dict(emb_dim=5, filter_order=64, w=10, lr=6e-4, wd=0, lr_pos_emb=0)),
        "sa_trunc": dict(mod="sa", d_model=8, l_max=50, L=60, B=1, kw=dict(emb_dim=5, filter_order=64, w=10, lr_pos_emb=0)),
    }
    for tag, c in cfgs.items():
        torch.manual_seed(2222)
        if c["mod"] == "src":
            op = hy.HyenaOperator(d_model=c["d_model"], l_max=c["l_max"], layer_idx=0, device=None, dtype=None, **c["kw"])
            shift = op.filter_fn.modulation.shift
        else:
            op = sa.HyenaOperator(d_model=c["d_model"], l_max=c["l_max"], **c["kw"])
            shift = op.filter_fn.modulation.shift
        sd = op.state_dict()
        for key, val in sd.items():
            out[f"{tag}/sd/{key}"] = np_(val)
        out[f"{tag}/shift"] = np.array(shift, dtype=np.float64)
        out[f"{tag}/l_max"] = np.array(c["l_max"])
        Lf = min(c["L"], c["l_max"])
        out[f"{tag}/filter"] = np_(op.filter_fn.filter(Lf))
        u = torch.randn(c["B"], c["L"], c["d_model"], requires_grad=True)
        w = torch.randn(c["B"], Lf, c["d_model"])
        y = op(u)
        (y * w).sum().backward()
        out[f"{tag}/u"], out[f"{tag}/w"], out[f"{tag}/y"], out[f"{tag}/du"] = np_(u), np_(w), np_(y), np_(u.grad)
        for name, prm in op.named_parameters():
            if prm.grad is not None:
                out[f"{tag}/grad/{name}"] = np_(prm.grad)
    np.savez_compressed(os.path.join(OUT, "operator.npz"), **out)


def gen_model(sa):
    """BASELINE config C1: standalone tiny HyenaDNA (2 layers, d_model=128, seqlen 1024) — scaled to
    seqlen 256 / d_model 32 to keep the fixture small; same code path."""
    out = {}
    torch.manual_seed(2222)
    model = sa.HyenaDNAModel(d_model=32, n_layer=2, d_inner=128, vocab_size=12, embed_dropout=0.0,
                             layer=dict(l_max=258, emb_dim=5, filter_order=64, short_filter_order=3, modulate=True,
                                        w=10, lr=6e-4, wd=0.0, lr_pos_emb=0.0))
    model.eval()
    for key, val in model.state_dict().items():
        out[f"sd/{key}"] = np_(val)
    ids = torch.randint(7, 11, (2, 256), generator=torch.Generator().manual_seed(0))
    h = model(ids)
    loss = h.float().pow(2).mean()
    loss.backward()
    out["ids"], out["hidden"], out["loss"] = ids.numpy(), np_(h), np_(loss)
    out["grad/backbone.layers.0.mixer.in_proj.weight"] = np_(model.backbone.layers[0].mixer.in_proj.weight.grad)
    out["grad/backbone.layers.1.mixer.filter_fn.bias"] = np_(model.backbone.layers[1].mixer.filter_fn.bias.grad)
    out["grad/backbone.embeddings.word_embeddings.weight"] = np_(model.backbone.embeddings.word_embeddings.weight.grad)
    np.savez_compressed(os.path.join(OUT, "model_tiny.npz"), **out)


def gen_tokenizer(sa):
    """The reference tokenizers cannot be constructed under transformers 5.5 (they pin 4.28): the HF
    base-class __init__ is bypassed so that the reference's OWN vocabulary / _tokenize /
    _convert_token_to_id / build_inputs_with_special_tokens run; HF's generic padding/truncation
    (padding='max_length', truncation=True, padding_side='left') cannot be executed here and is
    restated in oracle.tokenize_ref from the 4.28 documentation (that part stays unpinned)."""
    import transformers
    base = transformers.PreTrainedTokenizer
    spec = importlib.util.spec_from_file_location("ref_char_tok", os.path.join(REF, "src/dataloaders/datasets/hg38_char_tokenizer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    orig = base.__init__
    base.__init__ = lambda self, *a, **k: None
    try:
        for tag, cls in (("src", mod.CharacterTokenizer), ("sa", sa.CharacterTokenizer)):
            tok = cls.__new__(cls)
            try:
                cls.__init__(tok, characters=["A", "C", "G", "T", "N"], model_max_length=32)
            except Exception as e:  # attribute setters of the HF base may still object; vocab is set before/after
                print("tokenizer init note:", tag, repr(e))
            vocab = dict(tok._vocab_str_to_int)
            text = "ACGTNacgtn.XRY-ACCGT"
            ids = [tok._convert_token_to_id(t) for t in tok._tokenize(text)]
            out[f"{tag}/text"] = np.frombuffer(text.encode(), dtype=np.uint8)
            out[f"{tag}/ids"] = np.array(ids)
            out[f"{tag}/vocab_keys"] = np.array(list(vocab.keys()))
            out[f"{tag}/vocab_vals"] = np.array(list(vocab.values()))
            # special-token layout with the ids the vocab assigns to [SEP]/[CLS]
            for name in ("sep_token_id", "cls_token_id"):
                try:
                    setattr(type(tok), name, property(lambda self, n=name: vocab["[SEP]"] if n.startswith("sep") else vocab["[CLS]"]))
                except Exception:
                    pass
            out[f"{tag}/with_special"] = np.array(tok.build_inputs_with_special_tokens(ids))
    finally:
        base.__init__ = orig
    np.savez_compressed(os.path.join(OUT, "tokenizer.npz"), **out)


def gen_revcomp():
    """string_reverse_complement of src/dataloaders/datasets/hg38_dataset.py:28-38 (the rc_aug path, :118-119) on
    fixed and random strings over the FASTA alphabet incl. lower case, N and '.'."""
    for name in ("pyfaidx", "polars"):         # absent here; only used by the FASTA / BED readers of that module
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.Fasta = object
            sys.modules[name] = m
    # loaded by file path: the package __init__ chain (src.dataloaders) pulls in the whole training stack
    spec = importlib.util.spec_from_file_location("ref_hg38_dataset", os.path.join(REF, "src/dataloaders/datasets/hg38_dataset.py"))
    ds = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ds)
    rng = np.random.default_rng(0)
    alphabet = np.frombuffer(b"ACGTNacgtn.XRY-", dtype=np.uint8)
    out = {}
    cases = [b"", b"A", b"ACGTNacgtn.XRY-ACCGT", bytes(alphabet[rng.integers(0, len(alphabet), size=257)]),
             bytes(alphabet[rng.integers(0, 5, size=4096)])]
    for i, c in enumerate(cases):
        out[f"in{i}"] = np.frombuffer(c, dtype=np.uint8)
        out[f"out{i}"] = np.frombuffer(ds.string_reverse_complement(c.decode()).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "revcomp.npz"), **out)


def gen_features(hy, sa):
    """SURVEY section 8(f) rows that were built after the first fixtures: order-3 operators (hyena.py:475-484, standalone
    :286-288), the bidirectional long convolution (hyena.py:68-74), k_rev (:64-66), the H3 multi-head form
    (src/ops/fftconv.py:38-55 with head_dim > 1) and a bidirectional HyenaOperator."""
    out = {}
    g = torch.Generator().manual_seed(23)
    # long convolution variants of hyena.fftconv_ref
    for tag, (B, H, L) in {"bi_a": (2, 3, 64), "bi_b": (1, 2, 101), "bi_c": (2, 2, 300)}.items():
        u = torch.randn(B, H, L, generator=g, requires_grad=True)
        k = (torch.randn(H, L, generator=g) * torch.exp(-torch.arange(L) / (L / 3.0))).requires_grad_(True)
        kr = (torch.randn(H, L, generator=g) * torch.exp(-torch.arange(L) / (L / 5.0))).requires_grad_(True)
        D = torch.randn(H, generator=g, requires_grad=True)
        w = torch.randn(B, H, L, generator=g)
        out[f"{tag}_u"], out[f"{tag}_k"], out[f"{tag}_krev"], out[f"{tag}_D"], out[f"{tag}_w"] = map(np_, (u, k, kr, D, w))
        for name, kw in (("bidir", dict(bidirectional=True)), ("krev", dict(k_rev=kr)), ("bidir_krev", dict(k_rev=kr, bidirectional=True))):
            y = hy.fftconv_ref(u, k, D, None, gelu=False, **kw)
            gr = torch.autograd.grad((y * w).sum(), [u, k, D] + ([kr] if "k_rev" in kw else []))
            out[f"{tag}_{name}_y"] = np_(y)
            for n_, g_ in zip(["du", "dk", "dD", "dkrev"], gr):
                out[f"{tag}_{name}_{n_}"] = np_(g_)
    # H3 multi-head form (pure functions of src/ops/fftconv.py, exec'd as in gen_fftconv)
    src = open(os.path.join(REF, "src/ops/fftconv.py")).read().replace("from fftconv import fftconv_fwd, fftconv_bwd", "")
    src = src.replace("@torch.jit.script", "")
    ns = {"__name__": "ref_ops_fftconv"}
    exec(compile(src, "src/ops/fftconv.py", "exec"), ns)
    for hd in (2, 8):
        B, Hh, L = 2, 3, 80
        H = Hh * hd
        kk, vv, qq = (torch.randn(B, H, L, generator=g, requires_grad=True) for _ in range(3))
        ssm = (torch.randn(Hh, L, generator=g) * torch.exp(-torch.arange(L) / 20.0)).requires_grad_(True)
        D = torch.randn(Hh, generator=g, requires_grad=True)
        w = torch.randn(B, H, L, generator=g)
        y = ns["fftconv_h3_ref"](kk, ssm, D, qq, vv, head_dim=hd)
        gr = torch.autograd.grad((y * w).sum(), [kk, ssm, D, qq, vv])
        pre = f"h3_hd{hd}_"
        for n_, t_ in zip(["k", "v", "q", "ssm", "D", "w", "y", "dk", "dssm", "dD", "dq", "dv"],
                          [kk, vv, qq, ssm, D, w, y, *gr]):
            out[pre + n_] = np_(t_)
    # operators: order 3 (src '(v o)' / standalone '(o v)' filter layouts) and bidirectional
    cfgs = {
        "o3_src": dict(mod="src", d_model=8, l_max=64, L=50, B=2, kw=dict(order=3, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
        "o3_sa": dict(mod="sa", d_model=8, l_max=64, L=64, B=1, kw=dict(order=3, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
        "o4_src": dict(mod="src", d_model=4, l_max=40, L=40, B=2, kw=dict(order=4, emb_dim=3, filter_order=16, w=2, lr_pos_emb=0)),
        "bidir_src": dict(mod="src", d_model=8, l_max=72, L=70, B=2, kw=dict(bidirectional=True, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
    }
    for tag, c in cfgs.items():
        torch.manual_seed(2222)
        if c["mod"] == "src":
            op = hy.HyenaOperator(d_model=c["d_model"], l_max=c["l_max"], layer_idx=0, device=None, dtype=None, **c["kw"])
        else:
            op = sa.HyenaOperator(d_model=c["d_model"], l_max=c["l_max"], **c["kw"])
        for key, val in op.state_dict().items():
            out[f"{tag}/sd/{key}"] = np_(val)
        out[f"{tag}/shift"] = np.array(op.filter_fn.modulation.shift, dtype=np.float64)
        u = torch.randn(c["B"], c["L"], c["d_model"], requires_grad=True)
        w = torch.randn(c["B"], min(c["L"], c["l_max"]), c["d_model"])
        y = op(u)
        (y * w).sum().backward()
        out[f"{tag}/u"], out[f"{tag}/w"], out[f"{tag}/y"], out[f"{tag}/du"] = np_(u), np_(w), np_(y), np_(u.grad)
        for name, prm in op.named_parameters():
            if prm.grad is not None:
                out[f"{tag}/grad/{name}"] = np_(prm.grad)
    np.savez_compressed(os.path.join(OUT, "features.npz"), **out)


def gen_options(hy):
    """Round-2 second batch (options.npz): HyenaOperator options outside the HyenaDNA configs that the reference's forward
    really runs (num_heads > 1 and inner_factor != 1 raise inside the reference itself): num_blocks (hyena.py:447-453),
    outer_mixing (:476-479), post_order_ffn (:487-492), short_filter_order != 3 (:407-413), dropout in training (:481,
    CPU RNG stream seeded right before the call), and the LongConv layer (long_conv.py:19-175) causal / bidirectional."""
    out = {}
    cfgs = {
        "blocks2": dict(d_model=8, l_max=64, L=64, B=2, kw=dict(num_blocks=2, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
        "blocks4_o3": dict(d_model=4, l_max=96, L=96, B=1, kw=dict(num_blocks=4, order=3, emb_dim=3, filter_order=16, w=2, lr_pos_emb=0)),
        "outer": dict(d_model=6, l_max=50, L=50, B=2, kw=dict(outer_mixing=True, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
        "ffn_o3": dict(d_model=6, l_max=48, L=40, B=2, kw=dict(post_order_ffn=True, order=3, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
        "short5": dict(d_model=8, l_max=70, L=70, B=2, kw=dict(short_filter_order=5, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
        "drop": dict(d_model=8, l_max=64, L=64, B=2, train=True, kw=dict(dropout=0.25, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
        "gelu_act": dict(d_model=8, l_max=40, L=40, B=1, kw=dict(activation="gelu", emb_dim=5, filter_order=16, w=4, lr_pos_emb=0)),
    }
    for tag, c in cfgs.items():
        torch.manual_seed(2222)
        op = hy.HyenaOperator(d_model=c["d_model"], l_max=c["l_max"], layer_idx=0, device=None, dtype=None, **c["kw"])
        op.train(bool(c.get("train")))
        for key, val in op.state_dict().items():
            out[f"{tag}/sd/{key}"] = np_(val)
        u = torch.randn(c["B"], c["L"], c["d_model"], requires_grad=True)
        w = torch.randn(c["B"], min(c["L"], c["l_max"]), c["d_model"])
        torch.manual_seed(77)                     # the dropout masks are drawn from this stream
        y = op(u)
        (y * w).sum().backward()
        out[f"{tag}/u"], out[f"{tag}/w"], out[f"{tag}/y"], out[f"{tag}/du"] = np_(u), np_(w), np_(y), np_(u.grad)
        for name, prm in op.named_parameters():
            if prm.grad is not None:
                out[f"{tag}/grad/{name}"] = np_(prm.grad)
    lc = importlib.import_module("src.models.sequence.long_conv")
    lcfgs = {
        "lc_causal": dict(kw=dict(d_model=8, l_max=64, channels=1, lam=0.001), shape=(2, 8, 64)),
        "lc_bidir": dict(kw=dict(d_model=6, l_max=50, channels=1, bidirectional=True, lam=0.001), shape=(2, 6, 50)),
        "lc_ch2_bld": dict(kw=dict(d_model=8, l_max=64, channels=2, transposed=False, lam=0.0005, postact="glu", activation="gelu"), shape=(2, 40, 8)),
        "lc_bidir_short": dict(kw=dict(d_model=4, l_max=64, channels=2, bidirectional=True, lam=0.001, postact=None), shape=(1, 4, 33)),
    }
    for tag, c in lcfgs.items():
        torch.manual_seed(2222)
        m = lc.LongConv(**c["kw"])
        m.eval()
        for key, val in m.state_dict().items():
            out[f"{tag}/sd/{key}"] = np_(val)
        u = torch.randn(*c["shape"], requires_grad=True)
        y, _ = m(u)
        w = torch.randn(*y.shape)
        (y * w).sum().backward()
        out[f"{tag}/u"], out[f"{tag}/w"], out[f"{tag}/y"], out[f"{tag}/du"] = np_(u), np_(w), np_(y), np_(u.grad)
        for name, prm in m.named_parameters():
            if prm.grad is not None:
                out[f"{tag}/grad/{name}"] = np_(prm.grad)
    np.savez_compressed(os.path.join(OUT, "options.npz"), **out)


def gen_ingest():
    """ingest.npz: FastaInterval.__call__ (hg38_dataset.py:72-124, constructed without pyfaidx: seqs = dict of strings)
    and bert_mask (:238-286, CPU RNG seeded right before the call)."""
    for name in ("pyfaidx", "polars"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.Fasta = object
            sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("ref_hg38_dataset", os.path.join(REF, "src/dataloaders/datasets/hg38_dataset.py"))
    ds = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ds)
    rng = np.random.default_rng(3)
    chrom = bytes(np.frombuffer(b"ACGTNacgtn", dtype=np.uint8)[rng.integers(0, 10, size=5000)]).decode()
    out = {"chrom": np.frombuffer(chrom.encode(), dtype=np.uint8)}
    cases = [(100, 400, 300), (100, 400, 512), (0, 50, 300), (4900, 5000, 400), (10, 4000, 1000), (2400, 2600, 5600),
             (0, 5000, 5000), (7, 8, 16), (4990, 5000, 33), (1000, 1001, 1)]
    rows = []
    for pad in (False, True):
        fi = ds.FastaInterval.__new__(ds.FastaInterval)
        fi.seqs = {"chr1": chrom}
        fi.chr_lens = {"chr1": len(chrom)}
        fi.return_seq_indices, fi.shift_augs, fi.rc_aug, fi.pad_interval = False, None, False, pad
        for i, (s0, e0, ml) in enumerate(cases):
            seq = fi("chr1", s0, e0, ml)
            out[f"fetch/pad{int(pad)}/{i}"] = np.frombuffer(seq.encode(), dtype=np.uint8)
    out["fetch/cases"] = np.array(cases)
    for i, (shape, vocab, special) in enumerate([((4, 64), 12, [0, 1, 2, 3, 4, 5, 6]), ((2, 1000), 4096, [0, 1, 2, 3, 4]), ((1, 7), 12, [0, 1, 2, 3, 4, 5, 6])]):
        g = torch.Generator().manual_seed(100 + i)
        seq = torch.randint(5 if vocab > 12 else 7, vocab, shape, generator=g)
        seq[:, : shape[1] // 8] = 4            # left padding
        torch.manual_seed(500 + i)
        o, m, l = ds.bert_mask(seq.clone(), 3, 4, vocab, special_token_ids=special)
        out[f"bert/{i}/seq"], out[f"bert/{i}/out"], out[f"bert/{i}/mask"], out[f"bert/{i}/labels"] = seq.numpy(), o.numpy(), m.numpy(), l.numpy()
        out[f"bert/{i}/vocab"], out[f"bert/{i}/special"] = np.array(vocab), np.array(special)
    np.savez_compressed(os.path.join(OUT, "ingest.npz"), **out)


def main():
    sys.path.insert(0, REF)
    install_stubs()
    torch.set_num_threads(1)
    if "--only-ingest" in sys.argv:
        gen_ingest()
        print("ingest.npz", os.path.getsize(os.path.join(OUT, "ingest.npz")), "bytes")
        return
    if "--only-revcomp" in sys.argv:      # the other fixtures are not regenerated (their bytes are committed)
        gen_revcomp()
        print("revcomp.npz", os.path.getsize(os.path.join(OUT, "revcomp.npz")), "bytes")
        return
    hy = importlib.import_module("src.models.sequence.hyena")
    sa = importlib.import_module("standalone_hyenadna")
    assert hy.fftconv_func is None, "reference fused path unexpectedly importable"
    if "--only-features" in sys.argv:     # added in round 2; the round-1 fixtures keep their committed bytes
        gen_features(hy, sa)
        print("features.npz", os.path.getsize(os.path.join(OUT, "features.npz")), "bytes")
        return
    if "--only-options" in sys.argv:      # second round-2 batch
        gen_options(hy)
        print("options.npz", os.path.getsize(os.path.join(OUT, "options.npz")), "bytes")
        return
    gen_fftconv(hy, sa)
    gen_filter_and_operator(hy, sa)
    gen_model(sa)
    gen_tokenizer(sa)
    gen_revcomp()
    gen_features(hy, sa)
    gen_options(hy)
    gen_ingest()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
