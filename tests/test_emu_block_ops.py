"""No-GPU suite: the add + LayerNorm glue kernel (dna_b200/csrc/hy_addln.cu, run under the CPU emulator) against
the oracle restatement of the reference Block's add -> norm step (oracle/hyena_model_oracle.py::block_add_norm,
standalone_hyenadna.py:521-525), values and gradients."""
import pytest
import torch

import parity_cases as P
from oracle.hyena_model_oracle import block_add_norm


def _case(D, rows, xdt, rdt, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, D, generator=g).to(xdt)
    res = None if rdt is None else (torch.randn(rows, D, generator=g) * 3).to(rdt)
    w = torch.randn(D, generator=g)
    b = torch.randn(D, generator=g)
    return x, res, w, b


@pytest.mark.parametrize("D", [128, 256, 512, 1024])
@pytest.mark.parametrize("rows", [1, 7, 37])
def test_add_ln_fp32_matches_oracle(emu_lib, D, rows):
    from dna_b200 import block_ops
    x, res, w, b = _case(D, rows, torch.float32, torch.float32)
    norm = torch.nn.LayerNorm(D, eps=1e-5)
    with torch.no_grad():
        norm.weight.copy_(w); norm.bias.copy_(b)
    for r in (res, None):
        xr = x.clone().requires_grad_(True)
        rr = None if r is None else r.clone().requires_grad_(True)
        y, ro = block_ops.add_layer_norm(xr, rr, norm)
        xo = x.clone().requires_grad_(True)
        ro_in = None if r is None else r.clone().requires_grad_(True)
        wo, bo = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        y_ref, r_ref = block_add_norm(xo, ro_in, wo, bo, 1e-5)
        assert P.relerr(y, y_ref) <= 2e-6 and torch.equal(ro.detach(), r_ref.detach())
        gy, gr = torch.randn_like(y_ref), torch.randn_like(r_ref)
        norm.zero_grad()
        ((y * gy).sum() + (ro * gr).sum()).backward()
        ((y_ref * gy).sum() + (r_ref * gr).sum()).backward()
        assert P.relerr(xr.grad, xo.grad) <= 5e-6
        if r is not None:
            assert P.relerr(rr.grad, ro_in.grad) <= 5e-6
        assert P.relerr(norm.weight.grad, wo.grad) <= 5e-6 and P.relerr(norm.bias.grad, bo.grad) <= 5e-6


@pytest.mark.parametrize("xdt,rdt", [(torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16),
                                     (torch.float32, torch.bfloat16), (torch.bfloat16, None)])
def test_add_ln_mixed_dtypes(emu_lib, xdt, rdt):
    """bf16 hidden + fp32 stream is the autocast case; the stream is rounded exactly like torch's promoted add."""
    from dna_b200 import kernels as K
    D, rows = 256, 19
    x, res, w, b = _case(D, rows, xdt, rdt, seed=3)
    sdt = xdt if rdt is None else torch.promote_types(xdt, rdt)
    rin = None if res is None else res.to(sdt)
    for ydt in (torch.float32, torch.bfloat16):
        y, ro, mean, rstd = K.add_ln_fwd(x, rin, w, b, 1e-5, ydt, sdt, write_res=True)
        y_ref, r_ref = block_add_norm(x, res, w, b, 1e-5)          # LayerNorm over the promoted sum, in fp32
        assert ro.dtype == sdt and torch.equal(ro, r_ref)
        y_ref = torch.nn.functional.layer_norm(r_ref.float(), (D,), w, b, 1e-5)
        if ydt == torch.bfloat16:
            assert y.dtype == torch.bfloat16
            assert (y.float() - y_ref.to(torch.bfloat16).float()).abs().max() <= 2 ** -7 * y_ref.abs().max()
        else:
            assert P.relerr(y, y_ref) <= 2e-6
        # backward with bf16 dy (what the Linear after an autocast LayerNorm hands back)
        dy = torch.randn(rows, D).to(ydt)
        dr = torch.randn(rows, D).to(sdt)
        dx, dres, dg, db = K.add_ln_bwd(dy, dr, ro, mean, rstd, w, xdt, True, True)
        rr = r_ref.float().clone().requires_grad_(True)
        wo, bo = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        torch.nn.functional.layer_norm(rr, (D,), wo, bo, 1e-5).backward(dy.float())
        want = rr.grad + dr.float()
        assert dres.dtype == sdt and dx.dtype == xdt
        tol = 5e-6 if sdt == torch.float32 else 2 ** -7
        assert P.relerr(dres.float(), want) <= tol
        assert P.relerr(dx.float(), want) <= (5e-6 if xdt == torch.float32 else 2 ** -7)
        assert P.relerr(dg, wo.grad) <= 1e-5 and P.relerr(db, bo.grad) <= 1e-5


def test_add_ln_errors(emu_lib):
    from dna_b200 import _lib, block_ops, kernels as K
    assert not K.add_ln_supported(96) and K.add_ln_supported(256)
    x = torch.randn(4, 96)
    norm = torch.nn.LayerNorm(96)
    assert not block_ops.add_layer_norm_supported(norm, x)
    with pytest.raises(NotImplementedError):
        block_ops.add_layer_norm(x, None, norm)
    with pytest.raises(_lib.HyenaB200Error):
        K.add_ln_fwd(x, None, norm.weight.detach(), norm.bias.detach(), 1e-5, torch.float32, torch.float32, True)
    # residual_in_fp32 with a bf16 stream is the one dtype pattern the fused step does not reproduce
    n2 = torch.nn.LayerNorm(128)
    xb = torch.randn(4, 128).to(torch.bfloat16)
    assert not block_ops.add_layer_norm_supported(n2, xb, xb, residual_in_fp32=True)
    assert block_ops.add_layer_norm_supported(n2, xb, xb.float(), residual_in_fp32=True)


def test_harness_block_uses_fused_step(emu_lib):
    """The harness Block with the fused step equals the same Block with the reference's three statements."""
    from dna_b200.standalone import add_norm
    torch.manual_seed(0)
    D = 128
    norm = torch.nn.LayerNorm(D)
    drop = torch.nn.Dropout(0.0)
    h, r = torch.randn(2, 9, D), torch.randn(2, 9, D)
    y1, r1 = add_norm(drop, norm, h, r, False, True)
    y2, r2 = add_norm(drop, norm, h, r, False, False)
    assert P.relerr(y1, y2) <= 2e-6 and torch.equal(r1, r2)
    y3, r3 = add_norm(drop, norm, h, None, True, True)
    assert r3 is h and P.relerr(y3, torch.nn.functional.layer_norm(h, (D,), norm.weight, norm.bias)) <= 2e-6
    # active dropout keeps the unfused statements
    drop2 = torch.nn.Dropout(0.5).train()
    y4, r4 = add_norm(drop2, norm, h, r, False, True)
    assert not torch.equal(r4, r1)


@pytest.mark.parametrize("cfg", [(torch.float32, torch.float32, True), (torch.float32, torch.float32, False),
                                 (torch.bfloat16, torch.float32, True), (torch.bfloat16, torch.bfloat16, False)])
def test_add_ln_with_dropout_mask(emu_lib, cfg):
    """dropout -> add -> LayerNorm in one kernel for a given keep mask (embed_dropout = 0.1 of the training configs)."""
    xdt, rdt, with_res = cfg
    errs = P.add_ln_dropout_case("cpu", xdt, rdt, with_res)
    tol = 5e-6 if xdt == torch.float32 else 2e-2
    assert errs.pop("stream_equal") == 0.0
    for n, e in errs.items():
        assert e <= tol, (cfg, n, e)


def test_harness_block_fuses_active_dropout(emu_lib):
    """the first block's dropout1 (embed_dropout = 0.1, training mode) no longer leaves the fused path"""
    from dna_b200 import kernels as K
    from dna_b200.standalone import add_norm
    norm = torch.nn.LayerNorm(128)
    drop = torch.nn.Dropout(0.1)
    drop.train()
    x = torch.randn(9, 128)
    n0 = K.launch_count()
    h, r = add_norm(drop, norm, x, None, False, True)
    assert K.launch_count() - n0 == 1
    kept = r != 0
    assert torch.allclose(r[kept], (x / 0.9)[kept]) and 0.02 < (~kept).float().mean() < 0.25
