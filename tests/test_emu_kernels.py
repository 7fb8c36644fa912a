"""No-GPU suite: the CUDA kernel SOURCES (dna_b200/csrc) compiled by g++ against the execution-model
emulator tests/emu/cuda_emu.h and driven through the same ctypes binding as the shipped library.
This exercises every kernel's index algebra (all transform lengths, both regimes, every gating mode,
forward and backward) against the oracle where no GPU exists.  The sm_100a build itself is tested by
tests/test_gpu_*.py (-m gpu)."""
import ctypes
import os

import pytest
import torch

import parity_cases as P


@pytest.mark.parametrize("mode", ["plain", "gated", "shortconv"])
@pytest.mark.parametrize("shape", [(2, 3, 100), (3, 2, 257), (1, 2, 1000), (9, 1, 300), (1, 1, 4096), (2, 1, 2047), (1, 2, 1)])
def test_fused_regime_fp32(emu_lib, mode, shape):
    errs = P.conv_case(*shape, mode=mode, device="cpu")
    for name, e in errs.items():
        assert e <= P.FP32_TOL, (mode, shape, name, e)


@pytest.mark.parametrize("shape,mode", [((1, 2, 5000), "plain"), ((1, 2, 5000), "shortconv"), ((1, 1, 20000), "shortconv"),
                                        ((2, 1, 4097), "gated"), ((1, 1, 100000), "shortconv")])
def test_four_step_regime_production_rows(emu_lib, shape, mode):
    """M > 4096: M = M1 x 4096 (the shipped configuration), M1 = 2, 8, 32."""
    errs = P.conv_case(*shape, mode=mode, device="cpu")
    for name, e in errs.items():
        assert e <= P.FP32_TOL, (mode, shape, name, e)


@pytest.mark.parametrize("shape,mode", [((1, 2, 5000), "plain"), ((2, 2, 5000), "shortconv"), ((2, 1, 4097), "gated"),
                                        ((1, 1, 20000), "shortconv")])
def test_four_step_saved_spectrum_backward(emu_lib, shape, mode):
    """The forward keeps the row-transformed spectrum of g; the backward reads it back, transforms dy only and dD
    comes out as dk[:, 0] (hy_conv_fwd_args.gsave / hy_conv_bwd_args.gsave)."""
    errs = P.conv_case(*shape, mode=mode, device="cpu", gsave=True)
    for name, e in errs.items():
        assert e <= P.FP32_TOL, (mode, shape, name, e)


@pytest.mark.parametrize("L", [300, 1000, 4096, 16000, 65536])
def test_four_step_saved_spectrum_all_column_lengths(emu_lib, L):
    emu_lib.hy_debug_set_block.restype = ctypes.c_int
    emu_lib.hy_debug_set_block(256)
    try:
        errs = P.conv_case(1, 1, L, mode="shortconv", device="cpu", seed=L, gsave=True)
        if L <= 1000:
            errs.update({"plain_" + k: v for k, v in P.conv_case(3, 2, L, mode="plain", device="cpu", seed=L, gsave=True).items()})
    finally:
        emu_lib.hy_debug_set_block(0)
    for name, e in errs.items():
        assert e <= 5e-5, (L, name, e)


@pytest.mark.parametrize("L", [300, 1000, 2000, 4096, 8192, 16000, 32768, 65536, 100000])
def test_four_step_all_column_lengths(emu_lib, L):
    """Every column-transform length M1 = 2 .. 512 (1, 2 and 3 passes) with 256-point rows so the
    cases stay small: the same kernels the 1 M-nt configuration (M1 = 256) uses."""
    emu_lib.hy_debug_set_block.restype = ctypes.c_int
    emu_lib.hy_debug_set_block(256)
    try:
        errs = P.conv_case(1, 1, L, mode="shortconv", device="cpu", seed=L)
        if L <= 2000:
            errs.update({"plain_" + k: v for k, v in P.conv_case(2, 2, L, mode="plain", device="cpu", seed=L).items()})
    finally:
        emu_lib.hy_debug_set_block(0)
    for name, e in errs.items():
        assert e <= 5e-5, (L, name, e)      # dsb sums 1e5 fp32 terms


def _fft_len(lib, L):
    lib.hy_fft_len.restype = ctypes.c_int
    return lib.hy_fft_len(ctypes.c_int(L))


# L -> column length M1 with 256-point rows: every instance of the 5 * 2^a and 3 * 2^a families
ODD_COLS_AT_256 = {2049: 10, 2560: 10, 3071: 12, 3072: 12, 2500: 10, 3000: 12, 5000: 20, 6000: 24, 10000: 40, 12288: 48, 20000: 80, 24001: 96, 40000: 160,
                   49152: 192, 80000: 320, 98304: 384}


@pytest.mark.parametrize("L", sorted(ODD_COLS_AT_256))
def test_four_step_odd_column_lengths(emu_lib, L):
    """Transform lengths 5 * 2^k and 3 * 2^k (the reference transforms exactly 2L points, hyena.py:61-62): column plans
    with a radix-5 / radix-3 last pass (M1 = 10 ... 384; 1, 2 and 3 passes), forward and backward, both backward
    variants; same tolerance as the power-of-two lengths."""
    emu_lib.hy_debug_set_block.restype = ctypes.c_int
    emu_lib.hy_debug_set_block(256)
    try:
        assert _fft_len(emu_lib, L) == 256 * ODD_COLS_AT_256[L]
        errs = P.conv_case(1, 1, L, mode="shortconv", device="cpu", seed=L)
        errs.update({"gsave_" + k: v for k, v in P.conv_case(1, 1, L, mode="shortconv", device="cpu", seed=L, gsave=True).items()})
        if L <= 6000:
            errs.update({"plain_" + k: v for k, v in P.conv_case(2, 2, L, mode="plain", device="cpu", seed=L).items()})
            errs.update({"gated_" + k: v for k, v in P.conv_case(1, 2, L, mode="gated", device="cpu", seed=L).items()})
    finally:
        emu_lib.hy_debug_set_block(0)
    for name, e in errs.items():
        assert e <= 5e-5, (L, name, e)


@pytest.mark.parametrize("shape", [(1, 2, 40000), (1, 1, 45000), (1, 1, 70000)])
def test_odd_column_lengths_production_rows_bf16_and_fp32(emu_lib, shape):
    """4096-point rows (the shipped configuration): M = 10, 12 and 20 x 4096; fp32 parity and the staged bf16 paths."""
    L = shape[2]
    assert _fft_len(emu_lib, L) == 4096 * {40000: 10, 45000: 12, 70000: 20}[L]
    errs = P.conv_case(*shape, mode="shortconv", device="cpu", gsave=(L != 45000))
    for name, e in errs.items():
        assert e <= P.FP32_TOL, (shape, name, e)
    e_ours, e_ref, scale = P.bf16_forward_case(*shape, device="cpu")
    assert e_ours <= 2 * e_ref + scale * 2 ** -8, (e_ours, e_ref, scale)
    if L == 40000:
        errs = P.conv_case(*shape, mode="shortconv", device="cpu", dtype=torch.bfloat16, gsave=True)
        for name, e in errs.items():
            assert e <= 6e-2, (name, e)


def test_odd_lengths_switch(emu_lib):
    """hy_debug_set_odd_lengths(0) (HYENA_B200_POW2_ONLY=1) restores the power-of-two transform lengths."""
    emu_lib.hy_debug_set_odd_lengths.restype = ctypes.c_int
    assert _fft_len(emu_lib, 160000) == 40 * 4096
    assert _fft_len(emu_lib, 1000000) == 1 << 20 and _fft_len(emu_lib, 4000) == 4096
    emu_lib.hy_debug_set_odd_lengths(0)
    try:
        assert _fft_len(emu_lib, 160000) == 1 << 18
    finally:
        emu_lib.hy_debug_set_odd_lengths(1)


@pytest.mark.parametrize("shape", [(2, 3, 100), (1, 2, 1000), (2, 2, 3000)])
def test_bf16_forward_not_worse_than_reference_bf16(emu_lib, shape):
    e_ours, e_ref, scale = P.bf16_forward_case(*shape, device="cpu")
    assert e_ours <= 2 * e_ref + scale * 2 ** -8, (e_ours, e_ref, scale)


@pytest.mark.parametrize("shape", [(1, 2, 20000), (2, 1, 8192), (1, 1, 100000)])
def test_bf16_four_step_with_cp_async_staging(emu_lib, shape):
    """bf16 + aligned rows take the cp.async-staged prologue/epilogue of the column kernels (M1 = 8, 2, 32)."""
    e_ours, e_ref, scale = P.bf16_forward_case(*shape, device="cpu")
    assert e_ours <= 2 * e_ref + scale * 2 ** -8, (e_ours, e_ref, scale)
    if shape[2] == 20000:
        errs = P.conv_case(*shape, mode="shortconv", device="cpu", dtype=torch.bfloat16)
        for name, e in errs.items():
            assert e <= 6e-2, (name, e)
        # saved-spectrum backward: the dy-only phase A stages the x0 source row; same numbers as the recompute path
        errs2 = P.conv_case(*shape, mode="shortconv", device="cpu", dtype=torch.bfloat16, gsave=True)
        for name, e in errs2.items():
            assert e <= 6e-2 and abs(e - errs[name]) <= 2e-3, (name, e, errs[name])


@pytest.mark.parametrize("L", [25000, 24997])
def test_bf16_vectorised_gate_sweeps_of_the_column_kernels(emu_lib, L):
    """M1 = 128 (two column passes): the staged bf16 tiles are gated by the 8-samples-per-step sweeps (prologue of
    phase A for g and for dy, epilogue of phase C for z/y and for dx1/dv), the fp32 spectrum / dk rows by the
    16-byte cp.async / store fast paths; L = 24997 adds the partial chunk at the row end and unaligned fp32 rows."""
    emu_lib.hy_debug_set_block.restype = ctypes.c_int
    emu_lib.hy_debug_set_block(256)
    try:
        errs = P.conv_case(1, 2, L, mode="shortconv", device="cpu", dtype=torch.bfloat16, gsave=True)
        errs_r = P.conv_case(1, 2, L, mode="shortconv", device="cpu", dtype=torch.bfloat16, gsave=False)
        e_ours, e_ref, scale = P.bf16_forward_case(1, 2, L, device="cpu")
    finally:
        emu_lib.hy_debug_set_block(0)
    assert e_ours <= 2 * e_ref + scale * 2 ** -8, (e_ours, e_ref, scale)
    for name, e in errs.items():
        assert e <= 6e-2 and abs(e - errs_r[name]) <= 2e-3, (name, e, errs_r[name])


@pytest.mark.parametrize("mode", ["plain", "gated", "shortconv"])
def test_bf16_backward_close_to_reference_bf16(emu_lib, mode):
    # the reference's bf16 autograd rounds every intermediate to bf16: agreement is to a few bf16 ulps
    errs = P.conv_case(2, 2, 300, mode=mode, device="cpu", dtype=torch.bfloat16)
    for name, e in errs.items():
        assert e <= 6e-2, (mode, name, e)
    for name in ("out",):
        assert errs[name] <= 2e-2, (mode, name, errs[name])


@pytest.mark.parametrize("cfg", [(16, 64, 5, 2, 100, 130), (256, 64, 5, 2, 300, 400), (8, 16, 3, 2, 33, 40), (70, 64, 5, 1, 64, 64),
                                 (5, 32, 7, 3, 200, 200), (3, 8, 3, 0, 10, 12), (300, 64, 5, 2, 130, 130), (4, 16, 9, 1, 50, 50)])
def test_filter_kernel(emu_lib, cfg):
    e_ours, e_ref32 = P.filter_case(*cfg, device="cpu")
    # sin(10 x) amplifies fp32 rounding: the oracle's own fp32 path is ~1e-5 from fp64; we must be in that class
    assert e_ours <= 4 * e_ref32 + 1e-6, (cfg, e_ours, e_ref32)


@pytest.mark.parametrize("cfg", [(3, 50, 40, 1), (2, 20, 32, 1), (2, 20, 32, 0), (4, 100, 64, 3), (2, 37, 37, 5), (3, 64, 65, 9),
                                 (1, 5000, 4097, 1), (2, 9, 1, 1), (2, 0, 5, 1), (5, 70, 41, 3), (6, 3000, 2049, 1)])
def test_tokenizer_bit_exact(emu_lib, cfg):
    B, maxchars, max_length, flags = cfg
    assert P.tokenizer_case(B, max(maxchars, 1), max_length, flags, device="cpu")


@pytest.mark.parametrize("cfg", [(3, 50, True), (1, 1, False), (4, 4097, True), (2, 16, False), (5, 33, True)])
def test_reverse_complement_bit_exact(emu_lib, golden_dir, cfg):
    import os
    import numpy as np
    B, maxchars, with_apply = cfg
    g = np.load(os.path.join(golden_dir, "revcomp.npz")) if B == 3 else None
    assert P.revcomp_case(B, maxchars, "cpu", seed=maxchars, with_apply=with_apply, golden=g)


@pytest.mark.parametrize("shape,dtype,gsave", [((2, 2, 300), torch.float32, False), ((1, 2, 1001), torch.float32, False),
                                               ((2, 1, 5000), torch.float32, True), ((1, 2, 20000), torch.bfloat16, True),
                                               ((2, 2, 300), torch.bfloat16, False)])
def test_deferred_dx0_in_short_filter_backward(emu_lib, shape, dtype, gsave):
    """hy_conv_bwd_args.defer_dx0 + hy_shortconv_bwd_gate: same gradients as the two-step form."""
    a = P.conv_case(*shape, mode="shortconv", device="cpu", dtype=dtype, gsave=gsave)
    b = P.conv_case(*shape, mode="shortconv", device="cpu", dtype=dtype, gsave=gsave, defer=True)
    tol = P.FP32_TOL if dtype == torch.float32 else 6e-2
    for name in a:
        assert b[name] <= tol and abs(a[name] - b[name]) <= (1e-6 if dtype == torch.float32 else 2e-3), (name, a[name], b[name])


@pytest.mark.parametrize("L,gsave", [(300, False), (5000, False), (5000, True)])
def test_batch_accumulation_into_fewer_slots(emu_lib, L, gsave):
    """B > nslot: batches b, b + nslot, ... accumulate into one spectrum slot (dKacc +=), in every regime."""
    for mode in ("plain", "shortconv"):
        errs = P.conv_case(3, 2, L, mode=mode, device="cpu", gsave=gsave, nslot=2)
        for name, e in errs.items():
            assert e <= P.FP32_TOL, (mode, L, name, e)
    errs = P.conv_case(3, 1, L, mode="plain", device="cpu", gsave=gsave, nslot=1)
    assert max(errs.values()) <= P.FP32_TOL


@pytest.mark.parametrize("cfg", [(4, 300, 2, torch.float32), (4, 5000, 4, torch.float32), (4, 9000, 2, torch.bfloat16)])
def test_channel_slabs_equal_the_full_operator(emu_lib, cfg):
    """N > 1 along channels (SURVEY 8e, B = 1): every rank's slab result is the corresponding rows of the full run."""
    assert P.channel_slab_case(*cfg[:3], "cpu", dtype=cfg[3])


@pytest.mark.parametrize("cfg", [(64, 5, 2, 300, 320), (64, 5, 2, 64, 64), (16, 3, 1, 130, 130), (32, 7, 0, 100, 128)])
def test_filter_saved_trunk_backward_matches_recompute(emu_lib, cfg):
    assert P.filter_trunk_saved_case(*cfg, device="cpu") <= 1e-5


@pytest.mark.parametrize("cfg", [((3, 3, 8000), "shortconv", torch.float32), ((2, 5, 1000), "plain", torch.float32),
                                 ((1, 11, 30000), "shortconv", torch.bfloat16), ((9, 1, 4096), "gated", torch.float32)])
def test_persistent_pipeline_equals_per_phase_launches(emu_lib, cfg):
    """hy_conv_pipe.cuh: ONE persistent launch dealing A / B / C work items over a ring of 4 row buffers must give
    bit-identical results to the three launches per row group (more rows than ring buffers: buffers are reused)."""
    from dna_b200 import kernels as K
    shape, mode, dt = cfg
    emu_lib.hy_debug_set_block.restype = ctypes.c_int
    emu_lib.hy_debug_set_block(256)
    try:
        res, launches = [], []
        for pipe in (1, 0):
            emu_lib.hy_debug_set_conv_pipe(pipe)
            n0 = K.launch_count()
            res.append(P.conv_case(*shape, mode=mode, device="cpu", dtype=dt, gsave=True, seed=3))
            launches.append(K.launch_count() - n0)
    finally:
        emu_lib.hy_debug_set_conv_pipe(0)          # the library default (opt-in: HYENA_B200_CONV_PIPE=1)
        emu_lib.hy_debug_set_block(0)
    assert launches[0] < launches[1], launches          # spectrum / forward / backward / dk: one launch each
    assert res[0] == res[1], (res[0], res[1])
    for name, e in res[0].items():
        assert e <= (5e-5 if dt == torch.float32 else 6e-2), (name, e)


def test_persistent_pipeline_falls_back_for_odd_column_lengths(emu_lib):
    """The pipeline has instances for the power-of-two column lengths only: with it switched on, a 5 * 2^a length takes
    the per-phase launches (same launch count, same numbers)."""
    from dna_b200 import kernels as K
    emu_lib.hy_debug_set_block.restype = ctypes.c_int
    emu_lib.hy_debug_set_block(256)
    try:
        res, launches = [], []
        for pipe in (1, 0):
            emu_lib.hy_debug_set_conv_pipe(pipe)
            n0 = K.launch_count()
            res.append(P.conv_case(2, 2, 5000, mode="shortconv", device="cpu", gsave=True, seed=3))
            launches.append(K.launch_count() - n0)
    finally:
        emu_lib.hy_debug_set_conv_pipe(0)
        emu_lib.hy_debug_set_block(0)
    assert launches[0] == launches[1] and res[0] == res[1], (launches, res)


def test_fetch_intervals_bit_exact(emu_lib, golden_dir):
    import numpy as np
    assert P.fetch_intervals_case(np.load(os.path.join(golden_dir, "ingest.npz")), "cpu") > 20


@pytest.mark.parametrize("i", [0, 1, 2])
def test_bert_mask_bit_exact(emu_lib, golden_dir, i):
    import numpy as np
    assert P.bert_mask_case(np.load(os.path.join(golden_dir, "ingest.npz")), i, "cpu")


@pytest.mark.parametrize("cfg", [((3, 3, 1700), "shortconv", torch.float32), ((2, 5, 2048), "plain", torch.float32),
                                 ((2, 4, 4096), "gated", torch.float32), ((1, 7, 1300), "shortconv", torch.bfloat16)])
def test_single_kernel_regime_backward_with_saved_spectrum(emu_lib, cfg):
    """L <= 4096: the forward keeps the spectrum of g and k_fused_bwdg transforms dy only (one row buffer per sequence)."""
    shape, mode, dt = cfg
    for name, e in P.conv_case(*shape, mode=mode, device="cpu", dtype=dt, gsave=True, seed=4).items():
        assert e <= (5e-5 if dt == torch.float32 else 6e-2), (cfg, name, e)
