"""ctypes binding of libhyena_b200.so — the C-ABI declared in include/hyena_b200.h.

The product path loads ONLY the nvcc-built sm_100a library that lives next to this file
(``dna_b200/lib/libhyena_b200.so``) and fails loudly if it is missing or if no CUDA device is
usable: there is no CPU fallback (north_star).  The ``_use_library_for_tests`` hook exists so the
"not gpu" test-suite can point the same binding at a CPU execution-model emulation of the kernel
source (tests/emu); nothing in the package ever calls it.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# HYENA_B200_LIB: developer hook to A/B-test another sm_100a build of the same sources (never a CPU build)
LIB_PATH = os.environ.get("HYENA_B200_LIB") or os.path.join(_HERE, "lib", "libhyena_b200.so")

HY_F32, HY_BF16 = 0, 1
IN_PLAIN, IN_PREGATE, IN_SHORTCONV = 0, 1, 2
OUT_PLAIN, OUT_POSTGATE, OUT_SHORTCONV = 0, 1, 2
TOK_ADD_SEP, TOK_ADD_CLS, TOK_N_TO_PAD, TOK_NUC_ENCODE = 1, 2, 4, 8

_lock = threading.Lock()
_lib = None
_is_emulation = False


class HyenaB200Error(RuntimeError):
    pass


class ConvFwdArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("B", C.c_int), ("H", C.c_int), ("L", C.c_int),
        ("in_mode", C.c_int), ("out_mode", C.c_int),
        ("u", C.c_void_p), ("pre", C.c_void_p), ("u_bs", C.c_longlong), ("ldu", C.c_int),
        ("post", C.c_void_p), ("post_bs", C.c_longlong), ("ldpost", C.c_int),
        ("sw", C.c_void_p), ("sb", C.c_void_p), ("pb", C.c_void_p),
        ("Kf", C.c_void_p),
        ("out", C.c_void_p), ("ysave", C.c_void_p), ("out_bs", C.c_longlong), ("ldo", C.c_int),
        ("ws", C.c_void_p), ("ws_bytes", C.c_size_t),
        ("gsave", C.c_void_p),
    ]


class ConvBwdArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("B", C.c_int), ("H", C.c_int), ("L", C.c_int),
        ("in_mode", C.c_int), ("out_mode", C.c_int),
        ("u", C.c_void_p), ("pre", C.c_void_p), ("u_bs", C.c_longlong), ("ldu", C.c_int),
        ("post", C.c_void_p), ("post_bs", C.c_longlong), ("ldpost", C.c_int),
        ("sw", C.c_void_p), ("sb", C.c_void_p), ("pb", C.c_void_p),
        ("Kf", C.c_void_p),
        ("dout", C.c_void_p), ("out_bs", C.c_longlong), ("ldo", C.c_int),
        ("ysave", C.c_void_p), ("ys_bs", C.c_longlong), ("ldys", C.c_int),
        ("du", C.c_void_p), ("dpre", C.c_void_p), ("dpost", C.c_void_p),
        ("dKacc", C.c_void_p), ("nslot", C.c_int),
        ("dDpart", C.c_void_p),
        ("ws", C.c_void_p), ("ws_bytes", C.c_size_t),
        ("gsave", C.c_void_p),
        ("defer_dx0", C.c_int),
    ]


class PeerPullArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p * 8), ("flags", C.c_void_p * 8),
        ("G", C.c_int), ("self", C.c_int), ("epoch", C.c_uint), ("reduce", C.c_int),
        ("n_outer", C.c_int), ("n_inner", C.c_int), ("row_bytes", C.c_longlong),
        ("src_base", C.c_longlong), ("src_outer", C.c_longlong), ("src_inner", C.c_longlong),
        ("dst_outer", C.c_longlong), ("dst_inner", C.c_longlong), ("dst_peer", C.c_longlong),
        ("dst", C.c_void_p), ("max_ctas", C.c_int),
    ]


class FilterArgs(C.Structure):
    _fields_ = [
        ("L", C.c_int), ("D", C.c_int), ("order", C.c_int), ("emb_dim", C.c_int), ("n_inner", C.c_int),
        ("z", C.c_void_p), ("ldz", C.c_int),
        ("t", C.c_void_p),
        ("w_in", C.c_void_p), ("b_in", C.c_void_p),
        ("w_h", C.c_void_p), ("b_h", C.c_void_p),
        ("w_out", C.c_void_p),
        ("freq", C.c_void_p),
        ("deltas", C.c_void_p),
        ("shift", C.c_float), ("modulate", C.c_int),
    ]


# every symbol include/hyena_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "hy_init": (C.c_int, []),
    "hy_last_error": (C.c_char_p, []),
    "hy_version": (C.c_char_p, []),
    "hy_fft_len": (C.c_int, [C.c_int]),
    "hy_conv_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "hy_conv_gsave_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hy_conv_ndpart": (C.c_int, [C.c_int]),
    "hy_launch_count": (C.c_ulonglong, []),
    "hy_clock_probe": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hy_set_scratch_budget": (C.c_int, [C.c_size_t]),
    "hy_set_pipeline": (C.c_int, [C.c_int, C.c_size_t]),
    "hy_filter_spectrum": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "hy_conv_fwd": (C.c_int, [C.POINTER(ConvFwdArgs), C.c_void_p]),
    "hy_conv_bwd": (C.c_int, [C.POINTER(ConvBwdArgs), C.c_void_p]),
    "hy_conv_dk": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                             C.c_void_p, C.c_size_t, C.c_void_p]),
    "hy_shortconv_nchunk": (C.c_int, [C.c_int, C.c_int]),
    "hy_shortconv_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hy_shortconv_bwd_gate": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "hy_shortconv_fwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hy_filter_fwd": (C.c_int, [C.POINTER(FilterArgs), C.c_void_p, C.c_int, C.c_void_p]),
    "hy_filter_modulate_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int,
                                         C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hy_filter_trunk_bwd_layout": (C.c_int, [C.POINTER(FilterArgs), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hy_filter_fwd_save": (C.c_int, [C.POINTER(FilterArgs), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "hy_filter_trunk_bwd": (C.c_int, [C.POINTER(FilterArgs), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "hy_filter_trunk_save_layout": (C.c_int, [C.POINTER(FilterArgs), C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "hy_filter_fwd_save_trunk": (C.c_int, [C.POINTER(FilterArgs), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                           C.c_int, C.c_void_p]),
    "hy_filter_trunk_bwd_saved": (C.c_int, [C.POINTER(FilterArgs), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_void_p]),
    "hy_filter_out_bwd_supported": (C.c_int, [C.c_int, C.c_int]),
    "hy_filter_out_bwd_workspace_bytes": (C.c_size_t, [C.c_int]),
    "hy_filter_out_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "hy_tokenize": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p,
                              C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hy_reverse_complement": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                                        C.c_int, C.c_int, C.c_void_p]),
    "hy_add_ln_supported": (C.c_int, [C.c_int]),
    "hy_add_ln_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int,
                                C.c_void_p]),
    "hy_add_ln_bwd_parts": (C.c_int, [C.c_longlong, C.c_int]),
    "hy_add_ln_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_longlong, C.c_int, C.c_void_p]),
    "hy_add_ln_dropout_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int,
                                        C.c_void_p]),
    "hy_add_ln_dropout_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "hy_fetch_intervals": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p]),
    "hy_bert_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong,
                               C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hy_peer_flag_bytes": (C.c_size_t, []),
    "hy_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "hy_peer_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "hy_peer_close": (C.c_int, [C.c_void_p]),
    "hy_peer_free": (C.c_int, [C.c_void_p]),
    "hy_peer_pull": (C.c_int, [C.POINTER(PeerPullArgs), C.c_void_p]),
    "hy_peer_wait_done": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p]),
    "hy_peer_error": (C.c_int, [C.c_void_p]),
}


def _bind(path: str):
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def load_library():
    """Load (once) the sm_100a library. Raises HyenaB200Error when it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise HyenaB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). hyena-b200 has no CPU fallback.")
        _lib = _bind(LIB_PATH)
        return _lib


_inited_devices = set()
_cuda_ok = False


def lib():
    """The bound library, initialised (once per device: hy_init queries device properties, ~10 ms) on the
    current CUDA device."""
    global _cuda_ok
    l = _lib if _lib is not None else load_library()
    if _is_emulation:
        dev = -1
    else:
        import torch
        if not _cuda_ok:                       # is_available() costs ~0.5 ms (NVML): ask once
            if not torch.cuda.is_available():
                raise HyenaB200Error("hyena-b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            _cuda_ok = True
        dev = torch.cuda.current_device()
    if dev not in _inited_devices:
        check(l.hy_init(), l)
        _inited_devices.add(dev)
    return l


def check(rc: int, l=None):
    if rc != 0:
        l = l or _lib
        msg = l.hy_last_error().decode("utf-8", "replace") if l is not None else ""
        raise HyenaB200Error(f"libhyena_b200 error {rc}: {msg}")


def is_emulation() -> bool:
    return _is_emulation


def _use_library_for_tests(path: str):
    """TEST HOOK: bind a different build of the same C-ABI (the CPU emulation of the kernel source
    under tests/emu). Never called by package code."""
    global _lib, _is_emulation
    with _lock:
        _lib = _bind(path)
        _is_emulation = True
        _inited_devices.clear()
        return _lib


def current_stream_ptr():
    if _is_emulation:
        return None
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
