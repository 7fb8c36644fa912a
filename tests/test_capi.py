"""The C-ABI boundary (include/hyena_b200.h): the nvcc-built library loads without a GPU and exports
every declared symbol; the ctypes table matches the header; the product loader refuses to run on CPU."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "hyena_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hy_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(g.LIB)
    return g.LIB


def test_header_and_binding_agree():
    from dna_b200 import _lib
    assert header_functions() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(built_lib):
    import ctypes
    lib = ctypes.CDLL(built_lib)
    for name in header_functions():
        assert hasattr(lib, name), name
    lib.hy_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.hy_version()
    lib.hy_fft_len.restype = ctypes.c_int
    assert lib.hy_fft_len(1) == 256 and lib.hy_fft_len(1024) == 1024 and lib.hy_fft_len(1025) == 2048
    assert lib.hy_fft_len(1_000_000) == 1 << 20 and lib.hy_fft_len(160_000) == 40 * 4096
    assert lib.hy_fft_len(3_000_000) == -1
    lib.hy_conv_workspace_bytes.restype = ctypes.c_size_t
    assert lib.hy_conv_workspace_bytes(8, 256, 1024, 1) == 0                      # fused regime: no scratch
    assert lib.hy_conv_workspace_bytes(1, 256, 1_000_000, 1) == 256 * 8 * (1 << 20)  # 256 rows of 8 MB fit the 2 GB scratch budget
    # the SASS really is sm_100a
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_loader_has_no_cpu_fallback(built_lib):
    """Without a CUDA device every product entry point raises (run in a clean interpreter)."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import torch\n"
        "from dna_b200 import _lib\n"
        "from dna_b200.fftconv import fftconv_func\n"
        "from dna_b200.hyena import HyenaOperator\n"
        "if torch.cuda.is_available(): print('SKIP'); raise SystemExit(0)\n"
        "n = 0\n"
        "try: fftconv_func(torch.randn(1, 2, 64), torch.randn(2, 64), torch.randn(2), None, False)\n"
        "except _lib.HyenaB200Error: n += 1\n"
        "try: HyenaOperator(d_model=8, l_max=64)(torch.randn(1, 32, 8))\n"
        "except _lib.HyenaB200Error: n += 1\n"
        "print('RAISED', n)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "SKIP" in r.stdout or "RAISED 2" in r.stdout, r.stdout + r.stderr[-2000:]


def test_oracle_is_not_imported_by_the_package():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dna_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle timed on the host cores) runs without a GPU and prints ONE JSON line
    carrying the contract keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-len", "512"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "nt/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"] == "hyenadna-large-1m" and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "nt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
