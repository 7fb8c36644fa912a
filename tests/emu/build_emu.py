"""TEST INFRASTRUCTURE: build libhyena_b200_emu.so = the kernel sources of dna_b200/csrc compiled by
g++ against tests/emu/cuda_emu.h (a CPU execution-model emulation).  Used only by the "not gpu"
tests to exercise kernel logic in a container without a GPU.  Never loaded by the package."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "dna_b200", "csrc")
OUT_DIR = os.path.join(ROOT, "build", "emu")
LIB = os.path.join(OUT_DIR, "libhyena_b200_emu.so")
SOURCES = ["hy_core.cu", "hy_conv_api.cu", "hy_conv_f32.cu", "hy_conv_bf16.cu", "hy_conv_odd5_f32.cu", "hy_conv_odd5_bf16.cu", "hy_conv_odd3_f32.cu", "hy_conv_odd3_bf16.cu", "hy_conv_pipe_f32.cu", "hy_conv_pipe_bf16.cu", "hy_shortconv.cu",
           "hy_filter.cu", "hy_filter_tc.cu", "hy_filter_tc05.cu", "hy_tokenizer.cu", "hy_addln.cu", "hy_exchange.cu"]
FLAGS = ["-O1", "-std=c++20", "-fopenmp", "-fPIC", "-DHY_EMU_BUILD", "-I" + os.path.join(ROOT, "tests", "emu"),
         "-I" + CSRC, "-Wno-unused-but-set-variable"]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(ROOT, "tests", "emu", "cuda_emu.h"), os.path.join(ROOT, "include", "hyena_b200.h")]
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OUT_DIR, s.replace(".cu", ".o"))
        if not _newer(obj, [src] + headers):
            jobs.append(["g++"] + FLAGS + ["-x", "c++", "-c", src, "-o", obj])
    emu_src = os.path.join(ROOT, "tests", "emu", "cuda_emu.cpp")
    emu_obj = os.path.join(OUT_DIR, "cuda_emu.o")
    if not _newer(emu_obj, [emu_src] + headers):
        jobs.append(["g++"] + FLAGS + ["-c", emu_src, "-o", emu_obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emu build failed: %s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
        if verbose and r.stderr:
            print(r.stderr[-2000:])

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OUT_DIR, s.replace(".cu", ".o")) for s in SOURCES] + [emu_obj]
    if jobs or not os.path.exists(LIB):
        run(["g++", "-shared", "-fopenmp", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
