"""Tokenizer kernel against plain streams of the same size: B x 2^20 nucleotides (1 B in, 8 B of int64 id out each).
B must be large enough that the kernel outlasts the ~0.15 ms of host work per call (ctypes + torch.empty).
usage: python tools/bench_tokenizer.py [B]"""
import sys, os, torch
sys.path.insert(0, "/root/repo")
from dna_b200.tokenizer import CharacterTokenizer
dev = "cuda"
tokB, tokL = (int(sys.argv[1]) if len(sys.argv) > 1 else 512), 1 << 20
tb = torch.randint(65, 85, (tokB, tokL), dtype=torch.uint8, device=dev)
tokz = CharacterTokenizer(["A", "C", "G", "T", "N"], model_max_length=tokL + 1)
def t(f, n=20):
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: tokz.encode_bytes_cuda(tb, None, tokL + 1, add_special_tokens=True))
print("tokenize %.4f ms  %.0f GB/s" % (ms, 9.0 * tokB * tokL / ms / 1e6))
out = torch.empty(tokB, tokL + 1, dtype=torch.int64, device=dev)
ms = t(lambda: out.fill_(7))
print("fill int64 %.4f ms  %.0f GB/s (write only)" % (ms, 8.0 * tokB * tokL / ms / 1e6))
src = torch.empty(tokB, tokL + 1, dtype=torch.int64, device=dev)
ms = t(lambda: out.copy_(src))
print("copy int64 %.4f ms  %.0f GB/s (r+w)" % (ms, 16.0 * tokB * tokL / ms / 1e6))
ms = t(lambda: tb.to(torch.int64))
print("aten u8->i64 %.4f ms  %.0f GB/s" % (ms, 9.0 * tokB * tokL / ms / 1e6))
