"""Kernel-level breakdown of one bench step (torch.profiler) + CPU enqueue vs GPU time.
usage: python tools/prof_step.py [workload] [out.txt]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import bench
from dna_b200.standalone import HyenaDNAModel
from dna_b200.tokenizer import CharacterTokenizer
from dna_b200.dp import FlatGradAllReduce

wl = sys.argv[1] if len(sys.argv) > 1 else "hyenadna-large-1m"
cfg = bench.WORKLOADS[wl]
B, L = cfg["batch"], cfg["seqlen"]
dev = torch.device("cuda")
torch.manual_seed(2222)
model = HyenaDNAModel(d_model=cfg["d_model"], n_layer=cfg["n_layer"], d_inner=cfg["d_inner"], vocab_size=12, pad_vocab_size_multiple=8,
                      embed_dropout=0.0, lm_head=True, layer=dict(l_max=L + 2, **bench.LAYER_CFG)).to(dev).train()
red = FlatGradAllReduce(model.parameters())
opt = torch.optim.AdamW(model.parameters(), lr=6e-4, fused=True)
tok = CharacterTokenizer(["A", "C", "G", "T", "N"], model_max_length=L + 1)
host = torch.from_numpy(bench.synth_bytes(B, L, 0)).pin_memory()
dbytes = host.to(dev)

def step(src):
    ids = tok.encode_bytes_cuda(src, None, L + 1, add_special_tokens=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(ids[:, :-1])
    loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), ids[:, 1:].reshape(-1))
    red.zero(); loss.backward(); red.allreduce(); opt.step()
    return loss

for _ in range(3):
    step(dbytes)
torch.cuda.synchronize()
# CPU enqueue time vs GPU time
for mode in ("nosync", "sync"):
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        l = step(dbytes if mode == "nosync" else host.to(dev, non_blocking=True))
        t1 = time.perf_counter()
        if mode == "sync":
            l.item()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
    print(mode, "cpu-enqueue ms / total ms:", [(round(a, 1), round(b, 1)) for a, b in ts], flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(dbytes)
    torch.cuda.synchronize()
txt = prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=70, max_name_column_width=70)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
