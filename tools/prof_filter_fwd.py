"""Forward-only loop of the implicit-filter kernel for ncu.  usage: python tools/prof_filter_fwd.py L D iters [save]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dna_b200.hyena import HyenaFilter
L = int(sys.argv[1]); D = int(sys.argv[2]); iters = int(sys.argv[3]); save = len(sys.argv) > 4 and sys.argv[4] == "save"
torch.manual_seed(0)
f = HyenaFilter(D, emb_dim=5, order=64, seq_len=L + 2, w=10, lr_pos_emb=0.0, shift=0.05).cuda()
for _ in range(iters):
    if save:
        k = f.filter_cm(L)
    else:
        with torch.no_grad():
            k = f.filter_cm(L)
torch.cuda.synchronize()
print("ok", float(k[0, 0]))
