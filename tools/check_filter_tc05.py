"""tcgen05 filter forward (hy_filter_tc05.cu) against the FFMA kernel (hy_filter.cu) on the same inputs, and both timed.
usage: python tools/check_filter_tc05.py [L] [D]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dna_b200 import kernels as K, _lib
from dna_b200.hyena import HyenaFilter

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 256
lib = _lib.lib()
torch.manual_seed(0)


def fwd(f, L, tc, save):
    lib.hy_debug_set_filter_tc05(int(tc))
    lins = [m for m in f.implicit_filter if isinstance(m, torch.nn.Linear)]
    n_inner = len(lins) - 2
    w_h = torch.stack([l.weight for l in lins[1:-1]]).detach() if n_inner else None
    b_h = torch.stack([l.bias for l in lins[1:-1]]).detach() if n_inner else None
    return K.filter_fwd(f.pos_emb.z[0], f.pos_emb.t[0], lins[0].weight.detach(), lins[0].bias.detach(), w_h, b_h,
                        lins[-1].weight.detach(), f.implicit_filter[1].freq.detach().reshape(-1),
                        f.modulation.deltas.detach().reshape(-1), 0.05, True, L, save_h=save, save_trunk=save)


for (Lc, Dc, emb, order, ninner) in [(300, 16, 5, 64, 2), (4097, 256, 5, 64, 2), (1000, 70, 3, 16, 1), (129, 600, 5, 64, 0), (L, D, 5, 64, 2)]:
    f = HyenaFilter(Dc, emb_dim=emb, order=order, seq_len=Lc + 2, w=10, lr_pos_emb=0.0, num_inner_mlps=ninner, shift=0.05).cuda()
    for save in (False, True):
        ref = fwd(f, Lc, False, save)
        got = fwd(f, Lc, True, save)
        ref = ref if isinstance(ref, tuple) else (ref,)
        got = got if isinstance(got, tuple) else (got,)
        errs = []
        for r_, g_ in zip(ref, got):
            if r_.dim() == 3:      # a_save [layer][64][lda]: compare the written region
                r_, g_ = r_[:, :order, :Lc], g_[:, :order, :Lc]
            errs.append(((g_ - r_).abs().max() / r_.abs().max()).item())
        print(f"L={Lc} D={Dc} emb={emb} order={order} n_inner={ninner} save={save}: rel diff tc05 vs FFMA (k, h_last, a_save) =",
              ["%.2e" % e for e in errs], flush=True)

f = HyenaFilter(D, emb_dim=5, order=64, seq_len=L + 2, w=10, lr_pos_emb=0.0, shift=0.05).cuda()
for tc in (False, True):
    for save in (False, True):
        for _ in range(2):
            fwd(f, L, tc, save)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fwd(f, L, tc, save)
        e1.record()
        torch.cuda.synchronize()
        print(f"L={L} D={D} {'tcgen05' if tc else 'FFMA   '} save={save}: {e0.elapsed_time(e1) / 5:.3f} ms per forward", flush=True)
lib.hy_debug_set_filter_tc05(1)
