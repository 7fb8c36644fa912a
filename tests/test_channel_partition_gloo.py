"""Channel partition of ONE sequence over the ranks (BASELINE.json configs[3]) on CPU:
world_size-2 gloo run of the model-level split — sequence chunks outside the operator core, channel slabs inside it,
two all-to-alls per layer — against the single-process step.  The kernels run under the CPU emulator (tests/emu)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = dict(d_model=8, n_layer=2, d_inner=16, L=96, B=2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(emu_path):
    from dna_b200 import _lib
    _lib._use_library_for_tests(emu_path)
    from dna_b200.standalone import HyenaDNAModel
    torch.manual_seed(7)
    m = HyenaDNAModel(d_model=CFG["d_model"], n_layer=CFG["n_layer"], d_inner=CFG["d_inner"], vocab_size=12,
                      pad_vocab_size_multiple=8, embed_dropout=0.0, lm_head=True,
                      layer=dict(l_max=CFG["L"] + 2, emb_dim=5, filter_order=16, w=4, lr_pos_emb=0.0))
    # the init draws biases of zero: give every per-channel parameter a distinct value so a wrong slab shows
    g = torch.Generator().manual_seed(8)
    for n, p in m.named_parameters():
        if n.endswith("bias"):
            p.data.copy_(torch.randn(p.shape, generator=g) * 0.3)
    m.train()
    ids = torch.randint(7, 11, (CFG["B"], CFG["L"] + 1), generator=torch.Generator().manual_seed(9))
    return m, ids[:, :-1], ids[:, 1:]


def _loss(model, data, target, denom):
    logits = model(data)
    return torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), target.reshape(-1), reduction="sum") / denom


def _worker(rank, world, port, emu_path, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dna_b200.dp import ChannelPartition, FlatGradAllReduce, set_channel_partition
    model, data, target = _build(emu_path)
    part = ChannelPartition()
    assert set_channel_partition(model, part) == CFG["n_layer"]
    red = FlatGradAllReduce(model.parameters())
    lo, hi = part.chunk(CFG["L"])
    red.zero()
    loss = _loss(model, data[:, lo:hi], target[:, lo:hi], CFG["B"] * CFG["L"])
    loss.backward()
    red.allreduce(average=False)
    t = loss.detach().clone()
    dist.all_reduce(t)
    q.put((rank, float(t), red.flat.clone(), part.bytes_sent))
    dist.barrier()
    dist.destroy_process_group()


def test_channel_partition_step_matches_single_process(emu_lib):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    emu_path = build_emu.build()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, emu_path, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    model, data, target = _build(emu_path)
    loss = _loss(model, data, target, CFG["B"] * CFG["L"])
    loss.backward()
    seen, ref = set(), []
    for p in model.parameters():
        if id(p) not in seen:
            seen.add(id(p))
            ref.append(p.grad.reshape(-1))
    ref = torch.cat(ref)
    for rank, l, flat, sent in got:
        assert abs(l - float(loss.detach())) <= 1e-5 * abs(float(loss.detach())), (l, float(loss.detach()))
        err = (flat - ref).abs().max() / ref.abs().max()
        assert err <= 2e-5, (rank, float(err))
        assert sent > 0
