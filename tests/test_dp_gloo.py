"""N > 1 host logic on CPU: world_size-2 gloo run of the flat gradient all-reduce + batch sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dna_b200.dp import FlatGradAllReduce, shard_batch
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    red = FlatGradAllReduce(model.parameters())
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    lo, hi = shard_batch(8, rank, world)
    red.zero()
    model(x[lo:hi]).pow(2).sum().backward()
    red.allreduce(average=False)
    q.put((rank, red.flat.clone(), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_allreduce_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    model(x).pow(2).sum().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    shards = sorted(g[2] for g in got)
    assert shards == [(0, 4), (4, 8)]
    for _, flat, _ in got:
        assert torch.allclose(flat, ref, atol=1e-5)


def test_shard_helpers():
    from dna_b200.dp import channel_slab, shard_batch
    assert [shard_batch(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [channel_slab(256, r, 8) for r in (0, 7)] == [(0, 32), (224, 256)]
    with pytest.raises(ValueError):
        channel_slab(250, 0, 8)
