"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: sharding covers the batch exactly
once, the flat-gradient all-reduce reproduces the single-process gradient, clipping is consistent."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sparf_b200.distributed import FlatGradients, global_mean_scale, shard_range, shard_rays


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions_exactly():
    for n in (1, 7, 341, 1023, 4096):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    fg = FlatGradients([net])
    g = torch.Generator().manual_seed(1)
    x = torch.randn(103, 5, generator=g)
    y = torch.randn(103, 3, generator=g)
    idx = torch.randperm(103, generator=g)          # same permutation on every rank
    mine = shard_rays(idx, rank, world)
    fg.zero_()
    loss = torch.nn.functional.huber_loss(net(x[mine]), y[mine], delta=0.5) * 2 * global_mean_scale(len(mine), 103)
    loss.backward()
    fg.all_reduce()
    total = fg.clip_grad_norm_(0.1)
    if rank == 0:
        out.put((fg.flat.clone(), float(total), [int(i) for i in mine]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat2, total2, mine0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-process reference over the full batch
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    fg = FlatGradients([net])
    g = torch.Generator().manual_seed(1)
    x = torch.randn(103, 5, generator=g)
    y = torch.randn(103, 3, generator=g)
    (torch.nn.functional.huber_loss(net(x), y, delta=0.5) * 2).backward()
    total1 = fg.clip_grad_norm_(0.1)
    assert len(mine0) == 52
    assert abs(total1.item() - total2) < 1e-5 * max(1.0, total2)
    assert torch.allclose(fg.flat, flat2, rtol=1e-4, atol=1e-7)
