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


def test_flat_gradients_detect_detached_grads():
    """optimizer.zero_grad() (set_to_none=True) detaches every .grad from the flat buffer: the next collective must
    re-attach (None) or refuse (foreign storage) instead of silently reducing stale zeros; `progress` is left out."""
    import pytest

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(4, 3)
            self.progress = torch.nn.Parameter(torch.tensor(0.0))

    net = Net()
    fg = FlatGradients([net])
    assert fg.flat.numel() == 4 * 3 + 3 and all(p is not net.progress for p in fg.params)
    assert all(getattr(p, "_sparf_inplace_grad", False) for p in fg.params) and not hasattr(net.progress, "_sparf_inplace_grad")
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    net.lin(torch.ones(2, 4)).sum().backward()
    assert fg.flat.abs().sum() > 0
    opt.zero_grad()                                   # set_to_none=True: .grad = None
    assert net.lin.weight.grad is None
    fg.zero_()                                        # re-attaches
    assert net.lin.weight.grad.data_ptr() == fg.flat.data_ptr()
    net.lin(torch.ones(2, 4)).sum().backward()
    assert fg.flat.abs().sum() > 0
    net.lin.weight.grad = torch.zeros_like(net.lin.weight)     # foreign storage
    with pytest.raises(RuntimeError, match="no longer aliases"):
        fg.all_reduce()
