"""world_size-2 gloo tests of the data-parallel step (CPU): receiver sharding, flat-arena gradient mean, the row
all-gather that replaces the all-reduce of the per-ray / per-receiver tables, and the arena layout behind it."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch.utils.data import DistributedSampler

import avr_b200
from avr_b200.ddp import RowExchange, shard_receivers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                     # identical replicas
    lin = torch.nn.Linear(6, 4, bias=False)
    arena = avr_b200.GradArena(lin.parameters())
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10      # 8 "receivers"
    mine = shard_receivers(8, rank, world)
    arena.zero_()
    lin(x[mine]).square().mean().backward()                  # per-rank mean over its local receivers
    arena.all_reduce_mean()
    torch.save(arena.flat.clone(), os.path.join(out_dir, f"g{rank}.pt"))
    # the row exchange: every rank contributes n rows, everybody gets all of them in rank order, scaled to the mean
    ex = RowExchange()
    u = torch.full((3, 3), float(rank))
    rows = torch.arange(3 * 5, dtype=torch.float32).view(3, 5) + 100 * rank
    u2, rows2 = torch.full((2, 3), 10.0 + rank), torch.full((2, 5), 7.0 * (rank + 1))       # a second table, other row count
    (u_all, rows_all), (u2_all, rows2_all) = ex.gather([(u, rows), (u2, rows2)])
    assert u_all.shape == (3 * world, 3) and rows_all.shape == (3 * world, 5) and u2_all.shape == (2 * world, 3)
    for r in range(world):
        assert torch.equal(u_all[3 * r:3 * r + 3], torch.full((3, 3), float(r)))
        assert torch.equal(rows_all[3 * r:3 * r + 3], (torch.arange(15, dtype=torch.float32).view(3, 5) + 100 * r) / world)
        assert torch.equal(u2_all[2 * r:2 * r + 2], torch.full((2, 3), 10.0 + r))
        assert torch.equal(rows2_all[2 * r:2 * r + 2], torch.full((2, 5), 7.0 * (r + 1) / world))
    # an arena whose tail is exchanged as rows: the all-reduce must leave that tail alone
    a, b = torch.nn.Parameter(torch.zeros(10)), torch.nn.Parameter(torch.zeros(7))
    arena2 = avr_b200.GradArena([a, b])
    arena2._layout([a, b], n_reduced=1)
    arena2.flat.fill_(float(rank + 1))
    arena2.all_reduce_mean()
    assert arena2.reduce_numel == 12
    assert torch.equal(arena2.flat[:12], torch.full((12,), (1 + world) / 2)) and torch.equal(arena2.flat[12:], torch.full((8,), float(rank + 1)))
    assert a.grad.data_ptr() == arena2.flat.data_ptr() and b.grad.data_ptr() == arena2.flat[12:].data_ptr()
    dist.destroy_process_group()


def test_two_rank_gradient_mean_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    assert torch.equal(g0, g1)
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 4, bias=False)
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10
    lin(x).square().mean().backward()                        # global batch on one process
    assert torch.allclose(g0[: lin.weight.numel()].view_as(lin.weight), lin.weight.grad, rtol=1e-5, atol=1e-6)


def test_row_exchange_without_process_group_is_the_identity():
    ex = RowExchange()
    u, rows = torch.rand(4, 3), torch.rand(4, 6)
    ((u_all, rows_all),) = ex.gather([(u, rows)])
    assert torch.equal(u_all, u) and torch.equal(rows_all, rows) and ex.gather([]) == []


def test_shard_receivers_is_distributed_sampler_order():
    """avr_runner_ddp.py:131-137: DistributedSampler(shuffle=False) -- strided, padded by wrapping around so that every
    rank runs the same number of steps (a ragged last step would hang the per-step all-reduce)."""
    for n, w in ((8, 2), (9, 4), (3, 8), (3969, 8), (1, 4)):
        shards = [shard_receivers(n, r, w) for r in range(w)]
        assert len({len(s) for s in shards}) == 1
        assert sorted(set(i for s in shards for i in s)) == list(range(n))
        for r in range(w):
            assert shards[r] == list(DistributedSampler(range(n), num_replicas=w, rank=r, shuffle=False))
            assert shard_receivers(n, r, w, drop_last=True) == list(
                DistributedSampler(range(n), num_replicas=w, rank=r, shuffle=False, drop_last=True))
    assert shard_receivers(0, 0, 2) == []
