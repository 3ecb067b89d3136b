"""world_size-2 gloo test of the data-parallel step: receiver sharding + flat-arena gradient mean."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import avr_b200
from avr_b200.ddp import shard_receivers


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
    local = arena.flat.clone()
    arena.all_reduce_mean()
    torch.save(arena.flat.clone(), os.path.join(out_dir, f"g{rank}.pt"))
    # the overlapped exchange (GradArena.attach): the renderer announces each gradient from inside its backward pass
    class _Renderer:
        grad_ready_hook = grad_done_hook = None
    ren = _Renderer()
    arena.attach(ren)
    announced = local.clone()
    ren.grad_ready_hook(lin.weight, announced)
    ren.grad_done_hook()
    torch.save(announced, os.path.join(out_dir, f"h{rank}.pt"))
    avr_b200.GradArena.detach(ren)
    assert ren.grad_ready_hook is None
    dist.destroy_process_group()


def test_two_rank_gradient_mean_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    assert torch.equal(g0, g1)
    assert torch.equal(torch.load(tmp_path / "h0.pt"), g0) and torch.equal(torch.load(tmp_path / "h1.pt"), g0)
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 4, bias=False)
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10
    lin(x).square().mean().backward()                        # global batch on one process
    assert torch.allclose(g0[: lin.weight.numel()].view_as(lin.weight), lin.weight.grad, rtol=1e-5, atol=1e-6)


def test_shard_receivers_partition():
    for n, w in ((8, 2), (9, 4), (3, 8)):
        seen = sorted(i for r in range(w) for i in shard_receivers(n, r, w))
        assert seen == list(range(n))
