"""Data-parallel step on the GPU with the REAL renderer ("1-GPU vs N-GPU gradient equality on the same global
batch"): GradArena + row exchange, the reference's stock wrappers (``DDP(find_unused_parameters=True)``,
avr_runner_ddp.py:98; ``nn.DataParallel``, avr_runner.py:63) and concurrent calls from several host threads.

The two ranks talk over gloo (host-staged), so the test also runs on a box with ONE GPU (both ranks on cuda:0); with two
or more GPUs each rank takes its own device.  NCCL itself is exercised by ``bench.py --gpus N``.
"""
import copy
import os
import socket
import threading

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import avr_b200
from avr_b200.configs import tiny_config
from avr_b200.ddp import shard_receivers
from oracle import field_ref
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem(embed):
    cfg = tiny_config("AVRModel", n_azi=8, n_ele=4, n_samples=16, T=200)
    if embed:
        cfg["model"]["channel_embed"] = {"is_embed": True, "ch_num": 8, "connection_type": "add", "is_sigma_encoder": True,
                                         "is_sigma_decoder": False, "is_signal_network": True}
    ref = field_ref.trained_like_(field_ref.AVRModelRef(cfg["model"], seed=31), seed=32)
    gen = torch.Generator().manual_seed(9)
    rx = ((torch.rand(4, 3, generator=gen) * 2 - 1) * 2).float()
    tx = ((torch.rand(4, 3, generator=gen) * 2 - 1) * 2).float()
    ch = torch.tensor([5, 0, 3, 5]) if embed else None                 # ranks see DIFFERENT channels (0,2 -> 5,3; 1,3 -> 0,5)
    azi = torch.rand(8, generator=gen)
    G = torch.randn(4, 101, 2, generator=gen)
    return cfg, ref.state_dict(), rx, tx, ch, azi, G


def _render_grads(ren, native, dev, rx, tx, ch, azi, G, idx):
    out = ren(rx[idx].to(dev), tx[idx].to(dev), ch_idx=ch[idx].to(dev) if ch is not None else None, azi_rand=azi)
    (out * G[idx].to(dev)).sum().backward()
    return out


def _worker(rank, world, port, out_dir, embed):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, sd, rx, tx, ch, azi, G = _problem(embed)
    idx = shard_receivers(4, rank, world)
    # (1) GradArena with the row exchange
    native = avr_b200.AVRModel(cfg["model"])
    native.load_state_dict(sd)
    native = native.to(dev)
    ren = avr_b200.AVRRender(native, **cfg["render"])
    arena = avr_b200.GradArena(ren.parameters()).attach(ren)
    assert arena.reduce_numel < arena.numel()
    arena.zero_()
    _render_grads(ren, native, dev, rx, tx, ch, azi, G, idx)
    arena.all_reduce_mean()
    torch.cuda.synchronize()
    torch.save({n: (p.grad * world).cpu() for n, p in native.named_parameters()}, os.path.join(out_dir, f"arena{rank}.pt"))
    # (2) the reference's stock wrapper on an identical replica
    native2 = avr_b200.AVRModel(cfg["model"])
    native2.load_state_dict(sd)
    native2 = native2.to(dev)
    ren2 = avr_b200.AVRRender(native2, **cfg["render"])
    ddp = torch.nn.parallel.DistributedDataParallel(ren2, device_ids=None, find_unused_parameters=True)
    out = ddp(rx[idx].to(dev), tx[idx].to(dev), ch_idx=ch[idx].to(dev) if ch is not None else None, azi_rand=azi)
    (out * G[idx].to(dev)).sum().backward()
    torch.cuda.synchronize()
    torch.save({n: (p.grad * world).cpu() for n, p in native2.named_parameters() if p.grad is not None},
               os.path.join(out_dir, f"ddp{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("embed", [False, True])
def test_two_rank_gradients_equal_single_process(built_library, tmp_path, embed):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), embed), nprocs=world, join=True)
    cfg, sd, rx, tx, ch, azi, G = _problem(embed)
    dev = torch.device("cuda:0")
    native = avr_b200.AVRModel(cfg["model"])
    native.load_state_dict(sd)
    native = native.to(dev)
    ren = avr_b200.AVRRender(native, **cfg["render"])
    _render_grads(ren, native, dev, rx, tx, ch, azi, G, list(range(4)))        # the global batch on one process
    single = {n: p.grad.cpu() for n, p in native.named_parameters()}
    a0, a1 = torch.load(tmp_path / "arena0.pt"), torch.load(tmp_path / "arena1.pt")
    d0, d1 = torch.load(tmp_path / "ddp0.pt"), torch.load(tmp_path / "ddp1.pt")
    assert set(a0) == set(single)
    for name, g in single.items():
        assert torch.equal(a0[name], a1[name]), name                  # replicas hold the bit-identical mean
        assert rel_l2(a0[name], g) < 1e-5, (name, rel_l2(a0[name], g))
        assert torch.equal(d0[name], d1[name]) and rel_l2(d0[name], g) < 1e-5, name
    if embed:                                                          # rows of channels nobody rendered stay zero
        emb = [n for n in single if "embedding" in n]
        assert emb and all(float(a0[n][[1, 2, 4, 6, 7]].abs().max()) == 0 and float(a0[n][[0, 3, 5]].abs().min()) >= 0 for n in emb)


def test_concurrent_calls_from_host_threads(built_library):
    """nn.DataParallel drives forward from one Python thread per replica and autograd runs backward on its own device
    threads: the library keeps no global mutable state, so concurrent calls on separate streams must give exactly what
    sequential calls give."""
    cfg, sd, rx, tx, ch, azi, G = _problem(False)
    dev = torch.device("cuda:0")
    reps = []
    for _ in range(3):
        n = avr_b200.AVRModel(cfg["model"])
        n.load_state_dict(sd)
        n = n.to(dev)
        reps.append((n, avr_b200.AVRRender(n, **cfg["render"])))
    # sequential reference
    n0, r0 = reps[0]
    _render_grads(r0, n0, dev, rx, tx, None, azi, G, [0, 1, 2, 3])
    want = [p.grad.clone() for p in n0.parameters()]
    errors = []

    def run(k):
        try:
            with torch.cuda.stream(torch.cuda.Stream(dev)):
                for _ in range(3):
                    reps[k][0].zero_grad(set_to_none=True)
                    _render_grads(reps[k][1], reps[k][0], dev, rx, tx, None, azi, G, [0, 1, 2, 3])
                torch.cuda.current_stream().synchronize()
        except Exception as exc:                                       # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=run, args=(k,)) for k in (1, 2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    torch.cuda.synchronize()
    for k in (1, 2):
        for p, w in zip(reps[k][0].parameters(), want):
            assert torch.equal(p.grad, w)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="nn.DataParallel needs two GPUs to replicate")
def test_stock_data_parallel_wrapper(built_library):
    """avr_runner.py:63: ``nn.DataParallel(renderer)`` -- the batch is split over the visible GPUs by torch."""
    cfg, sd, rx, tx, ch, azi, G = _problem(False)
    dev = torch.device("cuda:0")
    native = avr_b200.AVRModel(cfg["model"])
    native.load_state_dict(sd)
    native = native.to(dev)
    ren = avr_b200.AVRRender(native, **cfg["render"])
    out = ren(rx.to(dev), tx.to(dev), azi_rand=azi)
    (out * G.to(dev)).sum().backward()
    want_out, want = out.detach().clone(), [p.grad.clone() for p in native.parameters()]
    native.zero_grad(set_to_none=True)
    dp = torch.nn.DataParallel(copy.copy(ren), device_ids=[0, 1])
    out = dp(rx.to(dev), tx.to(dev), azi_rand=azi.tolist())            # a tensor kwarg would be split across replicas
    assert rel_l2(out, want_out) < 1e-6
    (out * G.to(dev)).sum().backward()
    for p, w in zip(native.parameters(), want):
        assert rel_l2(p.grad, w) < 1e-5
