"""Data-parallel gradient exchange: one flat fp32 arena, one all-reduce per step (+ a row all-gather).

Replaces the DDP reducer of ``avr_runner_ddp.py:98,257`` (25 MB buckets + a used-parameter bitmap
all-reduce per step because of ``find_unused_parameters=True``).  Receivers are independent, every rank
holds a full replica, and the only exchange is the mean of the parameter gradients: all
``.grad`` tensors are views into one contiguous buffer, so the exchange is a single NCCL all-reduce
(NVLS in-switch reduction on NVSwitch) issued on the compute stream right after the backward kernels
(``all_reduce_mean``).

``attach(renderer)`` cuts that all-reduce to the gradients that are dense.  The per-ray and per-receiver hash
tables (``_dir_encoding``, ``_tx_encoding``: 76 of the 120 MB at simu, 183 of 229 MB at MeshRIR, 152 of 236 MB at RAF)
are touched by R = 650..3202 resp. bs points per step, so > 93 % of what the flat all-reduce ships for them is zeros.
With ``attach`` the backward pass all-gathers the pre-scatter rows instead (``[R, 3 + 40]`` floats per rank and table,
~350 KB; one collective at the end of the pass) and every rank scatters the rows of ALL ranks into its
own table gradient with the int64 fixed-point accumulator -- integer sums, so every replica holds the bit-identical
mean, exactly what an all-reduce guarantees.  Those tables then sit behind ``reduce_numel`` in the arena and the
all-reduce covers the dense prefix only (the per-point tables, the MLPs, channel embeddings).

The reference's stock ``DDP(...)`` wrapper also works on ``avr_b200.AVRRender`` (an ordinary ``nn.Module``,
``tests/test_ddp_gloo.py``); this arena is the B200-first path used by ``bench.py``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _active(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


class RowExchange:
    """All-gather of the ``(unit-cube input, gradient row)`` pairs of the per-ray / per-receiver encodings: ONE collective
    for all such tables of a backward pass, issued at its end.  (Starting it in the middle of the pass was measured
    slower: a collective that becomes resident next to the persistent one-CTA-per-SM GEMM kernels takes an SM away from
    the next of them, whose 148th CTA then runs after the other 147.)"""

    def __init__(self, group=None):
        self.group = group

    def gather(self, blocks):
        """``blocks``: list of ``(u[n_i,3], rows[n_i,w])`` of this rank (same shapes on every rank) ->
        list of ``(u_all[W*n_i,3], rows_all[W*n_i,w] / W)`` in rank order."""
        if not blocks:
            return []
        w = blocks[0][1].shape[1]
        if any(r.shape[1] != w for _, r in blocks):
            raise ValueError("row blocks of one exchange must share their width")
        packed = torch.cat([torch.cat([u, r], dim=1) for u, r in blocks], dim=0).contiguous()        # [sum n_i, 3 + w]
        world, dev = 1, packed.device
        if _active(self.group):
            world = dist.get_world_size(self.group)
            staged = packed
            if packed.is_cuda and dist.get_backend(self.group) != "nccl":     # gloo (tests): gather through host copies
                staged = packed.cpu()
            out = torch.empty(world * staged.shape[0], staged.shape[1], dtype=staged.dtype, device=staged.device)
            dist.all_gather_into_tensor(out, staged, group=self.group)
            packed = out.to(dev)
        packed = packed.view(world, -1, 3 + w)
        res, off = [], 0
        for u, r in blocks:
            n = u.shape[0]
            blk = packed[:, off:off + n, :].reshape(world * n, 3 + w)
            rows_all = blk[:, 3:].contiguous()
            if world > 1:
                rows_all = rows_all * (1.0 / world)                             # gradients are averaged over ranks (DDP)
            res.append((blk[:, :3].contiguous(), rows_all))
            off += n
        return res


class GradArena:
    """Owns a flat gradient buffer and points every parameter's ``.grad`` into it."""

    def __init__(self, parameters):
        self.params = [p for p in parameters if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        self._layout(self.params)

    def _layout(self, params, n_reduced=None):
        dev, dtype = params[0].device, params[0].dtype
        self.params, self.offsets, n = list(params), [], 0
        self.reduce_numel = None
        for k, p in enumerate(self.params):
            if p.device != dev or p.dtype != dtype:
                raise ValueError("all parameters must share device and dtype")
            if n_reduced is not None and k == n_reduced:
                self.reduce_numel = n
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4                       # keep every view 16-byte aligned
        if self.reduce_numel is None:
            self.reduce_numel = n
        self.flat = torch.zeros(n, device=dev, dtype=dtype)
        self.bind()

    def bind(self):
        """(Re)attach ``p.grad`` views, e.g. after ``optimizer.zero_grad(set_to_none=True)``."""
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)

    def zero_(self):
        self.flat.zero_()
        self.bind()

    def numel(self) -> int:
        return self.flat.numel()

    def all_reduce_mean(self, group=None, async_op=False):
        """grad <- mean over ranks (DistributedDataParallel semantics); covers ``flat[:reduce_numel]``, the rest was
        exchanged as rows inside the backward pass (``attach``)."""
        if not _active(group):
            return None
        world = dist.get_world_size(group)
        buf = self.flat[:self.reduce_numel]
        if buf.is_cuda and dist.get_backend(group) == "nccl":
            return dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=False)
        buf.div_(world)
        return work

    # -- row exchange for the per-ray / per-receiver tables ----------------------------------------------------------
    def attach(self, renderer, group=None):
        """Exchange the per-ray / per-receiver table gradients as rows (see the module docstring): re-lays the arena out
        with those tables last and installs the exchange on ``renderer`` (tensor-core path, ``avr_b200`` fields).
        Call ``all_reduce_mean`` after backward as before; it now covers the dense prefix only."""
        net = renderer.network_fn
        if getattr(renderer, "dense", "tc") != "tc" or not hasattr(net, "small_table_modules"):
            raise NotImplementedError("GradArena.attach needs an avr_b200 field rendered on the tensor-core path")
        small = {id(m.params) for m in net.small_table_modules()}
        dense = [p for p in self.params if id(p) not in small]
        rows = [p for p in self.params if id(p) in small]
        self._layout(dense + rows, n_reduced=len(dense))
        renderer.row_exchange = RowExchange(group)
        return self

    @staticmethod
    def detach(renderer):
        renderer.row_exchange = None

    def checksum(self) -> float:
        """float64 sum of the arena -- equal on every rank after the exchange (bench.py asserts it)."""
        return float(self.flat.double().sum())


def shard_receivers(n_receivers: int, rank: int, world: int, drop_last: bool = False):
    """Indices of the receivers rank ``rank`` renders: ``rank, rank+world, ...`` over the index list padded by wrapping
    around to a multiple of ``world`` -- ``DistributedSampler(shuffle=False)`` order and padding
    (avr_runner_ddp.py:131-137), so every rank gets the same number of steps; ``drop_last`` truncates instead."""
    if n_receivers <= 0:
        return []
    idx = list(range(n_receivers))
    if drop_last:
        idx = idx[: n_receivers - n_receivers % world]
    else:
        total = (n_receivers + world - 1) // world * world
        while len(idx) < total:                                  # DistributedSampler: indices += indices[:padding]
            idx += idx[: total - len(idx)]
    return idx[rank::world]
