"""Data-parallel gradient exchange: one flat fp32 arena, one all-reduce per step.

Replaces the DDP reducer of ``avr_runner_ddp.py:98,257`` (25 MB buckets + a used-parameter bitmap
all-reduce per step because of ``find_unused_parameters=True``).  Receivers are independent, every rank
holds a full replica, and the only exchange is the mean of the parameter gradients: all
``.grad`` tensors are views into one contiguous buffer, so the exchange is a single NCCL all-reduce
(NVLS in-switch reduction on NVSwitch) issued on the compute stream right after the backward kernels
(``all_reduce_mean``) -- or, with ``attach(renderer)``, one all-reduce per parameter tensor issued from INSIDE the
backward pass the moment that tensor's gradient is final: the signal network and the per-ray / per-receiver hash
tables (2/3 of the bytes) are reduced while the density path and the sigma encoder are still back-propagating,
and only the last table's exchange (38 of 120 MB at simu) is exposed.  Measured on 8 x B200 (``profiles/ddp_overlap_check.py``,
``profiles/r1/ddp_overlap.md``): bit-identical gradients at 2 ranks, but NOT faster -- 15.17 -> 15.10 ms at 2 GPUs,
15.44 -> 15.63 ms at 8: the persistent GEMM CTAs own every SM, so NCCL's kernels only make progress between them
and nine small all-reduces cost more latency than one large one.  ``attach`` therefore stays opt-in.

The reference's stock wrapper also works on ``avr_b200.AVRRender`` (it is an ordinary ``nn.Module``);
this arena is the B200-first path used by ``bench.py``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradArena:
    """Owns a flat gradient buffer and points every parameter's ``.grad`` into it."""

    def __init__(self, parameters):
        self.params = [p for p in parameters if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dtype = self.params[0].device, self.params[0].dtype
        self.offsets, n = [], 0
        for p in self.params:
            if p.device != dev or p.dtype != dtype:
                raise ValueError("all parameters must share device and dtype")
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4                       # keep every view 16-byte aligned
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
        """grad <- mean over ranks (DistributedDataParallel semantics)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        if self.flat.is_cuda and dist.get_backend(group) == "nccl":
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=False)
        self.flat.div_(world)
        return work


    # -- exchange overlapped with the backward pass -------------------------------------------------------------
    def attach(self, renderer, group=None):
        """Reduce every gradient inside ``renderer``'s backward pass (tensor-core path) instead of after it.

        The renderer announces ``(parameter, gradient)`` as soon as the producing kernels are enqueued; the all-reduce
        (mean) of that tensor starts on NCCL's stream behind them and runs next to the remaining backward kernels.
        ``done`` makes the compute stream wait for all of them before autograd accumulates the (already averaged)
        gradients into ``.grad``.  Do not call ``all_reduce_mean`` as well."""
        self._works = []

        def ready(param, grad):
            if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
                return
            if grad.is_cuda and dist.get_backend(group) == "nccl":
                self._works.append(dist.all_reduce(grad, op=dist.ReduceOp.AVG, group=group, async_op=True))
            else:
                dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
                grad.div_(dist.get_world_size(group))

        def done():
            for w in self._works:
                w.wait()                                         # stream-level wait, the host does not block
            self._works.clear()

        renderer.grad_ready_hook, renderer.grad_done_hook = ready, done
        return self

    @staticmethod
    def detach(renderer):
        renderer.grad_ready_hook = renderer.grad_done_hook = None


def shard_receivers(n_receivers: int, rank: int, world: int):
    """Indices of the receivers rank ``rank`` renders: ``rank, rank+world, ...``
    (``DistributedSampler`` order without shuffling, avr_runner_ddp.py:131-137)."""
    return list(range(rank, n_receivers, world))
