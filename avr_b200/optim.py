"""Fused training update over the flat arenas: clip-norm -> NaN/Inf scrub -> Adam in one pass.

Replaces ``avr_runner.py:192-200`` (``clip_grad_norm_(params, max_norm=1)``, the in-place NaN/Inf scrub loop over
every ``.grad`` and ``torch.optim.Adam.step()``).  Parameters are re-homed into one contiguous fp32 buffer (their
``nn.Parameter`` objects become views of it, so ``state_dict`` / the renderer keep working) next to the
``GradArena`` gradient buffer, and the whole update is one kernel of 28 bytes of HBM traffic per parameter.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .ddp import GradArena


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` semantics (lr, betas, eps, weight_decay as L2) + the reference's clipping and scrubbing.

    A ``torch.optim.Optimizer``: ``param_groups[0]["lr"]`` is what ``step`` uses, so ``CosineAnnealingLR``
    (avr_runner.py:71,200) drives it unchanged, and ``state_dict`` / ``load_state_dict`` speak ``torch.optim.Adam``'s
    format, so the ``optimizer_state_dict`` of a reference checkpoint (avr_runner.py:122,150) loads and saves.

    Known deviation: the update always covers the whole arena with ONE step count.  ``torch.optim.Adam`` skips a
    parameter whose ``.grad`` is ``None`` (no moment decay, no weight decay, its own step count); here a parameter that
    received no gradient in a step (``layer_embeddings`` when ``ch_idx`` is None) is updated with a zero gradient, like
    ``torch.optim.Adam`` after ``zero_grad(set_to_none=False)``.  The fused render step produces a gradient for every
    field parameter on every step, so the two agree whenever ``ch_idx`` is passed consistently.
    """

    def __init__(self, parameters, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=1.0,
                 arena: GradArena | None = None):
        params = [p for p in parameters if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                                      maximize=False, foreach=None, capturable=False, differentiable=False, fused=None))
        self.arena = arena if arena is not None else GradArena(params)
        if [id(p) for p in self.arena.params] != [id(p) for p in params]:
            raise ValueError("arena and optimiser must cover the same parameters in the same order")
        self.max_norm = max_norm
        self.step_count = 0
        flat = self.arena.flat
        if not flat.is_cuda:
            raise _lib.AVRLibraryError("FusedAdam needs CUDA parameters (no CPU fallback)")
        self.flat_params = torch.zeros_like(flat)
        with torch.no_grad():
            for p, off in zip(self.arena.params, self.arena.offsets):       # re-home the parameters
                view = self.flat_params[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.norm = torch.zeros(2, device=flat.device)
        self._ws = torch.empty(int(_lib.load().avr_adam_workspace_bytes()) // 4 + 4, device=flat.device)
        self._bind_state()

    # hyper-parameters live in the (single) param group, like torch.optim.Adam
    lr = property(lambda self: self.param_groups[0]["lr"], lambda self, v: self.param_groups[0].__setitem__("lr", v))
    betas = property(lambda self: self.param_groups[0]["betas"])
    eps = property(lambda self: self.param_groups[0]["eps"])
    weight_decay = property(lambda self: self.param_groups[0]["weight_decay"])

    def _bind_state(self):
        """Per-parameter ``state`` entries are views of the flat moment buffers (torch.optim.Adam's keys)."""
        for p, off in zip(self.arena.params, self.arena.offsets):
            n = p.numel()
            self.state[p] = {"step": torch.tensor(float(self.step_count)),
                             "exp_avg": self.exp_avg[off:off + n].view_as(p),
                             "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p)}

    def state_dict(self):
        for p in self.arena.params:
            self.state[p]["step"] = torch.tensor(float(self.step_count))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """Accepts a ``torch.optim.Adam`` state dict over the same parameters (e.g. from a reference checkpoint)."""
        if len(state_dict["param_groups"]) != 1:
            raise ValueError("FusedAdam keeps one parameter group")
        super().load_state_dict(state_dict)
        steps = set()
        with torch.no_grad():
            for p, off in zip(self.arena.params, self.arena.offsets):
                st = self.state.get(p, {})
                n = p.numel()
                if "exp_avg" in st:
                    self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                    self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.add(int(st["step"]))
                else:                                                     # Adam creates state lazily: never stepped
                    self.exp_avg[off:off + n].zero_()
                    self.exp_avg_sq[off:off + n].zero_()
                    steps.add(0)
        if len(steps) > 1:
            raise ValueError("parameters with different step counts cannot share the fused update")
        self.step_count = steps.pop() if steps else 0
        self._bind_state()

    def zero_grad(self, set_to_none: bool = False):
        self.arena.zero_()

    @torch.no_grad()
    def step(self, closure=None, lr=None, write_back_grad=False):
        """Apply one update; ``lr`` overrides ``param_groups[0]["lr"]`` for this step."""
        if closure is not None:
            raise NotImplementedError("closures are not supported")
        self.step_count += 1
        flat = self.arena.flat
        dev = flat.device.index if flat.device.index is not None else torch.cuda.current_device()
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        p = C.c_void_p
        _lib.check(_lib.load().avr_fused_adam_step(
            p(self.flat_params.data_ptr()), p(flat.data_ptr()), p(self.exp_avg.data_ptr()), p(self.exp_avg_sq.data_ptr()),
            flat.numel(), float(self.lr if lr is None else lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
            float(self.weight_decay), float(self.max_norm if self.max_norm else 0.0), self.step_count,
            1 if write_back_grad else 0, p(self.norm.data_ptr()), p(self._ws.data_ptr()), self._ws.numel() * 4, dev, st),
            "avr_fused_adam_step")

    def grad_norm(self) -> torch.Tensor:
        """Total gradient norm of the last step (device scalar; no sync)."""
        return self.norm[0]
