"""Checkpoints in the reference's on-disk format (``avr_runner.py:104-154``).

A reference checkpoint is ``torch.save`` of::

    {"current_iteration": int,
     "audionerf_network_state_dict": renderer.state_dict(),      # keys network_fn._pos_encoding.params, ...
     "optimizer_state_dict": torch.optim.Adam(...).state_dict(),
     "scheduler_state_dict": CosineAnnealingLR(...).state_dict()}

``avr_b200`` modules keep the reference's parameter names and tiny-cuda-nn's flat fp32 layouts,
and ``FusedAdam`` reads / writes ``torch.optim.Adam``'s state, so the file is interchangeable in both directions.
tiny-cuda-nn stores half-precision-padded copies nowhere in the state dict: ``params`` is its fp32 master tensor.
"""
from __future__ import annotations

import os

import torch

KEY_ITER, KEY_NET, KEY_OPT, KEY_SCHED = ("current_iteration", "audionerf_network_state_dict", "optimizer_state_dict",
                                         "scheduler_state_dict")


def _unwrap(renderer):
    return renderer.module if hasattr(renderer, "module") and isinstance(renderer.module, torch.nn.Module) else renderer


def save_checkpoint(path, renderer, optimizer=None, scheduler=None, current_iteration: int = 0) -> str:
    """Write ``path`` (``{:06d}.tar`` in the reference) with the reference's four keys."""
    blob = {KEY_ITER: int(current_iteration), KEY_NET: _unwrap(renderer).state_dict()}
    if optimizer is not None:
        blob[KEY_OPT] = optimizer.state_dict()
    if scheduler is not None:
        blob[KEY_SCHED] = scheduler.state_dict()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(blob, path)
    return path


def load_checkpoint(path, renderer, optimizer=None, scheduler=None, map_location=None, trusted: bool = False) -> int:
    """Load a checkpoint written by the reference runners or by ``save_checkpoint``; -> ``current_iteration``.

    Network keys may carry the ``module.`` prefix of ``DataParallel`` / ``DDP``; shapes must match the configured
    field exactly (a mismatch is an error, never a silent partial load).  The file is read with
    ``weights_only=True`` (tensors and plain containers -- all a reference checkpoint holds); ``trusted=True`` allows
    arbitrary pickled objects for files of known origin."""
    ckpt = torch.load(path, map_location=map_location, weights_only=not trusted)
    state = ckpt[KEY_NET]
    if state and all(k.startswith("module.") for k in state):
        state = {k[len("module."):]: v for k, v in state.items()}
    _unwrap(renderer).load_state_dict(state, strict=True)
    if optimizer is not None and KEY_OPT in ckpt:
        optimizer.load_state_dict(ckpt[KEY_OPT])
    if scheduler is not None and KEY_SCHED in ckpt:
        scheduler.load_state_dict(ckpt[KEY_SCHED])
    return int(ckpt.get(KEY_ITER, 0))


def latest_checkpoint(ckpts_dir):
    """The reference picks the lexicographically last ``*.tar`` of ``<logdir>/<expname>/ckpts`` (avr_runner.py:110-115)."""
    if not os.path.isdir(ckpts_dir):
        return None
    names = [f for f in sorted(os.listdir(ckpts_dir)) if "tar" in f]
    return os.path.join(ckpts_dir, names[-1]) if names else None
