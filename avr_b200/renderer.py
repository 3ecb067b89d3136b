"""``AVRRender`` -- drop-in for ``/root/reference/renderer.py`` (class ``AVRRender``, lines 14-124).

Same constructor (``AVRRender(networks_fn, **cfg['render'])``, keys read as at renderer.py:20-29, extras
ignored), same ``forward(rays_o, position_tx, direction_tx=None, ch_idx=None) -> [bs, T//2+1, 2]`` fp32 on
the input device, usable under ``torch.optim``, ``state_dict``, ``DDP`` and ``torch.no_grad()`` exactly
like the reference module (avr_runner.py:60-73,168-178; avr_runner_ddp.py:98).

Two execution paths, both made only of ``libavr_b200`` CUDA kernels (there is no CPU fallback -- CPU
tensors raise):

* ``networks_fn`` is an ``avr_b200.model`` field (has ``fused_plan``): the fused native step
  (``fused.FusedRenderFunction``).
* any other ``networks_fn(pts, view, tx[, dir_tx], ch_idx=...) -> (attn, signal)``: the ray-generation
  kernel materialises the network inputs the reference would build, the user network runs as is, and the
  compositing / spectrum kernels consume its outputs (``functional.CompositeFunction``).
"""
from __future__ import annotations

import inspect

import torch
import torch.nn as nn

from . import _lib, ops, tables
from .functional import CompositeFunction
from .fused import FusedRenderFunction, plan_modules
from .fused_tc import FusedRenderTC


class AVRRender(nn.Module):
    """Acoustic volume renderer (see module docstring)."""

    def __init__(self, networks_fn, **kwargs) -> None:
        super().__init__()
        self.network_fn = networks_fn
        self.n_samples = kwargs["n_samples"]
        self.near = kwargs["near"]
        self.far = kwargs["far"]
        self.n_azi = kwargs["n_azi"]
        self.n_ele = kwargs["n_ele"]
        self.speed = kwargs["speed"]
        self.fs = kwargs["fs"]
        self.pathloss = kwargs["pathloss"]
        self.xyz_min = kwargs["xyz_min"]
        self.xyz_max = kwargs["xyz_max"]
        #: receivers rendered per kernel pass (bounds the activation memory of large inference batches)
        self.max_receivers_per_pass = int(kwargs.get("max_receivers_per_pass", 8))
        #: "tc": dense layers on tcgen05 tensor cores (error-compensated 16-bit plane sets) + collapsed output layer;
        #: "simt": exact-fp32 FMA GEMMs and the literal signal tensor (slower; kept as an independent check)
        self.dense = kwargs.get("dense", "tc")
        if self.dense not in ("tc", "simt"):
            raise ValueError("dense must be 'tc' or 'simt'")
        #: accumulation of the hash-table gradients, applied to every ``Encoding`` of the field when given:
        #: "atomic" (default of ``Encoding``) or "deterministic" (bit-reproducible, ~0.7 ms/step slower on simu)
        grid_grad = kwargs.get("grid_grad")
        if grid_grad is not None:
            from .model import Encoding
            from .ops import GRID_GRAD_MODES
            if grid_grad not in GRID_GRAD_MODES:
                raise ValueError(f"grid_grad must be one of {GRID_GRAD_MODES}")
            for m in networks_fn.modules():
                if isinstance(m, Encoding):
                    m.grid_grad = grid_grad
        self._tables = {}
        #: data-parallel runs: exchange object installed by ``ddp.GradArena.attach`` -- the backward pass all-gathers the
        #: pre-scatter gradient rows of the per-ray / per-receiver tables instead of leaving their (almost all-zero)
        #: table gradients to the all-reduce (tensor-core path)
        self.row_exchange = None

    # -- configuration -----------------------------------------------------------------------------
    def render_cfg(self) -> dict:
        return {"n_samples": self.n_samples, "near": self.near, "far": self.far, "n_azi": self.n_azi,
                "n_ele": self.n_ele, "speed": self.speed, "fs": self.fs, "pathloss": self.pathloss,
                "xyz_min": self.xyz_min, "xyz_max": self.xyz_max}

    def tables_for(self, T: int, device) -> tables.RenderTables:
        key = (int(T), str(device))
        if key not in self._tables:
            self._tables[key] = tables.RenderTables(self.render_cfg(), int(T), device)
        return self._tables[key]

    # -- forward -----------------------------------------------------------------------------------
    def forward(self, rays_o, position_tx, direction_tx=None, ch_idx=None, azi_rand=None):
        """rays_o[bs,3] receiver positions, position_tx[bs,3], direction_tx[bs,3] (RAF), ch_idx[bs].

        ``azi_rand`` (``[n_azi]`` in [0,1), optional) injects the azimuth jitter that the reference draws
        from the global CPU generator on every call (renderer.py:149); tests use it for repeatability.
        """
        if not rays_o.is_cuda:
            raise _lib.AVRLibraryError("avr_b200.AVRRender runs on CUDA tensors only (no CPU fallback); "
                                       "the CPU reference path is oracle/render_ref.py")
        _lib.load()
        device = rays_o.device
        rays_o = rays_o.detach().contiguous().float()
        position_tx = position_tx.detach().to(device).contiguous().float()
        if direction_tx is not None:
            direction_tx = direction_tx.detach().to(device).contiguous().float()
        dirs = tables.direction_table(self.n_azi, self.n_ele, azi_rand).to(device, non_blocking=True)
        bs = position_tx.size(0)
        step = max(1, self.max_receivers_per_pass)
        outs = []
        for b0 in range(0, bs, step):
            sl = slice(b0, min(bs, b0 + step))
            outs.append(self._render_pass(rays_o[sl], position_tx[sl],
                                          direction_tx[sl] if direction_tx is not None else None,
                                          ch_idx[sl] if ch_idx is not None else None, dirs))
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    def graphed_inference(self, bs: int, device=None, direction_tx: bool = False, ch_idx: bool = False):
        """The ``torch.no_grad()`` forward for a fixed batch size as ONE CUDA graph: ``f = ren.graphed_inference(1)`` then
        ``f(rays_o, position_tx[, direction_tx][, ch_idx][, azi_rand=...]) -> [bs, F, 2]`` replays ~50 kernel launches with
        a single ``cudaGraphLaunch``.  For the reference's evaluation loops, which render one receiver per call
        (eval_rotate_doa_avr.py:104-110, avr_runner.py:235-247): at bs = 1 the eager step is bound by launch latency.
        The graph reads the field's CURRENT parameters on every replay.  The returned tensor is the graph's static output
        buffer: it is overwritten by the next call (clone it to keep it)."""
        return GraphedInference(self, int(bs), device, direction_tx, ch_idx)

    def _render_pass(self, rays_o, position_tx, direction_tx, ch_idx, dirs):
        net = self.network_fn
        bs = rays_o.size(0)
        plan = None
        if hasattr(net, "fused_plan"):
            plan = net.fused_plan(ch_idx) if "ch_idx" in inspect.signature(net.fused_plan).parameters else net.fused_plan()
            if plan.get("extras") and self.dense != "tc":
                # channel embeddings on the fp32 SIMT path: the field is evaluated on explicit points (its own forward:
                # hash-grid gather + fp32 FMA GEMMs, embedding rows added / concatenated by torch) and composited by the
                # generic path below -- an independent check of the tensor-core path's bias epilogue / row broadcasts
                plan = None
        if plan is not None:
            if plan["needs_dir_tx"] and direction_tx is None:
                raise ValueError("this field needs direction_tx (AVRModel_complex, model.py:291)")
            T = int(net.signal_output_dim)
            tab = self.tables_for(T, rays_o.device)
            geom = ops.make_geom(self.render_cfg(), bs, T)
            params = [m.params for m in plan_modules(plan)] + list(plan.get("extra_tensors", []))
            dtx = direction_tx if plan["needs_dir_tx"] else None
            if self.dense == "tc":
                if self.row_exchange is not None:                              # ddp.GradArena.attach
                    plan = dict(plan)
                    plan["row_exchange"] = self.row_exchange
                return FusedRenderTC.apply(plan, geom, tab.dev, ops.collapse_tspan(self.render_cfg()), rays_o,
                                           position_tx, dtx, dirs, *params)
            if self.row_exchange is not None:
                raise NotImplementedError("GradArena.attach (row exchange) is built on the tensor-core path (dense='tc')")
            return FusedRenderFunction.apply(plan, geom, tab.dev, rays_o, position_tx, dtx, dirs, *params)

        # generic networks_fn: build what renderer.py:54-62 builds, call the network, composite.
        cfg = self.render_cfg()
        geom0 = ops.make_geom(cfg, bs, 1)
        d_vals = torch.linspace(0., 1., self.n_samples) * (self.far - self.near) + self.near
        d_vals = d_vals.to(rays_o.device)
        pts, view, txn, _ = ops.sample_points(geom0, rays_o, position_tx, dirs, d_vals, want_delay=False)
        args = [pts, view, txn]
        if direction_tx is not None:
            args.append(direction_tx[:, None, :].expand(bs, pts.size(1), 3))
        kwargs = {}
        if ch_idx is not None and "ch_idx" in inspect.signature(net.forward).parameters:
            kwargs["ch_idx"] = ch_idx
        if self.row_exchange is not None:
            raise NotImplementedError("GradArena.attach (row exchange) needs an avr_b200 field on the tensor-core path")
        attn, signal = net(*args, **kwargs)
        T = signal.size(-1)
        geom = ops.make_geom(cfg, bs, T)
        tab = self.tables_for(T, rays_o.device)
        _, _, _, delay = ops.sample_points(geom, rays_o, position_tx, dirs, tab.dev["d"], want_pts=False)
        attn = attn.reshape(bs, geom.R, geom.S)
        signal = signal.reshape(bs, geom.R, geom.S, T)
        return CompositeFunction.apply(attn, signal, delay, geom, tab.dev)


class GraphedInference:
    """See ``AVRRender.graphed_inference``."""

    def __init__(self, ren: AVRRender, bs: int, device=None, with_direction_tx: bool = False, with_ch_idx: bool = False):
        if device is None:
            device = next(ren.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.AVRLibraryError("graphed inference needs a CUDA device")
        _lib.load()
        self.ren, self.bs, self.device = ren, bs, device
        n_rays = ren.n_azi * ren.n_ele + 2
        self.rx = torch.zeros(bs, 3, device=device)
        self.tx = torch.zeros(bs, 3, device=device)
        self.dtx = torch.zeros(bs, 3, device=device) if with_direction_tx else None
        if self.dtx is not None:
            self.dtx[:, 0] = 1.0
        self.ch = torch.zeros(bs, dtype=torch.long, device=device) if with_ch_idx else None
        self.dirs = tables.direction_table(ren.n_azi, ren.n_ele, torch.zeros(ren.n_azi)).to(device)
        self.dirs_host = torch.empty(n_rays, 3).pin_memory()
        self.copied = torch.cuda.Event()
        side = torch.cuda.Stream(device)                                   # warm-up off the capture: caches, allocator pools
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                ren._render_pass(self.rx, self.tx, self.dtx, self.ch, self.dirs)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = ren._render_pass(self.rx, self.tx, self.dtx, self.ch, self.dirs)

    @torch.no_grad()
    def __call__(self, rays_o, position_tx, direction_tx=None, ch_idx=None, azi_rand=None):
        if rays_o.shape[0] != self.bs:
            raise ValueError(f"this graph was captured for {self.bs} receivers per call")
        self.rx.copy_(rays_o, non_blocking=True)
        self.tx.copy_(position_tx, non_blocking=True)
        if self.dtx is not None:
            if direction_tx is None:
                raise ValueError("this graph was captured with direction_tx")
            self.dtx.copy_(direction_tx, non_blocking=True)
        if self.ch is not None:
            if ch_idx is None:
                raise ValueError("this graph was captured with ch_idx")
            self.ch.copy_(ch_idx, non_blocking=True)
        self.copied.synchronize()                                          # the previous call's upload has left the pinned buffer
        self.dirs_host.copy_(tables.direction_table(self.ren.n_azi, self.ren.n_ele, azi_rand))
        self.dirs.copy_(self.dirs_host, non_blocking=True)
        self.copied.record()
        self.graph.replay()
        return self.out
