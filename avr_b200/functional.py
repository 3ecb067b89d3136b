"""Autograd glue between PyTorch and the C-ABI kernels (SURVEY 8b "autograd glue").

Each ``torch.autograd.Function`` launches only ``libavr_b200`` kernels on the current stream and
returns gradients for the flat parameter tensors (``None`` for positions, which never require grad in
the reference -- avr_runner.py:168-176).  Backward runs on the autograd engine's device thread; every
C call therefore takes the device ordinal and the stream explicitly (``ops._ctx``).
"""
from __future__ import annotations

import torch

from . import ops


def unit_cube(x: torch.Tensor) -> torch.Tensor:
    """model.py:187-189: ``(x.view(-1, 3) + 1) / 2``."""
    return (x.reshape(-1, 3) + 1) / 2


class HashGridFunction(torch.autograd.Function):
    """Encode explicit unit-cube points; deterministic table gradient."""

    @staticmethod
    def forward(ctx, u, params, enc):
        out = torch.empty(u.shape[0], enc.n_output_dims, device=u.device)
        ops.grid_encode_fwd(enc.meta, u, params.detach(), out)
        ctx.enc = enc
        ctx.save_for_backward(u)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (u,) = ctx.saved_tensors
        enc = ctx.enc
        d_out = d_out.contiguous()
        acc = ops.GridGradAccumulator(enc.meta, d_out.device, u.shape[0], None, enc.grid_grad)
        acc.observe(d_out, 0, enc.n_output_dims)
        acc.add_points(u, d_out)
        return None, acc.finalize(), None


class DenseStack:
    """Forward / backward of one bias-free ReLU MLP over row-major activations (no autograd).

    The first layer may read its input in several column blocks (``parts``): ``h0 = relu(sum_p
    x_p @ W0[:, c_p:c_p+w_p]^T)`` -- this is how the concatenation of model.py:221,326 is avoided.
    """

    def __init__(self, net, params):
        self.net = net
        self.mats = net.matrices(params)

    def forward(self, parts, keep=True):
        """parts: list of (x[N,w], relu_in) covering W0's columns in order. -> (out[N,out_pad], acts)."""
        n = parts[0][0].shape[0]
        dev = parts[0][0].device
        acts = []
        h = torch.empty(n, self.mats[0].shape[0], device=dev)
        c0 = 0
        for k, (x, relu_in) in enumerate(parts):
            w = self.mats[0][:, c0:c0 + x.shape[1]]
            last = k == len(parts) - 1
            ops.linear_fwd(x, w, h, relu=last and len(self.mats) > 1, relu_in=relu_in, accum=k > 0)
            c0 += x.shape[1]
        assert c0 == self.mats[0].shape[1], "input blocks do not cover the first matrix"
        acts.append(h)
        for li in range(1, len(self.mats)):
            w = self.mats[li]
            y = torch.empty(n, w.shape[0], device=dev)
            ops.linear_fwd(h, w, y, relu=li < len(self.mats) - 1)
            if li < len(self.mats) - 1:
                acts.append(y)
            h = y
        return h, acts

    def backward(self, parts, acts, d_out, d_params, workspace, want_dx):
        """d_out[N,out_pad] -> fills d_params (flat, same layout) and returns d_x per part (or None).

        ``want_dx[p]`` is False, True, or a tuple ``(dst, accum, mask_src)`` naming the tensor to write
        the input gradient into (used to accumulate both consumers of ``sigma_feat`` in place).
        """
        d_mats = self.net.matrices(d_params)
        n_l = len(self.mats)
        g = d_out
        for li in range(n_l - 1, 0, -1):
            x = acts[li - 1]                                   # post-ReLU input of layer li
            ops.linear_bwd_weight(g, x, d_mats[li], workspace)
            gx = torch.empty_like(x)
            ops.linear_bwd_data(g, self.mats[li], gx, mask_src=x)
            g = gx
        # first layer: g = d(pre-activation of h0) already masked by relu'(h0)
        d_parts, c0 = [], 0
        for (x, relu_in), want in zip(parts, want_dx):
            wdt = x.shape[1]
            ops.linear_bwd_weight(g, x, d_mats[0][:, c0:c0 + wdt], workspace, relu_in=relu_in)
            if want is False or want is None:
                d_parts.append(None)
            else:
                if want is True:
                    dst, accum, mask_src = torch.empty(x.shape[0], wdt, device=x.device), False, (x if relu_in else None)
                else:
                    dst, accum, mask_src = want
                ops.linear_bwd_data(g, self.mats[0][:, c0:c0 + wdt], dst, mask_src=mask_src, accum=accum)
                d_parts.append(dst)
            c0 += wdt
        return d_parts

    def weight_grad_workspace(self, n_rows, device):
        nbytes = max(ops.gemm_workspace_bytes(o, i, n_rows) for (o, i) in self.net.shapes)
        return torch.empty(max(4, nbytes // 4), device=device)


class MLPFunction(torch.autograd.Function):
    """``tcnn.Network`` forward on explicit inputs ``x[N, n_in]`` (ones-padded to ``in_pad``)."""

    @staticmethod
    def forward(ctx, x, params, net):
        n = x.shape[0]
        if net.in_pad != net.n_input_dims:
            xp = torch.ones(n, net.in_pad, device=x.device)
            xp[:, : net.n_input_dims] = x
        else:
            xp = x
        stack = DenseStack(net, params.detach())
        out, acts = stack.forward([(xp, False)])
        ctx.net = net
        ctx.save_for_backward(xp, params, *acts)
        return out[:, : net.n_output_dims]

    @staticmethod
    def backward(ctx, d_y):
        xp, params, *acts = ctx.saved_tensors
        net = ctx.net
        n = xp.shape[0]
        d_out = torch.zeros(n, net.out_pad, device=xp.device)
        d_out[:, : net.n_output_dims] = d_y
        stack = DenseStack(net, params.detach())
        d_params = torch.empty_like(params)
        ws = stack.weight_grad_workspace(n, xp.device)
        (d_xp,) = stack.backward([(xp, False)], acts, d_out, d_params, ws, [ctx.needs_input_grad[0]])
        d_x = d_xp[:, : net.n_input_dims] if d_xp is not None else None
        return d_x, d_params, None


class CompositeFunction(torch.autograd.Function):
    """renderer.py:79-121 on explicit network outputs: ``(attn[bs,R,S], signal[bs,R,S,T]) -> [bs,F,2]``."""

    @staticmethod
    def forward(ctx, attn, signal, delay, geom, tables):
        attn = attn.contiguous().float()
        signal = signal.contiguous().float()
        w, _ = ops.ray_weights_fwd(geom, attn, 1, tables["delta"], -1.0)
        y = ops.composite_fwd(geom, signal, w, delay)
        out = ops.spectrum_fwd(geom, y, tables)
        ctx.geom, ctx.tables = geom, tables
        ctx.save_for_backward(attn, signal, delay, w)
        return out

    @staticmethod
    def backward(ctx, d_out):
        attn, signal, delay, w = ctx.saved_tensors
        geom, tables = ctx.geom, ctx.tables
        d_y = ops.spectrum_bwd(geom, d_out.contiguous().float(), tables)
        d_sig, d_w = ops.composite_bwd(geom, signal, w, delay, d_y, want_dsig=ctx.needs_input_grad[1],
                                       want_dw=ctx.needs_input_grad[0])
        d_attn = None
        if d_w is not None:
            d_attn = torch.empty_like(attn)
            ops.ray_weights_bwd(geom, attn, 1, tables["delta"], -1.0, d_w, d_attn, 1)
        return d_attn, d_sig, None, None, None
