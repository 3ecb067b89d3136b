"""The fused native render step: ray generation -> hash-grid encode -> MLPs -> composite -> spectrum.

One ``torch.autograd.Function`` spanning SURVEY 8a rows a1-a11 for a field built from
``avr_b200.model`` modules.  Nothing of shape ``[bs, P, 3]`` is materialised, the per-ray / per-receiver
encodings are evaluated on their ``R`` / ``bs`` distinct inputs only, ``concat`` is
replaced by column-block GEMMs, and the backward returns one flat gradient per parameter tensor with the
hash-table gradients accumulated deterministically.
"""
from __future__ import annotations

import torch

from . import ops
from .functional import DenseStack


def plan_modules(plan: dict):
    """Modules whose flat ``params`` the fused step differentiates (per-receiver row blocks carry no parameters here)."""
    segs = plan["x0"] + plan["tail"]
    return [m for (m, kind) in segs if kind != "receiver_rows"] + [plan["enc"], plan["dec"], plan["sig"]]


def _assemble(segments, total_width, geom, small_in, rays_o, pos_tx, dirs, d_vals, params_of, delay_slot):
    """Write the concatenated encodings of ``segments`` into a fresh ``[N, total_width]`` buffer."""
    n_rows = geom.bs * geom.R * geom.S
    dev = rays_o.device
    buf = torch.empty(n_rows, total_width, device=dev)
    col = 0
    for k, (mod, kind) in enumerate(segments):
        w = mod.n_output_dims
        last = k == len(segments) - 1
        n_ones = total_width - (col + w) if last else 0
        if kind == "point":
            delay = delay_slot.pop() if delay_slot else None
            ops.raygen_encode_fwd(geom, mod.meta, rays_o, pos_tx, dirs, d_vals, params_of(mod), buf, col0=col,
                                  n_ones=n_ones, delay=delay)
        else:
            u = small_in[kind]
            small = torch.empty(u.shape[0], w, device=dev)
            ops.grid_encode_fwd(mod.meta, u, params_of(mod), small)
            ops.rows_broadcast(geom, small, kind != "ray", buf, col)
            if n_ones:
                buf[:, col + w:] = 1.0
        col += w
    return buf


class FusedRenderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, geom, tables, rays_o, pos_tx, dir_tx, dirs, *params):
        mods = plan_modules(plan)
        pmap = {id(m): p.detach() for m, p in zip(mods, params)}
        params_of = lambda m: pmap[id(m)]                                    # noqa: E731
        enc_net, dec_net, sig_net = plan["enc"], plan["dec"], plan["sig"]
        feat_dim, T = plan["feat_dim"], geom.T
        if sig_net.out_pad != T:
            raise NotImplementedError("signal_output_dim must be a multiple of 8")
        if enc_net.out_pad != feat_dim:
            raise NotImplementedError("sigma feature width must be a multiple of 16")
        d_vals = tables["d"]
        u_view, u_tx, u_dtx = ops.aux_inputs(geom, pos_tx, dirs, dir_tx)
        small_in = {"ray": u_view, "receiver_tx": u_tx, "receiver_dir_tx": u_dtx}
        delay = torch.empty(geom.bs, geom.R, geom.S, dtype=torch.int32, device=rays_o.device)
        delay_slot = [delay]

        x0 = _assemble(plan["x0"], enc_net.in_pad, geom, small_in, rays_o, pos_tx, dirs, d_vals, params_of, delay_slot)
        if delay_slot:                                                       # no per-point encoding in x0
            _, _, _, delay = ops.sample_points(geom, rays_o, pos_tx, dirs, d_vals, want_pts=False)
        enc_stack = DenseStack(enc_net, params_of(enc_net))
        feat, acts_enc = enc_stack.forward([(x0, False)])
        dec_stack = DenseStack(dec_net, params_of(dec_net))
        dec_out, acts_dec = dec_stack.forward([(feat, True)])
        w, _ = ops.ray_weights_fwd(geom, dec_out, dec_out.stride(0), tables["delta"], plan["slope"])
        tail = _assemble(plan["tail"], sig_net.in_pad - feat_dim, geom, small_in, rays_o, pos_tx, dirs, d_vals,
                         params_of, [])
        sig_stack = DenseStack(sig_net, params_of(sig_net))
        sig, acts_sig = sig_stack.forward([(feat, plan["sig_relu_feat"]), (tail, False)])
        y = ops.composite_fwd(geom, sig, w, delay)
        out = ops.spectrum_fwd(geom, y, tables)

        if any(ctx.needs_input_grad[7:]):
            ctx.plan, ctx.geom, ctx.tables = plan, geom, tables
            ctx.small_in = small_in
            ctx.bufs = dict(x0=x0, feat=feat, acts_enc=acts_enc, dec_out=dec_out, acts_dec=acts_dec, tail=tail,
                            acts_sig=acts_sig, sig=sig, w=w, delay=delay)
            ctx.save_for_backward(rays_o, dirs, *params)
        return out

    @staticmethod
    def backward(ctx, d_out):
        plan, geom, tables, B = ctx.plan, ctx.geom, ctx.tables, ctx.bufs
        rays_o, dirs, *params = ctx.saved_tensors
        mods = plan_modules(plan)
        pmap = {id(m): p.detach() for m, p in zip(mods, params)}
        enc_net, dec_net, sig_net = plan["enc"], plan["dec"], plan["sig"]
        feat_dim = plan["feat_dim"]
        dev = d_out.device
        n_rows = geom.bs * geom.R * geom.S
        d_vals = tables["d"]
        grads = {}

        d_y = ops.spectrum_bwd(geom, d_out.contiguous().float(), tables)
        d_sig, d_w = ops.composite_bwd(geom, B["sig"], B["w"], B["delay"], d_y)
        B["sig"] = None
        d_sig = d_sig.view(n_rows, geom.T)

        stacks = {k: DenseStack(n, pmap[id(n)]) for k, n in (("enc", enc_net), ("dec", dec_net), ("sig", sig_net))}
        ws_bytes = max(ops.gemm_workspace_bytes(o, i, n_rows) for n in (enc_net, dec_net, sig_net) for (o, i) in n.shapes)
        ws = torch.empty(max(4, ws_bytes // 4), device=dev)

        feat = B["feat"]
        d_feat = torch.empty(n_rows, feat_dim, device=dev)
        d_tail = torch.empty_like(B["tail"])
        g_sig = torch.empty_like(pmap[id(sig_net)])
        stacks["sig"].backward([(feat, plan["sig_relu_feat"]), (B["tail"], False)], B["acts_sig"], d_sig, g_sig, ws,
                               [(d_feat, False, feat if plan["sig_relu_feat"] else None), (d_tail, False, None)])
        grads[id(sig_net)] = g_sig
        del d_sig
        B["acts_sig"] = None

        d_dec_out = torch.zeros_like(B["dec_out"])
        ops.ray_weights_bwd(geom, B["dec_out"], B["dec_out"].stride(0), tables["delta"], plan["slope"], d_w, d_dec_out,
                            d_dec_out.stride(0))
        g_dec = torch.empty_like(pmap[id(dec_net)])
        stacks["dec"].backward([(feat, True)], B["acts_dec"], d_dec_out, g_dec, ws, [(d_feat, True, feat)])
        grads[id(dec_net)] = g_dec

        d_x0 = torch.empty_like(B["x0"])
        g_enc = torch.empty_like(pmap[id(enc_net)])
        stacks["enc"].backward([(B["x0"], False)], B["acts_enc"], d_feat, g_enc, ws, [(d_x0, False, None)])
        grads[id(enc_net)] = g_enc

        grids = [m for (m, kind) in plan["x0"] + plan["tail"] if kind != "receiver_rows" and m.grid_grad == "deterministic"]
        scratch = torch.empty(max(int(m.meta.total) * 2 for m in grids), dtype=torch.int64, device=dev) if grids else None
        for segments, d_buf in ((plan["x0"], d_x0), (plan["tail"], d_tail)):
            col = 0
            for mod, kind in segments:
                wdt = mod.n_output_dims
                if kind == "point":
                    acc = ops.GridGradAccumulator(mod.meta, dev, n_rows, scratch, mod.grid_grad)
                    acc.observe(d_buf, col, wdt)
                    acc.add_rays(geom, rays_o, dirs, d_vals, d_buf, col, sample_step=tables.get("sample_step", 0.0))
                else:
                    small = ops.rows_reduce(geom, d_buf, col, wdt, kind != "ray")
                    acc = ops.GridGradAccumulator(mod.meta, dev, small.shape[0], scratch, mod.grid_grad)
                    acc.observe(small, 0, wdt)
                    acc.add_points(ctx.small_in[kind], small)
                grads[id(mod)] = acc.finalize()
                col += wdt
        ctx.bufs = None
        return (None,) * 7 + tuple(grads[id(m)] for m in mods)
