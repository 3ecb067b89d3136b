"""The fused native render step on the tensor cores (default path of ``AVRRender`` for ``avr_b200`` fields).

Same contract as ``fused.FusedRenderFunction`` (SURVEY 8a rows a1-a11 in one autograd node) with two
B200-first changes:

* every activation is an error-compensated set of 16-bit planes (bf16 triples / pairs, fp16 pairs: DESIGN.md 4) and
  every dense layer runs on ``tcgen05.mma`` with TMEM accumulators (``csrc/umma_gemm.cu``): fp32-grade accuracy at
  tensor-core rate;
* the ``width -> T`` output layer of the signal network is never evaluated per sample point: it is fused with
  the delay-masked ray reduction as a prefix sum over delay-sorted rays (``csrc/collapse.cu``), so the
  ``[bs,R,S,T]`` signal tensor and its gradient (0.84 GB / receiver each at simu) do not exist.
"""
from __future__ import annotations

import torch

from . import ops
from .fused import plan_modules
from .ops import PlanePair


FWD_KIND = ops.PLANES_BF16x3  # forward operands carry 24 mantissa bits (six products): ReLU decisions must match fp32
SIG_HIDDEN_F16 = True
# The 512 x 512 hidden layers of the signal network (tensor-bound, half of the forward MMA time) take fp16
# (hi, lo' * 2^11) pairs: 24 bits in two planes, THREE products instead of six.  Range: fp16's -- like tiny-cuda-nn's own
# activations -- inf / NaN beyond 65504 (loud), absolute error <= 1.5e-11 below 6e-5.  False keeps bf16 triples everywhere
# (a module constant, not an environment switch: tests set it explicitly).
SIG_TN_FROM_F16 = True
# The weight-gradient GEMM of such a layer needs its activation as a bf16 pair (the gradient operand needs bf16's range,
# and tcgen05.mma faults on a bf16 x f16 mix).  True (default): avr_umma_gemm_tn takes the fp16 pair and its epilogue warps,
# idle during the main loop, convert every tile to bf16 (hi, mid) in shared memory -- bit-identical gradients, 2.1 GB less
# HBM traffic and activation memory per simu step, the two forward layers 0.82 -> 0.62 and 0.79 -> 0.65 ms.  False: the
# producing layer writes each hidden activation twice (ops.UMMA_DUAL_COPY: fp16 pair + bf16 pair).
# History (profiles/r2/ab_tn_from_f16*.txt, interleaved runs on one box): on the single-CTA weight-gradient kernel (two
# 96 KB pipeline stages, the conversion on the stage's critical path: 0.61 -> 0.87 ms per 512 x 512 gradient) it was a net
# loss, 11.76 vs 11.65 ms per step; as CTA pairs (each CTA converts only its half of the B tile, three 64 KB stages) it
# is a net gain: 11.28 vs 11.54 ms.
BWD_PLANES = 2      # gradients enter linearly: 16 bits / three products are enough ...
DENSITY_BWD_PLANES = 3   # ... except along the sigma decoder: the density gradient of a ray sums to ~0 over its samples
                         # (sum_s w_s = 1), so the decoder's weight-gradient sums cancel to ~1/50 of their terms
                         # (measured, profiles/diag_tn.py) and 16-bit operands leave ~1e-3 there
NEAR_ZERO_GUARD = False  # forward layers: re-evaluate pre-activations too close to zero for the tensor core's
                         # accumulation error with fp32 FMAs (ops.NearZeroGuard).  Needed with ONE accumulator per tile
                         # (error ~6e-9*K, 3-5 flipped ReLU decisions per 1e7 vs 0-1 for fp32); with the main/small
                         # accumulator pair the forward is fp32-grade (2.1e-7 at K = 512, cuBLAS fp32: 2.4e-7) and the
                         # guard changes nothing measurable (profiles/r1/parity_real_networks.md), so it is off
FUSE_SIGMA_CHAIN = True   # sigma encoder -> sigma decoder (all 128-wide layers, the fp32 density head) in ONE launch with the
                          # activation tile resident in shared memory (ops.mlp_chain_fwd, csrc/mlp_chain.cu) instead of one
                          # GEMM launch per layer -- bit-identical outputs; False runs the layers one by one (A/B, tests)
LAST_GUARD_COUNTS = None  # diagnostics: per-layer number of re-evaluated elements of the latest forward (device tensor)


def _lib_chain_max():
    from ._lib import CHAIN_MAX_LAYERS
    return CHAIN_MAX_LAYERS


def _weight_planes(net, params, transpose, n, jobs=None, count=None):
    """Plane sets (kind ``n``: 2 / 3 bf16 planes or ops.PLANES_F16x2; a list gives one kind per matrix) of the first
    ``count`` matrices of ``net`` (``W[out,in]``) or of their transposes (``W^T[in,out]``).  With ``jobs`` (a list) the
    conversions are only queued there: the caller runs all matrices of the pass in ONE launch (ops.planes_split_many)."""
    out, own = [], jobs is None
    jobs = [] if own else jobs
    mats = net.matrices(params)
    for k, w in enumerate(mats[:count] if count is not None else mats):
        o, i = w.shape
        kind = n[k] if isinstance(n, (list, tuple)) else n
        pp = PlanePair.empty(i, o, w.device, kind=kind) if transpose else PlanePair.empty(o, i, w.device, kind=kind)
        jobs.append((w, pp, transpose))
        out.append(pp)
    if own:
        ops.planes_split_many(jobs)
    return out


def _assemble(segments, buf: PlanePair, col, end_col, geom, small_in, rays_o, pos_tx, dirs, d_vals, params_of, delay_slot,
              rows_of=None):
    """Encode ``segments`` into columns ``[col, ...)`` of ``buf``; pad up to ``end_col`` with ones."""
    dev = rays_o.device
    if not segments and end_col > col:
        segments = [(None, "ones")]
    pending = None                                       # (small table, per_receiver, column) of a broadcast not yet written
    for k, (mod, kind) in enumerate(segments):
        w = mod.n_output_dims if mod is not None else 0
        last = k == len(segments) - 1
        n_ones = end_col - (col + w) if last else 0
        small = None
        if kind == "point":
            delay = delay_slot.pop() if delay_slot else None
            ops.raygen_encode_fwd(geom, mod.meta, rays_o, pos_tx, dirs, d_vals, params_of(mod), buf, col0=col,
                                  n_ones=n_ones, delay=delay)
        elif kind == "receiver_rows":                                        # channel-embedding rows (model.py:201-203)
            small, per_rcv = rows_of[mod.key], True
        elif kind != "ones":
            u = small_in[kind]
            small = torch.empty(u.shape[0], w, device=dev)
            ops.grid_encode_fwd(mod.meta, u, params_of(mod), small)
            per_rcv = kind != "ray"
        # adjacent broadcast blocks go out in pairs: one launch, whole sectors (a 40-column block alone ends mid-sector)
        if small is not None and pending is not None and small.shape[1] % 8 == 0 and pending[0].shape[1] % 8 == 0 \
                and pending[2] % 8 == 0:
            ops.rows_broadcast2(geom, pending[0], pending[1], small, per_rcv, buf, pending[2])
            pending = None
        else:
            if pending is not None:
                ops.rows_broadcast(geom, pending[0], pending[1], buf, pending[2])
            pending = (small, per_rcv, col) if small is not None else None
        if kind != "point" and n_ones:                                       # tcnn pads the network input with ones
            c0 = buf.col0 + col + w
            buf.buf[0, :, c0:c0 + n_ones] = 1.0
            buf.buf[1:, :, c0:c0 + n_ones] = 0.0
        col += w
    if pending is not None:
        ops.rows_broadcast(geom, pending[0], pending[1], buf, pending[2])
    return col


class FusedRenderTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, geom, tables, tspan, rays_o, pos_tx, dir_tx, dirs, *params):
        mods = plan_modules(plan)
        pmap = {id(m): p.detach() for m, p in zip(mods, params)}
        params_of = lambda m: pmap[id(m)]                                    # noqa: E731
        # per-receiver channel-embedding rows: biases of hidden layers / blocks of network inputs
        roles = plan.get("extras", [])
        extras = [t.detach().float().contiguous() for t in params[len(mods):]]
        if len(extras) != len(roles):
            raise ValueError("plan extras and extra tensors do not match")
        bias_of = {(r[1], r[2]): t for r, t in zip(roles, extras) if r[0] == "bias"}
        rows_of = {r[1]: t for r, t in zip(roles, extras) if r[0] == "rows"}

        def layer_flags(net_key, li):
            b = bias_of.get((net_key, li))
            return (ops.UMMA_RELU | ops.UMMA_BIAS, dict(bias_rcv=b, geom=geom)) if b is not None else (ops.UMMA_RELU, {})

        enc_net, dec_net, sig_net = plan["enc"], plan["dec"], plan["sig"]
        feat_dim, T = plan["feat_dim"], geom.T
        if sig_net.out_pad != T or T % 8:
            raise NotImplementedError("signal_output_dim must be a multiple of 8")
        if enc_net.out_pad != feat_dim or feat_dim % 8:
            raise NotImplementedError("sigma feature width must be a multiple of 16")
        dev = rays_o.device
        n_rows = geom.bs * geom.R * geom.S
        d_vals = tables["d"]
        u_view, u_tx, u_dtx = ops.aux_inputs(geom, pos_tx, dirs, dir_tx)
        small_in = {"ray": u_view, "receiver_tx": u_tx, "receiver_dir_tx": u_dtx}
        delay = torch.empty(geom.bs, geom.R, geom.S, dtype=torch.int32, device=dev)
        delay_slot = [delay]
        guard = ops.NearZeroGuard(dev) if NEAR_ZERO_GUARD else None
        global LAST_GUARD_COUNTS
        LAST_GUARD_COUNTS = guard.counters if guard is not None else None

        # ---- sigma encoder ------------------------------------------------------------------------
        x0 = PlanePair.empty(n_rows, enc_net.in_pad, dev, kind=FWD_KIND)
        _assemble(plan["x0"], x0, 0, enc_net.in_pad, geom, small_in, rays_o, pos_tx, dirs, d_vals, params_of, delay_slot,
                  rows_of)
        if delay_slot:
            _, _, _, delay = ops.sample_points(geom, rays_o, pos_tx, dirs, d_vals, want_pts=False)
        f16 = SIG_HIDDEN_F16 and guard is None
        sig_mats = sig_net.matrices(params_of(sig_net))
        n_sig = len(sig_mats) - 1                                            # hidden layers (the output layer is collapsed)
        jobs = []                                                            # every weight matrix of the pass -> planes, one launch
        w_enc = _weight_planes(enc_net, params_of(enc_net), False, FWD_KIND, jobs)
        w_dec = _weight_planes(dec_net, params_of(dec_net), False, FWD_KIND, jobs)
        w_sig = _weight_planes(sig_net, params_of(sig_net), False,          # layer li > 0 multiplies an fp16-pair activation
                               [ops.PLANES_F16x2 if (f16 and li > 0) else FWD_KIND for li in range(n_sig)], jobs, count=n_sig)
        ops.planes_split_many(jobs)
        sig_in = PlanePair.empty(n_rows, sig_net.in_pad, dev, kind=FWD_KIND)
        feat_win = sig_in.window(0, feat_dim)                                # sigma_feat lands directly in the signal network's input buffer (no concat)
        bits_feat = ops.relu_bits_empty(n_rows, feat_dim, dev)               # (sigma_feat > 0)
        # Inference (torch.no_grad(): nothing requires grad) skips everything only the backward pass reads: saved planes of
        # the 128-wide layers, ReLU bitmasks, the bf16 copies of the signal network's fp16-pair activations.
        train = any(ctx.needs_input_grad[8:])
        chain = (FUSE_SIGMA_CHAIN and guard is None and not plan["sig_relu_feat"] and feat_dim == 128
                 and dec_net.in_pad == feat_dim and not plan.get("dec_tail") and enc_net.in_pad <= 128 and enc_net.in_pad % 16 == 0
                 and all(o == 128 for (o, _) in enc_net.shapes) and all(o == 128 for (o, _) in dec_net.shapes[:-1])
                 and dec_net.out_pad <= 64 and dec_net.out_pad % 16 == 0 and len(enc_net.shapes) + len(dec_net.shapes) <= _lib_chain_max())
        if chain:
            # ---- sigma encoder -> [relu] -> sigma decoder -> density head: one launch ----------------------------------
            acts_enc = [PlanePair.empty(n_rows if train else 0, 128, dev, kind=ops.PLANES_BF16x2) for _ in w_enc[:-1]]   # weight gradients read 16 bits
            bits_enc = [ops.relu_bits_empty(n_rows if train else 0, 128, dev) for _ in w_enc[:-1]]
            acts_dec = [PlanePair.empty(n_rows if train else 0, 128, dev, kind=FWD_KIND) for _ in w_dec[:-1]]             # ... 24 along the decoder
            bits_dec = [ops.relu_bits_empty(n_rows if train else 0, 128, dev) for _ in w_dec[:-1]]
            dec_in = PlanePair.empty(n_rows if train else 0, dec_net.in_pad, dev, kind=FWD_KIND)
            dec_out = torch.empty(n_rows, dec_net.out_pad, device=dev)
            def emb(net_key, li):
                # per-receiver channel-embedding row of a hidden layer (model.py:44-47): a bias before the activation
                b = bias_of.get((net_key, li))
                return dict(bias=b, bias_group_rows=geom.R * geom.S) if b is not None else {}

            if train:
                layers = [dict(w=wp, relu=True, save=a, bits=b, **emb("enc", li))
                          for li, (wp, a, b) in enumerate(zip(w_enc[:-1], acts_enc, bits_enc))]
                layers.append(dict(w=w_enc[-1], relu=True, save_raw=feat_win, save=dec_in, bits=bits_feat))   # raw feat + relu(feat)
                layers += [dict(w=wp, relu=True, save=a, bits=b, **emb("dec", li))
                           for li, (wp, a, b) in enumerate(zip(w_dec[:-1], acts_dec, bits_dec))]
            else:
                layers = [dict(w=wp, relu=True, **emb("enc", li)) for li, wp in enumerate(w_enc[:-1])]
                layers.append(dict(w=w_enc[-1], relu=True, save_raw=feat_win))
                layers += [dict(w=wp, relu=True, **emb("dec", li)) for li, wp in enumerate(w_dec[:-1])]
            layers.append(dict(w=w_dec[-1], relu=False, out_f32=dec_out))                                 # the |leaky_relu| kink follows
            ops.mlp_chain(x0, layers)
        else:
            acts_enc, bits_enc, h = [], [], x0
            for li in range(len(w_enc) - 1):
                y = PlanePair.empty(n_rows, w_enc[li].rows, dev, kind=FWD_KIND)
                bits = ops.relu_bits_empty(n_rows, y.cols, dev)
                fl, kw = layer_flags("enc", li)
                ops.umma_nt(h, w_enc[li], fl, y, bits_out=bits, guard=guard, **kw)
                acts_enc.append(y)
                bits_enc.append(bits)
                h = y
            if plan["sig_relu_feat"]:
                ops.umma_nt(h, w_enc[-1], ops.UMMA_RELU, feat_win, bits_out=bits_feat, guard=guard)        # both consumers read relu(feat)
                dec_in = feat_win
            else:
                dec_buf = PlanePair.empty(n_rows, dec_net.in_pad, dev, kind=FWD_KIND)
                ops.umma_nt(h, w_enc[-1], ops.UMMA_DUAL_RELU, feat_win, dec_buf.window(0, feat_dim), bits_out=bits_feat, guard=guard)   # raw feat + relu(feat)
                if dec_net.in_pad > feat_dim:                                    # decoder input = [relu(feat), embedding row]
                    _assemble(plan.get("dec_tail", []), dec_buf, feat_dim, dec_net.in_pad, geom, small_in, rays_o, pos_tx, dirs,
                              d_vals, params_of, [], rows_of)
                dec_in = dec_buf
            if dec_in.cols != dec_net.in_pad:
                raise NotImplementedError("sigma decoder input width does not match the sigma feature width")

            # ---- sigma decoder -> density -> ray weights ---------------------------------------------------
            acts_dec, bits_dec, h = [], [], dec_in
            for li in range(len(w_dec) - 1):
                y = PlanePair.empty(n_rows, w_dec[li].rows, dev, kind=FWD_KIND)
                bits = ops.relu_bits_empty(n_rows, y.cols, dev)
                fl, kw = layer_flags("dec", li)
                ops.umma_nt(h, w_dec[li], fl, y, bits_out=bits, guard=guard, **kw)
                acts_dec.append(y)
                bits_dec.append(bits)
                h = y
            dec_out = torch.empty(n_rows, dec_net.out_pad, device=dev)
            ops.umma_nt(h, w_dec[-1], ops.UMMA_OUT_F32, c_f32=dec_out, guard=guard)   # the |leaky_relu| kink
        w, _ = ops.ray_weights_fwd(geom, dec_out, dec_out.stride(0), tables["delta"], plan["slope"])

        # ---- signal network hidden layers + collapsed output layer ---------------------------------
        _assemble(plan["tail"], sig_in, feat_dim, sig_net.in_pad, geom, small_in, rays_o, pos_tx, dirs, d_vals,
                  params_of, [], rows_of)
        acts_sig, bits_sig, h = [], [], sig_in                               # acts_sig: what the weight gradients read
        for li in range(n_sig):
            last = li == n_sig - 1
            bits = ops.relu_bits_empty(n_rows, w_sig[li].rows, dev) if (train and not last) else None   # the last one is read by collapse
            fl, kw = layer_flags("sig", li)
            if f16:
                y = PlanePair.empty(n_rows, w_sig[li].rows, dev, kind=ops.PLANES_F16x2)
                if last or not train or SIG_TN_FROM_F16:                     # the weight-gradient GEMM converts the fp16 pair itself
                    ops.umma_nt(h, w_sig[li], fl, y, bits_out=bits, **kw)
                    acts_sig.append(y)
                else:
                    y_bf = PlanePair.empty(n_rows, w_sig[li].rows, dev)
                    ops.umma_nt(h, w_sig[li], fl | ops.UMMA_DUAL_COPY, y, y_bf, bits_out=bits, **kw)
                    acts_sig.append(y_bf)
            else:
                y = PlanePair.empty(n_rows, w_sig[li].rows, dev, kind=FWD_KIND)
                ops.umma_nt(h, w_sig[li], fl, y, bits_out=bits, guard=guard, **kw)
                acts_sig.append(y)
            bits_sig.append(bits)
            h = y
        sort = ops.delay_sort(geom, delay, w)
        y_t, prefix = ops.collapse_fwd(geom, h, sort, sig_mats[-1], tspan)
        out = ops.spectrum_fwd_tc(geom, y_t, tables)

        if train:
            ctx.plan, ctx.geom, ctx.tables, ctx.tspan = plan, geom, tables, tspan
            ctx.small_in, ctx.roles = small_in, roles
            ctx.bufs = dict(x0=x0, acts_enc=acts_enc, sig_in=sig_in, dec_in=dec_in, acts_dec=acts_dec, dec_out=dec_out,
                            acts_sig=acts_sig, sort=sort, prefix=prefix, bits_enc=bits_enc, bits_dec=bits_dec, bits_sig=bits_sig,
                            bits_feat=bits_feat, chain=chain)
            ctx.save_for_backward(rays_o, dirs, *params)
        return out

    @staticmethod
    def backward(ctx, d_out):
        plan, geom, tables, B = ctx.plan, ctx.geom, ctx.tables, ctx.bufs
        rays_o, dirs, *params = ctx.saved_tensors
        mods = plan_modules(plan)
        pmap = {id(m): p.detach() for m, p in zip(mods, params)}
        extra_grads = {}                                                     # role -> gradient of that extra tensor
        bias_layers = {(r[1], r[2]) for r in ctx.roles if r[0] == "bias"}

        def bias_grad(net_key, li, g):
            """d(embedding row) = sum of the layer's pre-activation gradient over each receiver's points."""
            if (net_key, li) in bias_layers:
                extra_grads[("bias", net_key, li)] = ops.rows_reduce(geom, g, 0, g.cols, True)

        enc_net, dec_net, sig_net = plan["enc"], plan["dec"], plan["sig"]
        feat_dim = plan["feat_dim"]
        dev = d_out.device
        n_rows = geom.bs * geom.R * geom.S
        d_vals = tables["d"]
        grads = {}
        # Data-parallel runs (ddp.GradArena.attach): the per-ray / per-receiver tables receive gradient from only R / bs
        # points, so instead of all-reducing their (almost all-zero) 38-145 MB gradients the ranks all-gather the few
        # pre-scatter rows and every rank scatters all of them locally -- ONE collective for all such tables, at the end of
        # the pass (a collective started mid-pass takes an SM from the persistent GEMM kernels: measured slower).
        exchange = plan.get("row_exchange")
        deferred = []

        grids = [m for (m, kind) in plan["x0"] + plan["tail"] if kind != "receiver_rows" and m.grid_grad == "deterministic"]
        scratch = torch.empty(max(int(m.meta.total) * 2 for m in grids), dtype=torch.int64, device=dev) if grids else None

        def scatter_segments(segments, d_buf):
            """Hash-table gradients (scatter; mode per Encoding.grid_grad) and embedding-row gradients of one input block."""
            col = 0
            for mod, kind in segments:
                wdt = mod.n_output_dims
                if kind == "receiver_rows":
                    extra_grads[("rows", mod.key)] = ops.rows_reduce(geom, d_buf, col, wdt, True)
                    col += wdt
                    continue
                if kind == "point":
                    acc = ops.GridGradAccumulator(mod.meta, dev, n_rows, scratch, mod.grid_grad)
                    acc.observe(d_buf, col, wdt)
                    acc.add_rays(geom, rays_o, dirs, d_vals, d_buf, col, sample_step=tables.get("sample_step", 0.0))
                else:
                    small = ops.rows_reduce(geom, d_buf, col, wdt, kind != "ray")
                    if exchange is not None:
                        deferred.append((mod, ctx.small_in[kind], small))
                        col += wdt
                        continue
                    acc = ops.GridGradAccumulator(mod.meta, dev, small.shape[0], scratch, mod.grid_grad)
                    acc.observe(small, 0, wdt)
                    acc.add_points(ctx.small_in[kind], small)
                grads[id(mod)] = acc.finalize()
                col += wdt

        ws_bytes = max(ops.umma_tn_workspace_bytes(o, i, n_rows) for n in (enc_net, dec_net, sig_net) for (o, i) in n.shapes)
        ws = torch.empty(max(4, ws_bytes // 4), device=dev)

        jobs = []                                                            # all transposed weight planes of the pass: one launch
        wt_sig_all = _weight_planes(sig_net, pmap[id(sig_net)], True, BWD_PLANES, jobs, count=len(sig_net.shapes) - 1)
        wt_dec = _weight_planes(dec_net, pmap[id(dec_net)], True, DENSITY_BWD_PLANES, jobs)
        wt_enc = _weight_planes(enc_net, pmap[id(enc_net)], True, BWD_PLANES, jobs)
        ops.planes_split_many(jobs)

        def hidden_backward(net, net_key, g, acts, bits, first_input, g_flat, wt):
            """Back-propagate through layers len(acts)..1 of ``net`` given g = d(pre-activation of the last
            hidden layer); fills the weight gradients of layers >= 1 and of layer 0; returns g at layer 0."""
            d_mats = net.matrices(g_flat)
            bias_grad(net_key, len(acts) - 1, g)
            for li in range(len(acts) - 1, 0, -1):
                x = acts[li - 1]
                ops.umma_tn(g, x, d_mats[li], ws)
                gx = PlanePair.empty(n_rows, x.cols, dev)
                ops.umma_nt(g, wt[li], ops.UMMA_MASK, gx, mask=bits[li - 1])
                g = gx
                bias_grad(net_key, li - 1, g)
            ops.umma_tn(g, first_input, d_mats[0], ws)
            return g, wt, d_mats

        # ---- spectrum, collapsed output layer --------------------------------------------------------
        d_y = ops.spectrum_bwd_tc(geom, d_out.contiguous().float(), tables)
        sig_mats = sig_net.matrices(pmap[id(sig_net)])
        g_sig = torch.empty_like(pmap[id(sig_net)])
        d_sig_mats = sig_net.matrices(g_sig)
        h_last = B["acts_sig"][-1]
        g = PlanePair.empty(n_rows, h_last.cols, dev)
        d_w = ops.collapse_bwd(geom, h_last, B["sort"], sig_mats[-1], d_y, ctx.tspan, B["prefix"], g, d_sig_mats[-1])
        B["prefix"] = None

        # ---- signal network hidden layers ----------------------------------------------------------------
        sig_in, dec_in = B["sig_in"], B["dec_in"]
        g, wt_sig, _ = hidden_backward(sig_net, "sig", g, B["acts_sig"], B["bits_sig"], sig_in, g_sig, wt_sig_all)
        grads[id(sig_net)] = g_sig
        wt0 = wt_sig[0]                                                      # W0^T planes [in_pad, width]
        tail_w = sig_net.in_pad - feat_dim
        if plan["sig_relu_feat"]:
            d_feat = PlanePair.empty(n_rows, feat_dim, dev)
            ops.umma_nt(g, wt0.row_window(0, feat_dim), ops.UMMA_MASK, d_feat, mask=B["bits_feat"])
            d_tail = PlanePair.empty(n_rows, tail_w, dev)
            ops.umma_nt(g, wt0.row_window(feat_dim, tail_w), 0, d_tail)
        else:
            # d(signal-network input) in ONE product -- g is read once instead of twice; d_feat and d_tail are its column windows
            d_sig_in = PlanePair.empty(n_rows, sig_net.in_pad, dev)
            ops.umma_nt(g, wt0, 0, d_sig_in)
            d_feat, d_tail = d_sig_in.window(0, feat_dim), d_sig_in.window(feat_dim, tail_w)
        B["acts_sig"] = None
        scatter_segments(plan["tail"], d_tail)
        d_tail = None

        # ---- density path: ray weights -> sigma decoder ---------------------------------------------------
        d_dec_out = torch.zeros_like(B["dec_out"])
        ops.ray_weights_bwd(geom, B["dec_out"], B["dec_out"].stride(0), tables["delta"], plan["slope"], d_w, d_dec_out,
                            d_dec_out.stride(0))
        g_dec = torch.empty_like(pmap[id(dec_net)])
        d_dec_mats = dec_net.matrices(g_dec)
        g = PlanePair.empty(n_rows, dec_net.out_pad, dev, n=DENSITY_BWD_PLANES)
        ops.planes_split(d_dec_out, g)
        acts_dec, acts_enc, x0 = B["acts_dec"], B["acts_enc"], B["x0"]
        n_dec = len(d_dec_mats)
        g_enc = torch.empty_like(pmap[id(enc_net)])
        d_enc_mats = enc_net.matrices(g_enc)
        n_enc = len(d_enc_mats)
        d_x0 = PlanePair.empty(n_rows, enc_net.in_pad, dev)
        if B.get("chain"):
            # ---- backward-data of sigma decoder -> d_feat (+= signal path) -> sigma encoder: ONE launch; the gradient
            # tile stays in shared memory, every layer's output is stored once for the weight-gradient GEMMs below
            g_dec_l = [g] + [PlanePair.empty(n_rows, 128, dev, n=DENSITY_BWD_PLANES) for _ in range(n_dec - 1)]     # g at layer n_dec-1 .. 1
            g_enc_l = [d_feat] + [PlanePair.empty(n_rows, 128, dev) for _ in range(n_enc - 1)]                       # g at layer n_enc-1 .. 1
            layers = []
            for k, li in enumerate(range(n_dec - 1, 0, -1)):
                layers.append(dict(w=wt_dec[li], mask=B["bits_dec"][li - 1], save=g_dec_l[k + 1]))
            layers.append(dict(w=wt_dec[0], mask=B["bits_feat"], save=d_feat, accumulate=True))                    # d_feat += relu'(feat) * ...
            for k, li in enumerate(range(n_enc - 1, 0, -1)):
                layers.append(dict(w=wt_enc[li], mask=B["bits_enc"][li - 1], save=g_enc_l[k + 1]))
            layers.append(dict(w=wt_enc[0], save=d_x0))
            ops.mlp_chain(g, layers)
            for k, li in enumerate(range(n_dec - 1, 0, -1)):
                ops.umma_tn(g_dec_l[k], acts_dec[li - 1], d_dec_mats[li], ws)
                bias_grad("dec", li - 1, g_dec_l[k + 1])                         # d(embedding row) of hidden layer li - 1
            ops.umma_tn(g_dec_l[-1], dec_in, d_dec_mats[0], ws)
            for k, li in enumerate(range(n_enc - 1, 0, -1)):
                ops.umma_tn(g_enc_l[k], acts_enc[li - 1], d_enc_mats[li], ws)
                bias_grad("enc", li - 1, g_enc_l[k + 1])
            ops.umma_tn(g_enc_l[-1], x0, d_enc_mats[0], ws)
            grads[id(dec_net)] = g_dec
            grads[id(enc_net)] = g_enc
        else:
            for li in range(n_dec - 1, 0, -1):
                x = acts_dec[li - 1]
                ops.umma_tn(g, x, d_dec_mats[li], ws)
                gx = PlanePair.empty(n_rows, x.cols, dev, n=DENSITY_BWD_PLANES)
                ops.umma_nt(g, wt_dec[li], ops.UMMA_MASK, gx, mask=B["bits_dec"][li - 1])
                g = gx
                bias_grad("dec", li - 1, g)
            ops.umma_tn(g, dec_in, d_dec_mats[0], ws)
            ops.umma_nt(g, wt_dec[0].row_window(0, feat_dim), ops.UMMA_MASK | ops.UMMA_ACCUM, d_feat, mask=B["bits_feat"])   # d_feat += relu'(feat) * ...
            d_dec_tail = None
            if dec_net.in_pad > feat_dim and plan.get("dec_tail"):
                d_dec_tail = PlanePair.empty(n_rows, dec_net.in_pad - feat_dim, dev)
                ops.umma_nt(g, wt_dec[0].row_window(feat_dim, dec_net.in_pad - feat_dim), 0, d_dec_tail)
            grads[id(dec_net)] = g_dec
            scatter_segments(plan.get("dec_tail", []), d_dec_tail)

            # ---- sigma encoder -----------------------------------------------------------------------------------
            g = d_feat
            for li in range(n_enc - 1, 0, -1):
                x = acts_enc[li - 1]
                ops.umma_tn(g, x, d_enc_mats[li], ws)
                gx = PlanePair.empty(n_rows, x.cols, dev)
                ops.umma_nt(g, wt_enc[li], ops.UMMA_MASK, gx, mask=B["bits_enc"][li - 1])
                g = gx
                bias_grad("enc", li - 1, g)
            ops.umma_tn(g, x0, d_enc_mats[0], ws)
            ops.umma_nt(g, wt_enc[0], 0, d_x0)
            grads[id(enc_net)] = g_enc
        scatter_segments(plan["x0"], d_x0)
        gathered = exchange.gather([(u, small) for _, u, small in deferred]) if deferred else []
        for (mod, _, _), (u_all, rows_all) in zip(deferred, gathered):
            # rows of ALL ranks (already scaled to the mean), scattered in rank order with the int64 accumulator: every
            # replica gets the bit-identical gradient, like after an all-reduce
            acc = ops.GridGradAccumulator(mod.meta, dev, rows_all.shape[0], scratch, "deterministic")
            acc.observe(rows_all, 0, rows_all.shape[1])
            acc.add_points(u_all, rows_all)
            grads[id(mod)] = acc.finalize()
        ctx.bufs = None
        return (None,) * 8 + tuple(grads[id(m)] for m in mods) + tuple(extra_grads[tuple(r)] for r in ctx.roles)
