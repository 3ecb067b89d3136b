"""Experiment: are the TC-path gradient errors caused by ReLU-mask disagreements (forward accuracy)?
Forward in exact fp32 (SIMT), activations then split to 2 planes, backward on the tensor cores."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, avr_b200
from avr_b200 import ops, fused, fused_tc
from avr_b200.functional import DenseStack
from avr_b200.ops import PlanePair
from avr_b200.configs import tiny_config
from oracle import field_ref, render_ref
from tests.helpers import rel_l2
DEV = "cuda:0"

def planes(x):
    return ops.planes_split(x.contiguous(), PlanePair.empty(x.shape[0], x.shape[1], DEV))

class Ctx:
    pass

def run(native, ren, rx, tx, dtx, azi, G):
    plan = native.fused_plan()
    T = native.signal_output_dim
    geom = ops.make_geom(ren.render_cfg(), rx.shape[0], T)
    tab = ren.tables_for(T, DEV).dev
    from avr_b200 import tables
    dirs = tables.direction_table(ren.n_azi, ren.n_ele, azi).to(DEV)
    mods = fused.plan_modules(plan)
    pmap = {id(m): m.params.detach() for m in mods}
    params_of = lambda m: pmap[id(m)]
    enc_net, dec_net, sig_net = plan["enc"], plan["dec"], plan["sig"]
    fd = plan["feat_dim"]
    u_view, u_tx, u_dtx = ops.aux_inputs(geom, tx, dirs, dtx)
    small_in = {"ray": u_view, "receiver_tx": u_tx, "receiver_dir_tx": u_dtx}
    delay = torch.empty(geom.bs, geom.R, geom.S, dtype=torch.int32, device=DEV)
    x0 = fused._assemble(plan["x0"], enc_net.in_pad, geom, small_in, rx, tx, dirs, tab["d"], params_of, [delay])
    feat, acts_enc = DenseStack(enc_net, params_of(enc_net)).forward([(x0, False)])
    dec_out, acts_dec = DenseStack(dec_net, params_of(dec_net)).forward([(feat, True)])
    w, _ = ops.ray_weights_fwd(geom, dec_out, dec_out.stride(0), tab["delta"], plan["slope"])
    tail = fused._assemble(plan["tail"], sig_net.in_pad - fd, geom, small_in, rx, tx, dirs, tab["d"], params_of, [])
    sig, acts_sig = DenseStack(sig_net, params_of(sig_net)).forward([(feat, plan["sig_relu_feat"]), (tail, False)])
    featp = feat.clamp_min(0) if plan["sig_relu_feat"] else feat
    sig_in = planes(torch.cat([featp, tail], 1))
    dec_in = sig_in.window(0, fd) if plan["sig_relu_feat"] else planes(feat.clamp_min(0))
    sort = ops.delay_sort(geom, delay, w)
    h = planes(acts_sig[-1])
    y = ops.collapse_fwd(geom, h, sort, sig_net.matrices(params_of(sig_net))[-1])
    out = ops.spectrum_fwd(geom, y, tab)
    ctx = Ctx()
    ctx.plan, ctx.geom, ctx.tables, ctx.tspan = plan, geom, tab, ops.collapse_tspan(ren.render_cfg())
    ctx.small_in = small_in
    ctx.bufs = dict(x0=planes(x0), acts_enc=[planes(a) for a in acts_enc], sig_in=sig_in, dec_in=dec_in,
                    acts_dec=[planes(a) for a in acts_dec], dec_out=dec_out, acts_sig=[planes(a) for a in acts_sig[:-1]] + [h], sort=sort)
    ctx.saved_tensors = (rx, dirs, *[m.params.detach() for m in mods])
    grads = fused_tc.FusedRenderTC.backward(ctx, G)[8:]
    return out, {id(m): g for m, g in zip(mods, grads)}, mods

for mc, kw, bs in [("AVRModel", dict(n_azi=12, n_ele=6, n_samples=24, T=400, width_sigma=64, width_signal=128), 3),
                   ("AVRModel_complex", dict(n_azi=10, n_ele=5, n_samples=16, T=480, fs=8000, xyz_min=-12, xyz_max=12), 2)]:
    cfg = tiny_config(mc, **kw)
    cls = field_ref.AVRModelRef if mc == "AVRModel" else field_ref.AVRModelComplexRef
    ref_net = field_ref.trained_like_(cls(cfg["model"], seed=21), seed=22)
    r = cfg["render"]
    gen = torch.Generator().manual_seed(5)
    c = (r["xyz_min"] + r["xyz_max"]) / 2
    rx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 2).float(); tx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 2).float()
    dtx = torch.nn.functional.normalize(torch.randn(bs, 3, generator=gen), dim=-1) if mc != "AVRModel" else None
    azi = torch.rand(r["n_azi"], generator=gen)
    G = torch.randn(bs, kw["T"] // 2 + 1, 2, generator=gen)
    ref_out = render_ref.RenderRef(ref_net, **r)(rx, tx, dtx, azi_rand=azi)
    (ref_out * G).sum().backward()
    ncls = avr_b200.AVRModel if mc == "AVRModel" else avr_b200.AVRModel_complex
    native = ncls(cfg["model"]); native.load_state_dict(ref_net.state_dict()); native = native.to(DEV)
    ren = avr_b200.AVRRender(native, **r)
    with torch.no_grad():
        out, grads, mods = run(native, ren, rx.to(DEV), tx.to(DEV), dtx.to(DEV) if dtx is not None else None, azi, G.to(DEV))
    print(mc, "fp32 forward + TC backward: out", "%.2e" % rel_l2(out, ref_out))
    rg = dict(ref_net.named_parameters())
    names = {id(m): n for n, m in native.named_modules()}
    for m in mods:
        n_ = names[id(m)] + ".params"
        print("   ", n_, "%.2e" % rel_l2(grads[id(m)], rg[n_].grad))
