"""Where does the tensor-core path's density gradient differ from the exact-fp32 SIMT path?  Captures dec_out, d_w and
d_dec_out of both paths on the same problem.   python profiles/diag_density.py <config> <n_azi> <n_ele> <seed>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avr_b200
from avr_b200 import ops
from avr_b200.configs import get_config
from oracle import field_ref
from oracle.render_ref import rel_l2
DEV = "cuda:0"
name, n_azi, n_ele, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cfg = get_config(name); cfg["render"]["n_azi"], cfg["render"]["n_ele"] = n_azi, n_ele
ref_net = field_ref.trained_like_(field_ref.AVRModelRef(cfg["model"], seed=seed), seed=seed + 1)
r = cfg["render"]; bs = 2
gen = torch.Generator().manual_seed(11)
c = (r["xyz_min"] + r["xyz_max"]) / 2
rx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 1.5).float(); tx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 1.5).float()
azi = torch.rand(n_azi, generator=gen); T = cfg["model"]["signal_output_dim"]; G = torch.randn(bs, T // 2 + 1, 2, generator=gen)
cap = {}
orig_bwd = ops.ray_weights_bwd
def spy_bwd(g, raw, ld_raw, delta, slope, d_w, d_raw, ld_draw):
    orig_bwd(g, raw, ld_raw, delta, slope, d_w, d_raw, ld_draw)
    cap[mode] = dict(dec_out=raw.clone(), d_w=d_w.clone(), d_dec_out=d_raw.clone())
ops.ray_weights_bwd = spy_bwd
orig_fwd = ops.ray_weights_fwd
def spy_fwd(g, raw, ld_raw, delta, slope, want_attn=False):
    w, a = orig_fwd(g, raw, ld_raw, delta, slope, want_attn)
    capw[mode] = w.clone()
    return w, a
capw = {}
ops.ray_weights_fwd = spy_fwd
grads = {}
for mode in ("tc", "simt"):
    native = avr_b200.AVRModel(cfg["model"]); native.load_state_dict(ref_net.state_dict()); native.to(DEV)
    out = avr_b200.AVRRender(native, **r, dense=mode)(rx.to(DEV), tx.to(DEV), azi_rand=azi)
    (out * G.to(DEV)).sum().backward()
    grads[mode] = {n: p.grad.clone() for n, p in native.named_parameters()}
a, b = cap["tc"], cap["simt"]
S = r["n_samples"]
print("dec_out col0 rel", rel_l2(a["dec_out"][:, 0].cpu(), b["dec_out"][:, 0].cpu()), " w rel", rel_l2(capw["tc"].cpu(), capw["simt"].cpu()))
print("d_w rel", rel_l2(a["d_w"].cpu(), b["d_w"].cpu()))
print("d_dec_out col0 rel", rel_l2(a["d_dec_out"][:, 0].cpu(), b["d_dec_out"][:, 0].cpu()))
# same d_w through both -> isolates ray_weights_bwd inputs
d1 = torch.zeros_like(a["d_dec_out"]); g = ops.make_geom(r, bs, T)
tab_delta = avr_b200.AVRRender(native, **r).tables_for(T, DEV).dev["delta"]
orig_bwd(g, a["dec_out"], 16, tab_delta, 0.01, b["d_w"], d1, 16)
print("d_dec_out with TC dec_out + SIMT d_w vs SIMT:", rel_l2(d1[:, 0].cpu(), b["d_dec_out"][:, 0].cpu()))
d2 = torch.zeros_like(a["d_dec_out"])
orig_bwd(g, b["dec_out"], 16, tab_delta, 0.01, a["d_w"], d2, 16)
print("d_dec_out with SIMT dec_out + TC d_w vs SIMT:", rel_l2(d2[:, 0].cpu(), b["d_dec_out"][:, 0].cpu()))
diff = (a["d_dec_out"][:, 0] - b["d_dec_out"][:, 0]).abs()
top = diff.topk(8)
print("largest |d_dec_out| differences (row, ray, sample, tc, simt, dec_out tc, dec_out simt, d_w tc, d_w simt, w):")
dw_t, dw_s = a["d_w"].reshape(-1), b["d_w"].reshape(-1)
for i in top.indices.tolist():
    print(f"  row {i} ray {(i // S) % g.R} s {i % S}: {a['d_dec_out'][i,0].item():+.6e} {b['d_dec_out'][i,0].item():+.6e} | x {a['dec_out'][i,0].item():+.6e} {b['dec_out'][i,0].item():+.6e} | d_w {dw_t[i].item():+.6e} {dw_s[i].item():+.6e} | w {capw['tc'].reshape(-1)[i].item():.3e}")
print("norms: d_dec_out", float(b["d_dec_out"][:, 0].norm()), " diff", float(diff.norm()))
for n in grads["tc"]:
    print(f"  {n:32s} tc vs simt {rel_l2(grads['tc'][n].cpu(), grads['simt'][n].cpu()):.2e}")
