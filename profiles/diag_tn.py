"""Every weight-gradient GEMM of one tensor-core backward pass against float64 on the same operands: how much of its
error is the operands' 16-bit planes (vs the 24 bits the forward stored) and how ill-conditioned the sum is.
    python profiles/diag_tn.py <config> <n_azi> <n_ele> <seed>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avr_b200
from avr_b200 import ops
from avr_b200.ops import PlanePair
from avr_b200.configs import get_config
from oracle import field_ref
DEV = "cuda:0"
name, n_azi, n_ele, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cfg = get_config(name); cfg["render"]["n_azi"], cfg["render"]["n_ele"] = n_azi, n_ele
ref_net = field_ref.trained_like_(field_ref.AVRModelRef(cfg["model"], seed=seed), seed=seed + 1)
r = cfg["render"]; bs = 2
gen = torch.Generator().manual_seed(11)
c = (r["xyz_min"] + r["xyz_max"]) / 2
rx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 1.5).float(); tx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 1.5).float()
azi = torch.rand(n_azi, generator=gen); T = cfg["model"]["signal_output_dim"]; G = torch.randn(bs, T // 2 + 1, 2, generator=gen)
def merged(pp, n):
    q = PlanePair(pp.buf[:n].contiguous(), pp.col0, pp.cols, pp.row0, pp.rows) if n < pp.n else pp
    return ops.planes_merge(q).double()
def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
orig = ops.umma_tn
def spy(a, b, c_f32, workspace, accumulate=False):
    orig(a, b, c_f32, workspace, accumulate)
    if accumulate:
        return
    a2, b2, b3 = merged(a, 2), merged(b, 2), merged(b, b.n)
    full, two = a2.t() @ b3, a2.t() @ b2
    cond = float((a2.abs().t() @ b3.abs()).norm() / full.norm())
    print(f"  TN M={a.cols:4d} N={b.cols:4d}: kernel vs f64(g16,x{8*b.n}) {rel(c_f32.double(), full):.2e}   vs f64(g16,x16) {rel(c_f32.double(), two):.2e}"
          f"   x16-vs-x{8*b.n} {rel(two, full):.2e}   cancellation |g|^T|x| / |g^T x| = {cond:.1f}")
ops.umma_tn = spy
native = avr_b200.AVRModel(cfg["model"]); native.load_state_dict(ref_net.state_dict()); native.to(DEV)
out = avr_b200.AVRRender(native, **r)(rx.to(DEV), tx.to(DEV), azi_rand=azi)
(out * G.to(DEV)).sum().backward()
