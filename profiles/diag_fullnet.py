"""Per-parameter errors of the real BASELINE fields on a reduced ray grid: tensor-core path (near-zero guard on / off) vs
exact-fp32 SIMT path vs the oracle's own fp32 noise (dense layers in float64).  Run on the GPU box.
    python profiles/diag_fullnet.py <config> <n_azi> <n_ele> [seed ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avr_b200
from avr_b200 import fused_tc
from avr_b200.configs import get_config
from oracle import field_ref, render_ref
from oracle.render_ref import rel_l2
from tests.helpers import oracle_fp32_noise
DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "simu"
n_azi, n_ele = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16, 8)
seeds = [int(s) for s in sys.argv[4:]] or [41, 43, 45]
for seed in seeds:
    cfg = get_config(name); cfg["render"]["n_azi"], cfg["render"]["n_ele"] = n_azi, n_ele
    mc = cfg["model_class"]
    cls = field_ref.AVRModelRef if mc == "AVRModel" else field_ref.AVRModelComplexRef
    ncls = avr_b200.AVRModel if mc == "AVRModel" else avr_b200.AVRModel_complex
    ref_net = field_ref.trained_like_(cls(cfg["model"], seed=seed), seed=seed + 1)
    r = cfg["render"]; bs = 2
    gen = torch.Generator().manual_seed(11)
    c = (r["xyz_min"] + r["xyz_max"]) / 2
    rx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 1.5).float(); tx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 1.5).float()
    dtx = torch.nn.functional.normalize(torch.randn(bs, 3, generator=gen), dim=-1) if mc != "AVRModel" else None
    azi = torch.rand(n_azi, generator=gen); T = cfg["model"]["signal_output_dim"]; G = torch.randn(bs, T // 2 + 1, 2, generator=gen)
    ref_out = render_ref.RenderRef(ref_net, **r)(rx, tx, dtx, azi_rand=azi); (ref_out * G).sum().backward()
    n_out, noise = oracle_fp32_noise(ref_net, r, rx, tx, G, dtx=dtx, azi_rand=azi)
    refg = dict(ref_net.named_parameters())
    res = {}
    for mode in ("tc", "tc-noguard", "simt"):
        fused_tc.NEAR_ZERO_GUARD = mode == "tc"
        native = ncls(cfg["model"]); native.load_state_dict(ref_net.state_dict()); native.to(DEV)
        out = avr_b200.AVRRender(native, **r, dense=mode.split("-")[0])(rx.to(DEV), tx.to(DEV), dtx.to(DEV) if dtx is not None else None, azi_rand=azi)
        (out * G.to(DEV)).sum().backward()
        res[mode] = (rel_l2(out.cpu(), ref_out), {n: rel_l2(p.grad.cpu(), refg[n].grad) for n, p in native.named_parameters()})
        if mode == "tc":
            print("guard counts per layer:", fused_tc.LAST_GUARD_COUNTS.tolist()[:14])
    print(f"seed {seed}: out tc {res['tc'][0]:.2e} noguard {res['tc-noguard'][0]:.2e} simt {res['simt'][0]:.2e} oracle-noise {n_out:.2e}")
    for n in refg:
        print(f"   {n:32s} tc {res['tc'][1][n]:.2e}  noguard {res['tc-noguard'][1][n]:.2e}  simt {res['simt'][1][n]:.2e}  oracle noise {noise[n]:.2e}")
