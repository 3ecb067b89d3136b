import sys, os
sys.path.insert(0, os.getcwd())
import torch, avr_b200
from avr_b200 import ops
from avr_b200.ops import PlanePair
from avr_b200.configs import tiny_config
from oracle import field_ref, render_ref
from tests.helpers import rel_l2
DEV="cuda:0"
# 1. raw GEMM error levels
g = torch.Generator().manual_seed(0)
for (M,N,K) in [(4096,512,512),(4096,128,128)]:
    A, B = torch.randn(M,K,generator=g), torch.randn(N,K,generator=g)
    a = ops.planes_split(A.to(DEV), PlanePair.empty(M,K,DEV)); b = ops.planes_split(B.to(DEV), PlanePair.empty(N,K,DEV))
    Aq, Bq = ops.planes_merge(a).cpu().double(), ops.planes_merge(b).cpu().double()
    c32 = torch.empty(M,N,device=DEV); ops.umma_nt(a,b,ops.UMMA_OUT_F32,c_f32=c32)
    print("gemm", M,N,K, "vs exact inputs", rel_l2(c32, A.double()@B.double().t()), "vs quantised inputs", rel_l2(c32, Aq@Bq.t()),
          "fp32 torch", rel_l2((A@B.t()), A.double()@B.double().t()))
# 2. end-to-end
for mc, kw, bs in [("AVRModel", dict(n_azi=12, n_ele=6, n_samples=24, T=400, width_sigma=64, width_signal=128), 3),
                   ("AVRModel_complex", dict(n_azi=10, n_ele=5, n_samples=16, T=480, fs=8000, xyz_min=-12, xyz_max=12), 2)]:
    cfg = tiny_config(mc, **kw)
    cls = field_ref.AVRModelRef if mc == "AVRModel" else field_ref.AVRModelComplexRef
    ref_net = field_ref.trained_like_(cls(cfg["model"], seed=21), seed=22)
    ref64 = field_ref.trained_like_(cls(cfg["model"], seed=21), seed=22).double()
    r = cfg["render"]
    gen = torch.Generator().manual_seed(5)
    c = (r["xyz_min"] + r["xyz_max"]) / 2
    rx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 2).float(); tx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 2).float()
    dtx = torch.nn.functional.normalize(torch.randn(bs, 3, generator=gen), dim=-1) if mc != "AVRModel" else None
    azi = torch.rand(r["n_azi"], generator=gen)
    G = torch.randn(bs, kw["T"] // 2 + 1, 2, generator=gen)
    ref_out = render_ref.RenderRef(ref_net, **r)(rx, tx, dtx, azi_rand=azi)
    (ref_out * G).sum().backward()
    for dense in ("simt", "tc"):
        ncls = avr_b200.AVRModel if mc == "AVRModel" else avr_b200.AVRModel_complex
        native = ncls(cfg["model"]); native.load_state_dict(ref_net.state_dict()); native = native.to(DEV)
        ren = avr_b200.AVRRender(native, **r, dense=dense)
        out = ren(rx.to(DEV), tx.to(DEV), dtx.to(DEV) if dtx is not None else None, azi_rand=azi)
        (out * G.to(DEV)).sum().backward()
        print(mc, dense, "out", "%.2e" % rel_l2(out, ref_out))
        rg = dict(ref_net.named_parameters())
        for n_, p in native.named_parameters():
            print("   ", n_, "%.2e" % rel_l2(p.grad, rg[n_].grad))
