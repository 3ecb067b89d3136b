"""Per-parameter gradient errors of the channel-embedding variants vs the oracle and vs the oracle's own fp32 noise
(run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avr_b200
from avr_b200.configs import tiny_config
from oracle import field_ref, render_ref
from oracle.render_ref import rel_l2
from tests.helpers import oracle_fp32_noise
DEV = "cuda:0"
combos = [("concat", 1, 0, 0), ("add", 1, 1, 1)]
for conn, enc, dec, sig in combos:
    cfg = tiny_config("AVRModel", n_azi=12, n_ele=6, n_samples=24, T=400, width_sigma=64, width_signal=128)
    if conn != "none":
        cfg["model"]["channel_embed"] = {"is_embed": True, "ch_num": 8, "connection_type": conn, "is_sigma_encoder": bool(enc),
                                         "is_sigma_decoder": bool(dec), "is_signal_network": bool(sig), "emb_dim_sigma_encoder": 8,
                                         "emb_dim_sigma_decoder": 16, "emb_dim_signal_network": 24}
    ref_net = field_ref.trained_like_(field_ref.AVRModelRef(cfg["model"], seed=21), seed=22)
    native = avr_b200.AVRModel(cfg["model"]); native.load_state_dict(ref_net.state_dict()); native.to(DEV)
    r = cfg["render"]; bs = 3
    gen = torch.Generator().manual_seed(5)
    rx = ((torch.rand(bs, 3, generator=gen) * 2 - 1) * 2).float(); tx = ((torch.rand(bs, 3, generator=gen) * 2 - 1) * 2).float()
    ch = torch.tensor([5, 0, 5]) if conn != "none" else None
    azi = torch.rand(r["n_azi"], generator=gen); G = torch.randn(bs, 201, 2, generator=gen)
    kw = dict(azi_rand=azi) if ch is None else dict(ch_idx=ch, azi_rand=azi)
    ref_out = render_ref.RenderRef(ref_net, **r)(rx, tx, None, **kw)
    (ref_out * G).sum().backward()
    n_out, noise = oracle_fp32_noise(ref_net, r, rx, tx, G, **kw)
    ren = avr_b200.AVRRender(native, **r)
    out = ren(rx.to(DEV), tx.to(DEV), ch_idx=ch.to(DEV) if ch is not None else None, azi_rand=azi)
    (out * G.to(DEV)).sum().backward()
    print(conn, enc, dec, sig, "out", f"{rel_l2(out.cpu(), ref_out):.2e} noise {n_out:.2e}")
    refg = dict(ref_net.named_parameters())
    tc = {n: p.grad.clone() for n, p in native.named_parameters()}
    native.zero_grad(set_to_none=True)

    class Plain(torch.nn.Module):                                  # no fused_plan: generic path, exact-fp32 SIMT GEMMs
        def __init__(self, net):
            super().__init__(); self.net = net
        def forward(self, pts, view, tx, ch_idx=None):
            return self.net(pts, view, tx, ch_idx=ch_idx)
    out_g = avr_b200.AVRRender(Plain(native), **r)(rx.to(DEV), tx.to(DEV), ch_idx=ch.to(DEV), azi_rand=azi)
    (out_g * G.to(DEV)).sum().backward()
    print("   generic-path out", f"{rel_l2(out_g.cpu(), ref_out):.2e}")
    for n, p in native.named_parameters():
        print(f"   {n:55s} tc {rel_l2(tc[n].cpu(), refg[n].grad):.2e}  generic fp32 {rel_l2(p.grad.cpu(), refg[n].grad):.2e}  oracle noise {noise[n]:.2e}")
