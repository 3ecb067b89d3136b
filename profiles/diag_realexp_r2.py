"""Localise the 1e-3 disagreement of BOTH GPU paths with the oracle on real_exp_ch_emb_1 (16 x 8 rays, seed 41).

Evaluates the same problem through (a) the oracle, (b) the generic path of AVRRender (explicit points -> standalone
field: hash-grid gather/scatter on explicit points + fp32 SIMT GEMMs -> composite kernels), capturing d(attn) and
d(signal), (c) dense='simt', (d) dense='tc'; prints every parameter's distance to the oracle and the distance of the
captured network-output gradients.  Run on the GPU box: python profiles/diag_realexp_r2.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avr_b200                                         # noqa: E402
from avr_b200.configs import get_config                 # noqa: E402
from oracle import field_ref, render_ref                # noqa: E402

DEV = "cuda:0"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


class Tap(torch.nn.Module):
    """Wraps a field so that AVRRender takes its generic path; keeps the network outputs to read their gradients."""

    def __init__(self, field):
        super().__init__()
        self.field = field
        self.signal_dim = field.signal_output_dim

    def forward(self, pts, view, tx, ch_idx=None):
        attn, signal = self.field(pts, view, tx)
        attn.retain_grad(); signal.retain_grad()
        self.attn, self.signal = attn, signal
        return attn, signal


def main():
    for name, seed, spread in (("real_exp_ch_emb_1", 41, 1.5), ("real_exp_ch_emb_1", 43, 1.5), ("simu", 41, 1.5)):
        cfg = get_config(name)
        cfg["render"]["n_azi"], cfg["render"]["n_ele"] = 16, 8
        r = cfg["render"]
        gen = torch.Generator().manual_seed(11)
        c = (r["xyz_min"] + r["xyz_max"]) / 2
        rx = (c + (torch.rand(2, 3, generator=gen) * 2 - 1) * spread).float()
        tx = (c + (torch.rand(2, 3, generator=gen) * 2 - 1) * spread).float()
        azi = torch.rand(16, generator=gen)
        T = cfg["model"]["signal_output_dim"]
        G = torch.randn(2, T // 2 + 1, 2, generator=gen)
        ref = field_ref.trained_like_(field_ref.AVRModelRef(cfg["model"], seed=seed), seed=seed + 1)
        tap_ref = Tap(ref)
        tap_ref.signal_output_dim = T
        out = render_ref.RenderRef(tap_ref, **r)(rx, tx, None, azi_rand=azi)
        (out * G).sum().backward()
        ref_g = {n: p.grad for n, p in ref.named_parameters()}
        ref_dattn, ref_dsig = tap_ref.attn.grad, tap_ref.signal.grad
        native = avr_b200.AVRModel(cfg["model"])
        native.load_state_dict(ref.state_dict())
        native = native.to(DEV)
        print(f"\n=== {name} 16x8 seed {seed}: IR norm {float(out.norm()):.3e}")
        res = {}
        # generic path
        native.zero_grad(set_to_none=True)
        tap = Tap(native)
        o = avr_b200.AVRRender(tap, **r)(rx.to(DEV), tx.to(DEV), azi_rand=azi)
        (o * G.to(DEV)).sum().backward()
        print(f"generic: IR {rel(o, out):.1e}  d_attn {rel(tap.attn.grad.reshape(-1), ref_dattn.reshape(-1)):.1e}  d_signal "
              f"{rel(tap.signal.grad.reshape(-1), ref_dsig.reshape(-1)):.1e}  attn {rel(tap.attn.reshape(-1), tap_ref.attn.reshape(-1)):.1e} "
              f"signal {rel(tap.signal.reshape(-1), tap_ref.signal.reshape(-1)):.1e}")
        res["generic"] = {n: p.grad.detach().cpu().clone() for n, p in native.named_parameters()}
        for dense in ("simt", "tc"):
            native.zero_grad(set_to_none=True)
            o = avr_b200.AVRRender(native, **r, dense=dense)(rx.to(DEV), tx.to(DEV), azi_rand=azi)
            (o * G.to(DEV)).sum().backward()
            res[dense] = {n: p.grad.detach().cpu().clone() for n, p in native.named_parameters()}
        for n in ref_g:
            print(f"  {n:30s} |g| {float(ref_g[n].norm()):.2e}  " + "  ".join(f"{k} {rel(res[k][n], ref_g[n]):.1e}" for k in res) +
                  f"   simt-vs-generic {rel(res['simt'][n], res['generic'][n]):.1e}  tc-vs-generic {rel(res['tc'][n], res['generic'][n]):.1e}")


if __name__ == "__main__":
    main()
