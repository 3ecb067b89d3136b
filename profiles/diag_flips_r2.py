"""Which side moves when the tensor-core path and the fp32 oracle disagree by > 1e-4 on a hash-table gradient?

For the reduced-ray real-field cases that failed the 1e-4 bar in round 2's first GPU run, evaluates the same problem with
  o32   the fp32 CPU oracle                      o64   the oracle with float64 dense layers (geometry / hash cells fp32)
  tc    the shipped tensor-core path             tc+guard  ... with the near-zero guard (fp32 re-evaluation of |v| ~ 0)
  tc/b3 ... with 3-plane (24-bit) backward GEMMs tc/bf16   ... with bf16 triples instead of fp16 pairs in the signal net
  simt  the fp32 FMA path
and prints every pairwise distance on the two most sensitive gradients.  Run on the GPU box: python profiles/diag_flips_r2.py"""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avr_b200                                         # noqa: E402
from avr_b200 import fused_tc                           # noqa: E402
from avr_b200.configs import get_config                 # noqa: E402
from oracle import field_ref, render_ref                # noqa: E402

DEV = "cuda:0"


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def main():
    cases = [("simu", 16, 8, 43), ("simu", 16, 8, 41), ("real_exp_ch_emb_1", 16, 8, 41), ("meshrir", 10, 6, 45)]
    for name, n_azi, n_ele, seed in cases:
        cfg = get_config(name)
        cfg["render"]["n_azi"], cfg["render"]["n_ele"] = n_azi, n_ele
        r = cfg["render"]
        gen = torch.Generator().manual_seed(11)
        c = (r["xyz_min"] + r["xyz_max"]) / 2
        rx = (c + (torch.rand(2, 3, generator=gen) * 2 - 1) * 1.5).float()
        tx = (c + (torch.rand(2, 3, generator=gen) * 2 - 1) * 1.5).float()
        azi = torch.rand(n_azi, generator=gen)
        T = cfg["model"]["signal_output_dim"]
        G = torch.randn(2, T // 2 + 1, 2, generator=gen)
        ref = field_ref.trained_like_(field_ref.AVRModelRef(cfg["model"], seed=seed), seed=seed + 1)
        evals = {}
        for key, net, g in (("o32", copy.deepcopy(ref), G), ("o64", copy.deepcopy(ref).double(), G.double())):
            out = render_ref.RenderRef(net, **r)(rx, tx, None, azi_rand=azi)
            (out * g).sum().backward()
            evals[key] = {n: p.grad.float() for n, p in net.named_parameters()}
        native = avr_b200.AVRModel(cfg["model"])
        native.load_state_dict(ref.state_dict())
        native = native.to(DEV)

        def run(dense="tc", **patch):
            saved = {k: getattr(fused_tc, k) for k in patch}
            for k, v in patch.items():
                setattr(fused_tc, k, v)
            try:
                native.zero_grad(set_to_none=True)
                ren = avr_b200.AVRRender(native, **r, dense=dense)
                out = ren(rx.to(DEV), tx.to(DEV), azi_rand=azi)
                (out * G.to(DEV)).sum().backward()
                torch.cuda.synchronize()
                return {n: p.grad.detach().cpu().clone() for n, p in native.named_parameters()}
            finally:
                for k, v in saved.items():
                    setattr(fused_tc, k, v)

        evals["tc"] = run()
        evals["tc+guard"] = run(NEAR_ZERO_GUARD=True)
        evals["tc/b3"] = run(BWD_PLANES=3)
        evals["tc/bf16"] = run(SIG_HIDDEN_F16=False)
        evals["simt"] = run(dense="simt")
        keys = list(evals)
        for pname in ("_pos_encoding.params", "_dir_encoding.params", "_model_decoder_sigma.params"):
            print(f"\n{name} {n_azi}x{n_ele} seed {seed}  {pname}: rel-L2 between evaluations (row vs column)")
            print(" " * 10 + "".join(f"{k:>10s}" for k in keys))
            for a in keys:
                print(f"{a:>10s}" + "".join(f"{rel(evals[a][pname], evals[b][pname]):10.1e}" for b in keys))


if __name__ == "__main__":
    main()
