"""Parity of the CUDA path against the committed golden vectors (unmodified reference renderer) -- a table
of relative-L2 errors for the rendered IR and every parameter gradient, both dense-layer modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, avr_b200
from tests.helpers import GOLDEN_CASES, case_config, load_golden, oracle_field, rel_l2
DEV = "cuda:0"
print("| case | dense | IR rel-L2 | worst grad rel-L2 (tensor) |")
print("|---|---|---:|---|")
for name in GOLDEN_CASES[:3]:
    g = load_golden(name)
    mc, cfg = case_config(name)
    ref_net = oracle_field(mc, cfg["model"], g)
    for dense in ("tc", "simt"):
        cls = avr_b200.AVRModel if mc == "AVRModel" else avr_b200.AVRModel_complex
        net = cls(cfg["model"]); net.load_state_dict(ref_net.state_dict()); net = net.to(DEV)
        ren = avr_b200.AVRRender(net, **cfg["render"], dense=dense)
        dtx = g["dir_tx"].to(DEV) if "dir_tx" in g else None
        out = ren(g["rx"].to(DEV), g["tx"].to(DEV), dtx, azi_rand=g["azi_rand"])
        (out * g["G"].to(DEV)).sum().backward()
        errs = {n: rel_l2(p.grad, g["grad/" + n]) for n, p in net.named_parameters()}
        worst = max(errs, key=errs.get)
        print(f"| {name} | {dense} | {rel_l2(out, g['out']):.2e} | {errs[worst]:.2e} ({worst}) |")
