"""Bit-exactness of sample positions / delay indices vs the oracle on a seeded case (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avr_b200 import ops, tables
from avr_b200.configs import tiny_config
from oracle import render_ref
DEV = "cuda:0"
cfg = tiny_config("AVRModel", n_azi=10, n_ele=5, n_samples=16, T=320, width_sigma=64, width_signal=128)
r = cfg["render"]; bs = 3; T = 320
gen = torch.Generator().manual_seed(9)
rx, tx = torch.randn(bs, 3, generator=gen), torch.randn(bs, 3, generator=gen)
azi = torch.rand(r["n_azi"], generator=gen)
dirs = render_ref.direction_table(r["n_azi"], r["n_ele"], azi)
d = torch.linspace(0., 1., r["n_samples"]) * (r["far"] - r["near"]) + r["near"]
pts_n, view, tx_n, _ = render_ref.sample_geometry(rx, tx, dirs, d, r)
delay_ref = render_ref.source_delay(pts_n, tx_n, r, T, r["n_samples"])
geom = ops.make_geom(r, bs, T)
p, v, t, delay = ops.sample_points(geom, rx.to(DEV), tx.to(DEV), dirs.to(DEV), d.to(DEV))
print("pts equal", torch.equal(p.cpu().view_as(pts_n), pts_n), "tx equal", torch.equal(t.cpu().view_as(tx_n), tx_n))
dr = delay_ref.reshape(-1).long(); dg = delay.cpu().reshape(-1).long()
bad = (dr != dg).nonzero().flatten()
print("delay mismatches", bad.numel(), "of", dr.numel())
span = (r["xyz_max"] - r["xyz_min"]) / 2; mid = (r["xyz_max"] + r["xyz_min"]) / 2
for i in bad[:5].tolist():
    diff = (tx_n.reshape(-1, 3)[i] - pts_n.reshape(-1, 3)[i])
    den = diff * span + mid
    n32 = torch.norm(den)
    n64 = torch.norm(den.double())
    print(i, "ref", int(dr[i]), "gpu", int(dg[i]), "dist*fs/speed f32", float(n32 * r["fs"] / r["speed"]), "f64", float(n64) * r["fs"] / r["speed"],
          "den", den.tolist())
    # ways of forming the norm in fp32
    x, y, z = [torch.tensor(float(c)) for c in den]
    import numpy as np
    a, b, c = np.float32(den[0]), np.float32(den[1]), np.float32(den[2])
    s1 = np.float32(np.float32(a * a) + np.float32(b * b)) + np.float32(c * c)
    print("   plain sum sq", float(np.sqrt(np.float32(s1))) , "torch.norm", float(n32), "linalg", float(torch.linalg.vector_norm(den)),
          "batched", float(torch.norm(den.view(1, 3).expand(8, 3).contiguous(), dim=-1)[0]))
