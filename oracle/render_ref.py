"""fp32 torch restatement of the reference CPU renderer.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows ``/root/reference/renderer_cpu.py``:

=====================  ==========================================================
here                   reference
=====================  ==========================================================
``direction_table``    ``ray_directions``            renderer_cpu.py:111-143
``static_tables``      d_vals :46, tau/shift :69-70, path loss :82-86,
                       phase :91, interval widths :163-164
``sample_geometry``    ray points :47, normalise :50-52,105-106
``source_delay``       tx->point delay indices :76-77,108-109
``ray_weights``        ``acoustic_render`` alpha / transmittance  :159-168
``composite``          masks :72-80, rfft*phase :90-92, weighted sums :95-101,170
``RenderRef``          ``AVRRender``                 renderer_cpu.py:5-102
=====================  ==========================================================

PINNED: ``tests/test_oracle_vs_reference.py`` runs the unmodified reference module beside this
file (bit-exact geometry / delay indices, rendered IR to fp32 round-off) whenever
``/root/reference`` exists, and ``tests/test_oracle_golden.py`` checks it everywhere against
``tests/golden/*.npz`` (generated from the unmodified reference by ``oracle/make_golden.py``).

The only intended differences from the reference are structural: the azimuth jitter can be
injected (``azi_rand``) instead of always being drawn from the global CPU generator, and
``ch_idx`` is accepted (the GPU twin ``renderer.py:31`` has it, the CPU twin does not).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn


def direction_table(n_azi: int, n_ele: int, azi_rand: torch.Tensor | None = None,
                    random_azi: bool = True) -> torch.Tensor:
    """Unit ray directions ``[n_azi*n_ele + 2, 3]``; ray ``r = a*n_ele + e``, poles last.

    With ``azi_rand=None`` the two generator draws of the reference (``rand(n_azi)`` then the
    unused ``rand(n_ele)``) are made, in that order, from the global CPU generator.
    """
    azi = torch.linspace(0, np.pi * 2, n_azi + 1)[:-1]
    if azi_rand is None:
        azi_rand = torch.rand(n_azi)
        torch.rand(n_ele)                                   # drawn and multiplied by 0 upstream
    if random_azi:
        azi = azi + (np.pi * 2 / n_azi) * azi_rand
    ele = torch.acos(2 * torch.linspace(0, 1, n_ele + 2)[1:-1] - 1)
    azi_g, ele_g = torch.meshgrid(azi, ele, indexing="ij")
    azi_f, ele_f = azi_g.flatten(), ele_g.flatten()
    sin_e = torch.sin(ele_f)
    body = torch.stack([torch.cos(azi_f) * sin_e, torch.sin(azi_f) * sin_e, torch.cos(ele_f)], dim=1)
    poles = torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, -1.0]])
    return torch.cat([body, poles], dim=0)


def static_tables(cfg: dict, T: int) -> dict:
    """Everything that depends only on the render config and the IR length ``T``."""
    S, near, far = cfg["n_samples"], cfg["near"], cfg["far"]
    fs, speed = cfg["fs"], cfg["speed"]
    F = T // 2 + 1
    d = torch.linspace(0., 1., S) * (far - near) + near
    tau = fs * d / speed                                    # fractional rx delay in samples
    shift = torch.round(tau)                                # half-to-even
    prev = int(0.1 / speed * fs)
    pl = cfg["pathloss"] / (torch.arange(0, T * 2.5) / fs * speed + 1e-3)
    pl[0:prev] = pl[prev + 1]
    phase = torch.exp(-1j * 2 * np.pi / T * torch.arange(0, F).unsqueeze(0) * tau.unsqueeze(1))
    delta = torch.cat([d[1:] - d[:-1], torch.tensor([1e10])])
    return {"d": d, "tau": tau, "shift": shift, "pl": pl, "phase": phase, "delta": delta,
            "prev": prev, "T": T, "F": F}


def _to_unit(x, lo, hi):
    return 2 * (x - lo) / (hi - lo) - 1


def _from_unit(x, lo, hi):
    return (x + 1) / 2 * (hi - lo) + lo


def sample_geometry(rays_o, position_tx, dirs, d, cfg):
    """-> (pts_n[bs,P,3], view[bs,P,3], tx_n[bs,P,3], ray_pts[bs,R,S,3]); point ``p = r*S + s``."""
    lo, hi = cfg["xyz_min"], cfg["xyz_max"]
    bs, R, S = rays_o.size(0), dirs.size(0), d.numel()
    ray_pts = rays_o[:, None, None, :] + (dirs[:, None, :] * d[None, :, None])[None]
    pts_n = _to_unit(ray_pts.reshape(bs, -1, 3), lo, hi)
    view = -1 * dirs[None, :, None, :].expand(bs, R, S, 3).reshape(bs, -1, 3)
    tx_n = _to_unit(position_tx[:, None, :].expand(bs, R * S, 3), lo, hi)
    return pts_n, view, tx_n, ray_pts


def source_delay(pts_n, tx_n, cfg, T, S):
    """tx -> sample-point delay in whole samples, ``[bs,R,S]`` (fp32-valued integers).

    NB: de-normalising a *difference* adds (hi+lo)/2 per axis -- kept bug-for-bug (:76).
    """
    lo, hi = cfg["xyz_min"], cfg["xyz_max"]
    bs = pts_n.size(0)
    dist = torch.linalg.vector_norm(_from_unit(tx_n - pts_n, lo, hi), dim=-1).reshape(bs, -1, S)
    return torch.clamp(torch.round(dist * cfg["fs"] / cfg["speed"]), min=0, max=T - 1)


def ray_weights(attn, delta):
    """alpha = 1-exp(-sigma*delta); w = alpha * prod_{j<s}(1-alpha_j+1e-6)."""
    alpha = 1. - torch.exp(-attn * delta)
    ones = torch.ones_like(alpha[..., :1])
    trans = torch.cumprod(torch.cat([ones, 1. - alpha + 1e-6], -1), -1)[..., :-1]
    return trans * alpha, alpha, trans


def composite(attn, signal, delay, tab):
    """Literal (reference-order) compositing.  attn[bs,R,S], signal[bs,R,S,T] -> [bs,F,2]."""
    T = tab["T"]
    t = torch.arange(T)
    tail = torch.where((torch.arange(T - 1, -1, -1)[None, :] - tab["shift"][:, None]) > 0, 1, 0)
    sig = signal * tail
    sig = sig * (t >= delay.unsqueeze(-1))
    idx = tab["shift"].numpy().astype(int)
    pl_all = torch.stack([tab["pl"][i:i + T] for i in idx])
    spec = torch.fft.rfft(sig.float() * pl_all, dim=-1) * tab["phase"]
    w, _, _ = ray_weights(attn, tab["delta"])
    per_ray = torch.sum(spec * w[..., None], -2)
    rec = torch.sum(per_ray, dim=-2)
    return torch.stack([rec.real, rec.imag], dim=-1)


def composite_reordered(attn, signal, delay, tab):
    """Same result with the ray sum moved in front of the FFT (last block).

    This is the order the CUDA kernels use; kept here so the reassociation error (~3e-7 rel-L2)
    is measured on the CPU, independent of any kernel.
    """
    T = tab["T"]
    t = torch.arange(T)
    w, _, _ = ray_weights(attn, tab["delta"])
    y = torch.sum(signal * (w[..., None] * (t >= delay.unsqueeze(-1))), dim=1)      # [bs,S,T]
    m1 = (t[None, :] < (T - 1 - tab["shift"][:, None])).float()
    idx = (tab["shift"].long()[:, None] + t[None, :])
    z = y * m1 * tab["pl"][idx]
    rec = torch.sum(torch.fft.rfft(z, dim=-1) * tab["phase"], dim=1)
    return torch.stack([rec.real, rec.imag], dim=-1)


class RenderRef(nn.Module):
    """Drop-in for ``renderer_cpu.AVRRender`` (same ctor kwargs / forward contract)."""

    KEYS = ("n_samples", "near", "far", "n_azi", "n_ele", "speed", "fs", "pathloss", "xyz_min", "xyz_max")

    def __init__(self, networks_fn, **kwargs):
        super().__init__()
        self.network_fn = networks_fn
        self.cfg = {k: kwargs[k] for k in self.KEYS}
        self.reordered = False

    def forward(self, rays_o, position_tx, direction_tx=None, ch_idx=None, azi_rand=None):
        cfg = self.cfg
        bs = position_tx.size(0)
        dirs = direction_table(cfg["n_azi"], cfg["n_ele"], azi_rand)
        d = torch.linspace(0., 1., cfg["n_samples"]) * (cfg["far"] - cfg["near"]) + cfg["near"]
        if direction_tx is not None:
            n_pts = dirs.size(0) * d.numel()
            dir_tx = direction_tx[:, None, :].expand(bs, n_pts, 3)
        pts_n, view, tx_n, _ = sample_geometry(rays_o, position_tx, dirs, d, cfg)
        kw = {"ch_idx": ch_idx} if ch_idx is not None else {}           # renderer.py:68-73 (the GPU twin passes it on)
        if direction_tx is not None:
            attn, signal = self.network_fn(pts_n, view, tx_n, dir_tx, **kw)
        else:
            attn, signal = self.network_fn(pts_n, view, tx_n, **kw)
        S = cfg["n_samples"]
        attn = attn.view(bs, -1, S)
        signal = signal.view(bs, -1, S, signal.size(-1))
        tab = static_tables(cfg, signal.size(-1))
        delay = source_delay(pts_n, tx_n, cfg, signal.size(-1), S)
        fn = composite_reordered if self.reordered else composite
        return fn(attn, signal, delay, tab)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| in float64."""
    a, b = a.detach().double(), b.detach().double()
    den = float(torch.linalg.vector_norm(b))
    return float(torch.linalg.vector_norm(a - b)) / (den if den > 0 else 1.0)
