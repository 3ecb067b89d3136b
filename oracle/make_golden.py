"""Regenerate ``tests/golden/*.npz`` from the UNMODIFIED reference renderer.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Run in the build container only:

    python -m oracle.make_golden

Every vector below is produced by ``/root/reference/renderer_cpu.py`` (``AVRRender.forward``,
``ray_directions``) executed as it lies; the field network it calls is ``oracle/field_ref.py``
(tiny-cuda-nn itself cannot be installed -- see that file's header).  Captured per case:

* inputs: receiver / transmitter positions, transmitter direction, the azimuth jitter the
  reference drew (recovered by replaying the CPU generator from the stored seed), all
  parameter tensors, the cotangent ``G`` used for the backward pass;
* what the reference handed to the network: ``pts``, ``view``, ``tx`` (bit-exact targets for
  the ray-generation / sampling kernel);
* the network outputs ``attn``, ``signal``; the rendered ``out[bs,F,2]``;
* ``d(sum(out*G))/d(params)`` from torch autograd through the reference.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from avr_b200.configs import tiny_config          # noqa: E402
from oracle import field_ref                      # noqa: E402
from oracle.reference_shim import load_reference_renderer  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class _Tap(torch.nn.Module):
    """Records what the reference passes to / gets from the network."""

    def __init__(self, net):
        super().__init__()
        self.net = net
        self.seen = {}

    def forward(self, pts, view, tx, dir_tx=None):
        self.seen = {"pts": pts.detach().clone(), "view": view.detach().clone(), "tx": tx.detach().clone()}
        if dir_tx is not None:
            self.seen["dir_tx"] = dir_tx.detach().clone()
            attn, sig = self.net(pts, view, tx, dir_tx)
        else:
            attn, sig = self.net(pts, view, tx)
        self.seen["attn"], self.seen["signal"] = attn.detach().clone(), sig.detach().clone()
        return attn, sig


class StubField(torch.nn.Module):
    """Renderer-only case: attn / signal are free parameters (isolates renderer_cpu.py)."""

    def __init__(self, bs, n_pts, T, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.signal_output_dim = T
        self.attn = torch.nn.Parameter(torch.rand(bs, n_pts, 1, generator=g) * 0.8)
        self.signal = torch.nn.Parameter(torch.randn(bs, n_pts, T, generator=g))

    def forward(self, pts, view, tx, dir_tx=None):
        return self.attn, self.signal


CASES = {
    # name: (model_class, tiny_config kwargs, bs, jitter seed, centre half-width)
    "avrmodel_sym": ("AVRModel", dict(), 2, 11, 3.0),
    "avrmodel_box0_10": ("AVRModel", dict(xyz_min=0, xyz_max=10, fs=4000), 2, 12, 2.0),
    "avrmodel_complex": ("AVRModel_complex", dict(xyz_min=-12, xyz_max=12, speed=346.8, pathloss=0.5, fs=8000), 3, 13, 2.0),
    "stub_renderer_only": ("stub", dict(n_azi=5, n_ele=4, n_samples=7, T=240), 2, 14, 3.0),
}


def build_case(name):
    model_class, kw, bs, seed, half = CASES[name]
    cfg = tiny_config("AVRModel" if model_class == "stub" else model_class, **kw)
    render, T = cfg["render"], cfg["model"]["signal_output_dim"]
    R = render["n_azi"] * render["n_ele"] + 2
    g = torch.Generator().manual_seed(seed)
    c = (render["xyz_min"] + render["xyz_max"]) / 2
    rx = (c + (torch.rand(bs, 3, generator=g) * 2 - 1) * half).float()
    tx = (c + (torch.rand(bs, 3, generator=g) * 2 - 1) * half).float()
    dir_tx = None
    if model_class == "AVRModel_complex":
        ang = torch.rand(bs, generator=g) * 2 * np.pi
        dir_tx = torch.stack([torch.cos(ang), torch.sin(ang), torch.zeros(bs)], dim=1).float()
    if model_class == "stub":
        net = StubField(bs, R * render["n_samples"], T, seed)
    elif model_class == "AVRModel":
        net = field_ref.trained_like_(field_ref.AVRModelRef(cfg["model"], seed=seed), seed=seed)
    else:
        net = field_ref.trained_like_(field_ref.AVRModelComplexRef(cfg["model"], seed=seed), seed=seed)
    G = torch.randn(bs, T // 2 + 1, 2, generator=g)
    return cfg, net, rx, tx, dir_tx, G, seed


def run_reference(name):
    ref = load_reference_renderer()
    cfg, net, rx, tx, dir_tx, G, seed = build_case(name)
    tap = _Tap(net)
    renderer = ref.AVRRender(tap, **cfg["render"])
    torch.manual_seed(seed)
    out = renderer(rx, tx, dir_tx) if dir_tx is not None else renderer(rx, tx)
    (out * G).sum().backward()
    torch.manual_seed(seed)
    azi_rand = torch.rand(cfg["render"]["n_azi"])
    torch.manual_seed(seed)
    dirs, _, _ = ref.ray_directions(cfg["render"]["n_azi"], cfg["render"]["n_ele"])
    blob = {"rx": rx, "tx": tx, "G": G, "azi_rand": azi_rand, "dirs": dirs, "out": out.detach(),
            "seed": torch.tensor(seed)}
    if dir_tx is not None:
        blob["dir_tx"] = dir_tx
    for k, v in tap.seen.items():
        if k == "signal" and name != "stub_renderer_only":
            v = v[:, ::7].contiguous()                     # keep the file small: every 7th point
        if k in ("view", "tx", "dir_tx"):
            v = v[:, :: max(1, v.shape[1] // 16)].contiguous()
        blob["net_" + k] = v
    for pname, p in net.named_parameters():
        blob["param/" + pname] = p.detach()
        blob["grad/" + pname] = p.grad.detach()
    return {k: v.numpy() for k, v in blob.items()}


CRITERION_CASES = {
    # name: (bs, T, train config).  Weights: avr_meshrir.yml:35-40 / avr_pra_1.yml:36-41; DAS terms: 8 microphones.
    "criterion_meshrir_w": (4, 1600, {"spec_loss_weight": 1, "amplitude_loss_weight": 0.5, "angle_loss_weight": 0.5,
                                      "time_loss_weight": 100, "energy_loss_weight": 5, "multistft_loss_weight": 1}),
    "criterion_small": (3, 400, {"spec_loss_weight": 2, "amplitude_loss_weight": 4, "angle_loss_weight": 1,
                                 "time_loss_weight": 50, "energy_loss_weight": 1, "multistft_loss_weight": 1}),
    "criterion_das": (8, 1600, {"spec_loss_weight": 1, "amplitude_loss_weight": 0.5, "angle_loss_weight": 0.5,
                                "time_loss_weight": 100, "energy_loss_weight": 5, "multistft_loss_weight": 1,
                                "das_reg_loss_weight": 0.3, "das_ce_loss_weight": 0.2, "beta": 100.0}),
}
CRITERION_RENDER = {"fs": 16000, "speed": 343.8}


def run_reference_criterion(name):
    """The unmodified ``utils/criterion.py`` (stand-in auraloss = oracle restatement, see reference_shim) on seeded
    spectra shaped like rendered IRs: a decaying random time signal -> rfft."""
    from oracle.reference_shim import load_reference_criterion
    bs, T, cfg = CRITERION_CASES[name]
    crit = load_reference_criterion().Criterion(cfg, CRITERION_RENDER)
    g = torch.Generator().manual_seed(len(name) + bs)
    env = torch.exp(-torch.arange(T) / (0.15 * T))
    ori = torch.fft.rfft(torch.randn(bs, T, generator=g) * env)
    pred = torch.fft.rfft((torch.randn(bs, T, generator=g) * 0.7 + 0.3 * torch.fft.irfft(ori)) * env)
    pred = pred.to(torch.complex64).requires_grad_()
    ori = ori.to(torch.complex64)
    outs = crit(pred, ori)
    coef = torch.tensor([1.0, 0.9, 1.1, 0.8, 1.2, 0.7, 1.3, 0.6])              # distinct cotangents per loss term
    (torch.stack([o.float() for o in outs[:8]]) * coef).sum().backward()
    blob = {"pred": torch.view_as_real(pred.detach()), "ori": torch.view_as_real(ori), "coef": coef,
            "losses": torch.stack([o.detach().float() for o in outs[:8]]),
            "ori_time": outs[8].detach(), "pred_time": outs[9].detach(),
            "grad_pred": torch.view_as_real(pred.grad)}
    return {k: v.numpy() for k, v in blob.items()}


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name in CRITERION_CASES:
        blob = run_reference_criterion(name)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, losses={np.round(blob['losses'], 4)}")
    for name in CASES:
        blob = run_reference(name)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, |out|max={np.abs(blob['out']).max():.3e}")


if __name__ == "__main__":
    main()
