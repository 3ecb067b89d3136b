"""Small host-side tables of the renderer (SURVEY 8b "small-table policy").

O(R + S*T + T*F) values that depend only on the render config (plus, for the direction table, on the
per-call azimuth jitter).  They are evaluated on the CPU with the same torch expressions the reference
uses so that they are bit-identical to ``renderer_cpu.py`` (torch's vectorised ``linspace`` / ``cos`` /
``acos`` / complex ``exp`` are not reproducible by hand-rolled device code), then uploaded once and
cached.  All per-(b,r,s[,t]) arithmetic lives in the CUDA kernels.
"""
from __future__ import annotations

import numpy as np
import torch


def direction_table(n_azi: int, n_ele: int, azi_rand: torch.Tensor | None = None) -> torch.Tensor:
    """Ray directions ``[n_azi*n_ele+2, 3]`` (renderer.py:133-165): jittered azimuth grid x equal-area
    elevation grid, ray ``r = a*n_ele + e``, the two poles appended.

    ``azi_rand=None`` draws ``rand(n_azi)`` and then the (unused) ``rand(n_ele)`` from the global CPU
    generator, exactly like the reference does on every forward (also in eval).
    """
    if azi_rand is None:
        azi_rand = torch.rand(n_azi)
        torch.rand(n_ele)
    elif not torch.is_tensor(azi_rand):             # a plain sequence of floats (nn.DataParallel scatters tensors)
        azi_rand = torch.tensor(list(azi_rand), dtype=torch.float32)
    azi = torch.linspace(0, np.pi * 2, n_azi + 1)[:-1] + (np.pi * 2 / n_azi) * azi_rand.to("cpu", torch.float32)
    ele = torch.acos(2 * torch.linspace(0, 1, n_ele + 2)[1:-1] - 1)
    azi = azi[:, None].expand(n_azi, n_ele).reshape(-1)
    ele = ele[None, :].expand(n_azi, n_ele).reshape(-1)
    out = torch.empty(n_azi * n_ele + 2, 3)
    out[:-2, 0] = torch.mul(torch.cos(azi), torch.sin(ele))
    out[:-2, 1] = torch.mul(torch.sin(azi), torch.sin(ele))
    out[:-2, 2] = torch.cos(ele)
    out[-2] = torch.tensor([0.0, 0.0, 1.0])
    out[-1] = torch.tensor([0.0, 0.0, -1.0])
    return out


def dft_matrix(T: int) -> torch.Tensor:
    """``[T, ldd]`` real-DFT matrix, columns (cos, -sin)(2 pi f t / T) interleaved, ``ldd = ceil8(2F)``.

    Angles are reduced exactly (``f*t mod T`` in integers) and evaluated in float64 before rounding.
    """
    F = T // 2 + 1
    ldd = (2 * F + 7) // 8 * 8
    k = (torch.arange(T, dtype=torch.int64)[:, None] * torch.arange(F, dtype=torch.int64)[None, :]) % T
    ang = k.double() * (2.0 * np.pi / T)
    m = torch.zeros(T, ldd, dtype=torch.float64)
    m[:, 0:2 * F:2] = torch.cos(ang)
    m[:, 1:2 * F:2] = -torch.sin(ang)
    return m.float()


class RenderTables:
    """Static tables of one ``(render config, T)`` pair on one device."""

    def __init__(self, cfg: dict, T: int, device):
        S, near, far = int(cfg["n_samples"]), cfg["near"], cfg["far"]
        fs, speed = cfg["fs"], cfg["speed"]
        F = T // 2 + 1
        d = torch.linspace(0., 1., S) * (far - near) + near                       # renderer.py:54
        tau = fs * d / speed                                                      # :80
        shift = torch.round(tau)                                                  # :81
        prev = int(0.1 / speed * fs)                                              # :96
        pl = cfg["pathloss"] / (torch.arange(0, T * 2.5) / fs * speed + 1e-3)     # :97-98
        pl[0:prev] = pl[prev + 1]                                                 # :99
        t = torch.arange(T)
        idx = shift.long()[:, None] + t[None, :]
        if int(idx.max()) >= pl.numel():
            raise ValueError("far*fs/speed too large for the IR length: the reference path-loss table "
                             "(2.5*T entries, renderer.py:97-100) would be overrun")
        tail_mask = (torch.arange(T - 1, -1, -1)[None, :] - shift[:, None]) > 0  # :82
        gain = torch.where(tail_mask, pl[idx], torch.zeros(()))                   # :82-83 x :100
        phase = torch.exp(-1j * 2 * np.pi / T * torch.arange(0, F).unsqueeze(0) * tau.unsqueeze(1))   # :108
        delta = torch.cat([d[1:] - d[:-1], torch.tensor([1e10])])                # :185-186
        self.T, self.F, self.S = T, F, S
        #: spacing of a ray's samples in unit-cube coordinates (avr_raygen_encode_bwd's run-merging scatter)
        self.sample_step = float(far - near) / max(1, S - 1) / float(cfg["xyz_max"] - cfg["xyz_min"])
        self.host = {"d": d, "tau": tau, "shift": shift, "pl": pl, "delta": delta}
        self.dev = {
            "d": d.to(device), "delta": delta.to(device),
            "gain": gain.float().contiguous().to(device),
            "phase": torch.view_as_real(phase.to(torch.complex64)).contiguous().to(device),
            "dft": dft_matrix(T).to(device),
            "sample_step": self.sample_step,
        }
