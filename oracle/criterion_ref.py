"""fp32 torch restatement of the reference training loss (``/root/reference/utils/criterion.py:7-126``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

* ``CriterionRef.forward`` follows ``utils/criterion.py:69-126`` line by line (irfft, the n_fft=256 STFT energy-decay
  term with its doubled squaring, the spectral / amplitude / angle / time L1 terms, the multi-resolution STFT term
  called as ``(ori, pred)``, the optional delay-and-sum terms).  PINNED: ``tests/test_oracle_vs_reference.py``
  executes the unmodified ``utils/criterion.py`` (with the stand-in ``auraloss`` below) and compares every output.
* ``MultiResolutionSTFTLossRef`` restates ``auraloss.freq.MultiResolutionSTFTLoss`` / ``STFTLoss`` as published in
  auraloss 0.4.0 (``requirements.txt`` does not pin a version; the package is absent here and cannot be installed):
  hann windows, ``|STFT| = sqrt(clamp(re^2 + im^2, 1e-8))``, spectral convergence ``||y - x||_F / ||y||_F`` over the
  whole batch + L1 of log magnitudes + ``w_lin_mag`` x L1 of magnitudes, averaged over the resolutions.
  PARITY UNPINNED at the auraloss boundary.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class STFTLossRef(nn.Module):
    def __init__(self, fft_size, hop_size, win_length, w_sc=1.0, w_log_mag=1.0, w_lin_mag=0.0, eps=1e-8):
        super().__init__()
        self.fft_size, self.hop_size, self.win_length = fft_size, hop_size, win_length
        self.w_sc, self.w_log_mag, self.w_lin_mag, self.eps = w_sc, w_log_mag, w_lin_mag, eps
        self.register_buffer("window", torch.hann_window(win_length))

    def magnitude(self, x):
        s = torch.stft(x, self.fft_size, self.hop_size, self.win_length, self.window.to(x.device, x.dtype), return_complex=True)
        return torch.sqrt(torch.clamp(s.real ** 2 + s.imag ** 2, min=self.eps))

    def forward(self, inp, target):
        x_mag = self.magnitude(inp.reshape(-1, inp.size(-1)))
        y_mag = self.magnitude(target.reshape(-1, target.size(-1)))
        sc = torch.norm(y_mag - x_mag, p="fro") / torch.norm(y_mag, p="fro")
        log_mag = F.l1_loss(torch.log(x_mag), torch.log(y_mag))
        lin_mag = F.l1_loss(x_mag, y_mag)
        return self.w_sc * sc + self.w_log_mag * log_mag + self.w_lin_mag * lin_mag


class MultiResolutionSTFTLossRef(nn.Module):
    def __init__(self, fft_sizes=(1024, 2048, 512), hop_sizes=(120, 240, 50), win_lengths=(600, 1200, 240),
                 w_sc=1.0, w_log_mag=1.0, w_lin_mag=0.0, **_unused):
        super().__init__()
        self.losses = nn.ModuleList(STFTLossRef(f, h, w, w_sc, w_log_mag, w_lin_mag)
                                    for f, h, w in zip(fft_sizes, hop_sizes, win_lengths))

    def forward(self, x, y):
        total = 0.0
        for f in self.losses:
            total = total + f(x, y)
        return total / len(self.losses)


MRSTFT_KW = dict(w_lin_mag=1, fft_sizes=[512, 256, 128, 64], win_lengths=[300, 150, 75, 30], hop_sizes=[60, 30, 8, 4])   # criterion.py:33


def beamforming_power(sig, fs, speed, angles_rad):
    """``utils/criterion.py:35-67``: delay-and-sum power over 360 look directions for an 8-microphone circle."""
    m = sig.shape[0]
    assert m == 8, f"Expected 8 microphones, but got {m}"
    time_sig = torch.real(torch.fft.irfft(sig, dim=-1))
    n_fft = 512
    freqs = torch.fft.rfftfreq(n_fft, 1 / fs).to(sig.device)
    x = torch.fft.rfft(time_sig, n=n_fft, dim=-1)
    mic_angles = torch.linspace(math.pi / 2, math.pi / 2 + 2 * math.pi, m + 1)[:-1].to(sig.device)
    mic_pos = torch.stack([torch.cos(mic_angles), torch.sin(mic_angles)], dim=-1)
    mic_pos = mic_pos - mic_pos.mean(dim=0)
    u = torch.stack([torch.cos(angles_rad), torch.sin(angles_rad)], dim=-1).to(sig.device)       # [K,2]
    delays = (u @ mic_pos.t()) / speed                                                           # [K,M]
    steering = torch.exp(-1j * 2 * math.pi * delays[:, :, None] * freqs[None, None, :])
    beam = torch.einsum("mf,kmf->kf", x, steering.to(x.dtype)) / m
    power = torch.abs(beam) ** 2
    power = power / (torch.sum(power, dim=0, keepdim=True) + 1e-8)
    return torch.sum(power, dim=-1)


class CriterionRef(nn.Module):
    def __init__(self, cfg, cfg_render):
        super().__init__()
        self.w = {k: cfg[k] for k in ("spec_loss_weight", "amplitude_loss_weight", "angle_loss_weight", "time_loss_weight",
                                      "energy_loss_weight", "multistft_loss_weight")}
        self.das_reg_w, self.das_ce_w = cfg.get("das_reg_loss_weight", 0.0), cfg.get("das_ce_loss_weight", 0.0)
        self.beta = cfg.get("beta", 100.0)
        self.angles_rad = torch.deg2rad(torch.arange(0.0, 360.0, 1.0))
        self.fs, self.speed = cfg_render["fs"], cfg_render["speed"]
        self.mrstft = MultiResolutionSTFTLossRef(**MRSTFT_KW)

    @staticmethod
    def _edc(spec_energy):
        e = torch.log10(torch.flip(torch.cumsum(torch.flip(spec_energy, [-1]) ** 2, dim=-1), [-1]) + 1e-9)   # :81
        return e - e[:, [0]]

    def forward(self, pred_sig, ori_sig):
        pred_time = torch.real(torch.fft.irfft(pred_sig, dim=-1))                                # :71-72
        ori_time = torch.real(torch.fft.irfft(ori_sig, dim=-1))
        pred_spec = torch.abs(torch.stft(pred_time, n_fft=256, return_complex=True))             # :74-75
        ori_spec = torch.abs(torch.stft(ori_time, n_fft=256, return_complex=True))
        pred_energy = self._edc(torch.sum(pred_spec ** 2, dim=1))                                # :77-84
        ori_energy = self._edc(torch.sum(ori_spec ** 2, dim=1))
        l1 = F.l1_loss
        spec = (l1(pred_sig.real, ori_sig.real) + l1(pred_sig.imag, ori_sig.imag)) * self.w["spec_loss_weight"]   # :86-88
        amp = l1(torch.abs(pred_sig), torch.abs(ori_sig)) * self.w["amplitude_loss_weight"]      # :90
        pa, oa = torch.angle(pred_sig), torch.angle(ori_sig)
        ang = (l1(torch.cos(pa), torch.cos(oa)) + l1(torch.sin(pa), torch.sin(oa))) * self.w["angle_loss_weight"]   # :92-93
        time = l1(ori_time, pred_time) * self.w["time_loss_weight"]                              # :95
        energy = l1(ori_energy, pred_energy) * self.w["energy_loss_weight"]                      # :97
        mr = self.mrstft(ori_time.unsqueeze(1), pred_time.unsqueeze(1)) * self.w["multistft_loss_weight"]   # :99
        das_reg = torch.tensor(0.0, device=pred_sig.device)
        das_ce = torch.tensor(0.0, device=pred_sig.device)
        if self.das_reg_w > 0 or self.das_ce_w > 0:                                              # :105-124
            ang_tab = self.angles_rad.to(pred_sig.device)
            p_pred = beamforming_power(pred_sig, self.fs, self.speed, ang_tab)
            p_ori = beamforming_power(ori_sig, self.fs, self.speed, ang_tab)
            if self.das_ce_w > 0:
                das_ce = F.cross_entropy(p_pred.unsqueeze(0), torch.argmax(p_ori).unsqueeze(0)) * self.das_ce_w
            if self.das_reg_w > 0:
                a_pred = torch.sum(torch.softmax(self.beta * p_pred, dim=0) * ang_tab)
                a_ori = torch.sum(torch.softmax(self.beta * p_ori, dim=0) * ang_tab)
                das_reg = (l1(torch.sin(a_pred), torch.sin(a_ori)) + l1(torch.cos(a_pred), torch.cos(a_ori))) * self.das_reg_w
        return spec, amp, ang, time, energy, mr, das_reg, das_ce, ori_time, pred_time
