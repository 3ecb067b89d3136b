"""The training loss, drop-in for the reference ``utils/criterion.py``.

``Criterion(cfg_train, cfg_render)(pred_sig, ori_sig)`` takes the complex spectra ``[bs, T//2+1]`` the runners build
from the renderer output (``avr_runner.py:178-181``) and returns the reference's ten-tuple
``(spec, amplitude, angle, time, energy, multi_stft, das_reg, das_ce, ori_time, pred_time)``.

Everything up to the multi-resolution STFT term runs in the hand-written kernels of ``csrc/criterion.cu``: each
kernel emits the fixed-order partial sums of its term AND the term's gradient with respect to the predicted
spectrum, so ``backward`` is a single weighted sum of six ``[bs, F, 2]`` arrays -- the ~40 launches, four
``torch.stft`` calls and the ``auraloss`` dependency of the reference are gone.  The two optional delay-and-sum terms
(``das_*_loss_weight``; eight-microphone batches only) are short torch expressions on the GPU.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, tables
from .ops import _ctx, _p

#: (n_fft, hop, win_length) of the multi-resolution STFT term and its linear-magnitude weight (criterion.py:33)
MRSTFT_RESOLUTIONS = ((512, 60, 300), (256, 30, 150), (128, 8, 75), (64, 4, 30))
MRSTFT_W_LIN = 1.0
MRSTFT_EPS = 1e-8                  # auraloss STFTLoss eps
ENERGY_NFFT, ENERGY_HOP = 256, 64  # torch.stft(x, n_fft=256): hop = n_fft // 4, rectangular window (criterion.py:74-75)


def _call(name, *args):
    _lib.check(getattr(_lib.load(), name)(*args), name)


class _CriterionFn(torch.autograd.Function):
    """(pred[bs,F,2], ori[bs,F,2]) -> (losses[6], ori_time[bs,T], pred_time[bs,T])."""

    @staticmethod
    def forward(ctx, pred, ori, crit):
        ctx.set_materialize_grads(False)
        pred = pred.detach().contiguous().float()
        ori = ori.detach().contiguous().float()
        if not pred.is_cuda:
            raise _lib.AVRLibraryError("avr_b200.Criterion needs CUDA tensors (there is no CPU path)")
        bs, n_f, _ = pred.shape
        T = 2 * (n_f - 1)
        dev_t = pred.device
        dev, st = _ctx(pred)
        want_grad = ctx.needs_input_grad[0]
        dft = crit.dft_for(T, dev_t)
        ldd = dft.stride(0)
        w = crit.weights
        if bs == 0:
            return pred.new_zeros(6), pred.new_zeros(0, T), pred.new_zeros(0, T)

        # time signals of both spectra in one launch (criterion.py:71-72)
        both = torch.cat([pred, ori], 0)
        x = torch.empty(2 * bs, T, device=dev_t)
        _call("avr_crit_irfft", _p(both), 2 * bs, T, _p(dft), ldd, _p(x), dev, st)
        x_pred, x_ori = x[:bs].clone(), x[bs:].clone()                # separate outputs, not two views of one buffer

        # spectral / amplitude / angle L1 (criterion.py:86-93)
        n_bins = float(bs * n_f)
        grads = torch.empty(6, bs, n_f, 2, device=dev_t) if want_grad else None
        part_f = torch.empty(bs, 3, device=dev_t)
        _call("avr_crit_freq_terms", _p(pred), _p(ori), bs, n_f, w["spec"] / n_bins, w["amp"] / n_bins, w["angle"] / n_bins,
              _p(part_f), _p(grads), dev, st)
        # time L1 (criterion.py:95), energy-decay L1 (criterion.py:74-84,97)
        d_x = torch.zeros(3, bs, T, device=dev_t) if want_grad else None        # time, energy, multi-res STFT
        part_t = torch.empty(bs, device=dev_t)
        _call("avr_crit_time_l1", _p(x_pred), _p(x_ori), bs, T, w["time"] / float(bs * T), _p(part_t),
              _p(d_x[0]) if want_grad else None, dev, st)
        m_e = 1 + T // ENERGY_HOP
        nbytes = int(_lib.load().avr_crit_energy_workspace_bytes(bs, T, ENERGY_HOP))
        ws = torch.empty(max(1, nbytes // 4), device=dev_t)
        part_e = torch.empty(bs, device=dev_t)
        _call("avr_crit_energy", _p(x_pred), _p(x_ori), bs, T, ENERGY_NFFT, ENERGY_HOP, w["energy"] / float(bs * m_e), _p(part_e),
              _p(d_x[1]) if want_grad else None, _p(ws), nbytes, dev, st)
        # multi-resolution STFT (criterion.py:33,99): input = ori, target = pred
        mr = pred.new_zeros(())
        scale_r = w["mrstft"] / len(MRSTFT_RESOLUTIONS)
        for n_fft, hop, win in MRSTFT_RESOLUTIONS:
            window = crit.window_for(win, dev_t)
            m, k = 1 + T // hop, n_fft // 2 + 1
            s_pred = torch.empty(bs, m, k, 2, device=dev_t)
            mag_ori = torch.empty(bs, m, k, device=dev_t)
            part = torch.empty(bs * m, 4, device=dev_t)
            _call("avr_crit_stft_fwd", _p(x_pred), _p(x_ori), bs, T, n_fft, hop, win, _p(window), MRSTFT_EPS, _p(s_pred),
                  _p(mag_ori), _p(part), dev, st)
            sums = part.sum(0)
            n_el = float(bs * m * k)
            mr = mr + scale_r * (torch.sqrt(sums[0]) / torch.sqrt(sums[1]) + sums[2] / n_el + MRSTFT_W_LIN * sums[3] / n_el)
            if want_grad:
                frames = torch.empty(bs, m, win, device=dev_t)
                _call("avr_crit_stft_bwd", _p(s_pred), _p(mag_ori), _p(sums), bs, T, n_fft, hop, win, _p(window), MRSTFT_EPS,
                      scale_r, MRSTFT_W_LIN, _p(frames), _p(d_x[2]), dev, st)
        sf = part_f.sum(0)
        losses = torch.stack([sf[0] * (w["spec"] / n_bins), sf[1] * (w["amp"] / n_bins), sf[2] * (w["angle"] / n_bins),
                              part_t.sum() * (w["time"] / float(bs * T)), part_e.sum() * (w["energy"] / float(bs * m_e)), mr])
        if want_grad:
            # the three time-domain gradients go back through the irfft together
            _call("avr_crit_irfft_adjoint", _p(d_x), 3 * bs, T, _p(dft), ldd, _p(grads[3:]), dev, st)
            ctx.save_for_backward(grads)
            ctx.dft, ctx.T = dft, T
        ctx.mark_non_differentiable(x_ori)
        return losses, x_ori, x_pred

    @staticmethod
    def backward(ctx, d_losses, _d_ori_time, d_pred_time):
        (grads,) = ctx.saved_tensors
        d_pred = None
        if d_losses is not None:
            d_pred = torch.einsum("k,kbfc->bfc", d_losses.float(), grads)
        if d_pred_time is not None:                                  # someone differentiates the returned time signal
            dev, st = _ctx(grads)
            g = d_pred_time.contiguous().float()
            extra = torch.empty(g.shape[0], grads.shape[2], 2, device=g.device)
            _call("avr_crit_irfft_adjoint", _p(g), g.shape[0], ctx.T, _p(ctx.dft), ctx.dft.stride(0), _p(extra), dev, st)
            d_pred = extra if d_pred is None else d_pred + extra
        return d_pred, None, None


def beamforming_power(sig: torch.Tensor, fs: float, speed: float, angles_rad: torch.Tensor) -> torch.Tensor:
    """Delay-and-sum power of an eight-microphone circular array over the look directions (criterion.py:35-67)."""
    m = sig.shape[0]
    if m != 8:
        raise ValueError(f"the delay-and-sum terms expect the 8 microphones of one array per batch, got {m}")
    n_fft = 512
    spec = torch.fft.rfft(torch.fft.irfft(sig, dim=-1), n=n_fft, dim=-1)                       # [M, 257]
    freqs = torch.fft.rfftfreq(n_fft, 1 / fs).to(sig.device)
    phi = torch.linspace(math.pi / 2, math.pi / 2 + 2 * math.pi, m + 1)[:-1].to(sig.device)
    mic = torch.stack([torch.cos(phi), torch.sin(phi)], -1)
    mic = mic - mic.mean(0)
    look = torch.stack([torch.cos(angles_rad), torch.sin(angles_rad)], -1)                    # [K, 2]
    delays = (look @ mic.t()) / speed                                                          # [K, M]
    steer = torch.exp(-2j * math.pi * delays[:, :, None] * freqs)                              # [K, M, F]
    power = torch.abs((steer * spec[None]).sum(1) / m) ** 2                                    # [K, F]
    power = power / (power.sum(0, keepdim=True) + 1e-8)
    return power.sum(-1)


class Criterion(nn.Module):
    """Same constructor and ``forward`` contract as ``utils/criterion.py:7-126``."""

    def __init__(self, cfg: dict, cfg_render: dict):
        super().__init__()
        self.spec_loss_weight = cfg["spec_loss_weight"]
        self.amplitude_loss_weight = cfg["amplitude_loss_weight"]
        self.angle_loss_weight = cfg["angle_loss_weight"]
        self.time_loss_weight = cfg["time_loss_weight"]
        self.energy_loss_weight = cfg["energy_loss_weight"]
        self.multi_stft_weight = cfg["multistft_loss_weight"]
        self.das_reg_loss_weight = cfg.get("das_reg_loss_weight", 0.0)
        self.das_ce_loss_weight = cfg.get("das_ce_loss_weight", 0.0)
        self.beta = cfg.get("beta", 100.0)
        self.angles_rad = torch.deg2rad(torch.arange(0.0, 360.0, 1.0))
        self.K = len(self.angles_rad)
        self.fs = cfg_render["fs"]
        self.sound_speed = cfg_render["speed"]
        self._dft, self._windows = {}, {}

    @property
    def weights(self) -> dict:
        return {"spec": float(self.spec_loss_weight), "amp": float(self.amplitude_loss_weight),
                "angle": float(self.angle_loss_weight), "time": float(self.time_loss_weight),
                "energy": float(self.energy_loss_weight), "mrstft": float(self.multi_stft_weight)}

    def dft_for(self, T: int, device) -> torch.Tensor:
        key = (int(T), str(device))
        if key not in self._dft:
            self._dft[key] = tables.dft_matrix(T).to(device)
        return self._dft[key]

    def window_for(self, win: int, device) -> torch.Tensor:
        key = (int(win), str(device))
        if key not in self._windows:
            self._windows[key] = torch.hann_window(win).to(device)
        return self._windows[key]

    def forward(self, pred_sig: torch.Tensor, ori_sig: torch.Tensor):
        if not pred_sig.is_complex() or not ori_sig.is_complex():
            raise TypeError("pred_sig / ori_sig are complex spectra [bs, T//2+1] (avr_runner.py:178-179)")
        if pred_sig.shape != ori_sig.shape or pred_sig.dim() != 2 or pred_sig.shape[1] < 3:
            raise ValueError("pred_sig and ori_sig must both be [bs, T//2+1]")
        losses, ori_time, pred_time = _CriterionFn.apply(torch.view_as_real(pred_sig.to(torch.complex64)),
                                                         torch.view_as_real(ori_sig.to(torch.complex64)), self)
        das_reg = torch.tensor(0.0, device=pred_sig.device)
        das_ce = torch.tensor(0.0, device=pred_sig.device)
        if self.das_reg_loss_weight > 0 or self.das_ce_loss_weight > 0:                    # criterion.py:105-124
            ang = self.angles_rad.to(pred_sig.device)
            p_pred = beamforming_power(pred_sig, self.fs, self.sound_speed, ang)
            p_ori = beamforming_power(ori_sig, self.fs, self.sound_speed, ang)
            if self.das_ce_loss_weight > 0:
                das_ce = F.cross_entropy(p_pred[None], torch.argmax(p_ori)[None]) * self.das_ce_loss_weight
            if self.das_reg_loss_weight > 0:
                a_pred = (torch.softmax(self.beta * p_pred, 0) * ang).sum()
                a_ori = (torch.softmax(self.beta * p_ori, 0) * ang).sum()
                das_reg = ((torch.sin(a_pred) - torch.sin(a_ori)).abs() +
                           (torch.cos(a_pred) - torch.cos(a_ori)).abs()) * self.das_reg_loss_weight
        return (losses[0], losses[1], losses[2], losses[3], losses[4], losses[5], das_reg, das_ce, ori_time, pred_time)
