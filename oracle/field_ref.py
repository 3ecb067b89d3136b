"""fp32 torch restatement of the tiny-cuda-nn field used by the reference ``model.py``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED at the tiny-cuda-nn
boundary: tcnn is not vendored/pinned/installed (``/root/reference/requirements.txt:11``);
what follows is the published HashGrid / bias-free-MLP algorithm as recorded in
SURVEY.md Appendix B, driven exactly the way ``/root/reference/model.py`` drives tcnn:

* ``AVRModelRef``         <- ``model.py:63-235``  (``AVRModel``; MeshRIR / Simu / Real_env)
* ``AVRModelComplexRef``  <- ``model.py:238-331`` (``AVRModel_complex``; RAF)

Parameters live in one flat fp32 ``params`` tensor per encoding / network, named like the
reference's sub-modules (``_pos_encoding.params`` ...), so a state-dict is interchangeable
with ``avr_b200.model``.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_U32 = 0xFFFFFFFF
_PRIME_Y = 2654435761
_PRIME_Z = 805459861


def _pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def hashgrid_geometry(cfg: dict):
    """Per-level (scale, resolution, entries, offset) -- SURVEY App. B.1."""
    n_levels = int(cfg.get("n_levels", 16))
    n_feat = int(cfg.get("n_features_per_level", 2))
    log2_size = int(cfg.get("log2_hashmap_size", 19))
    base = int(cfg.get("base_resolution", 16))
    pls = float(cfg.get("per_level_scale", 2.0))
    log2_pls = np.float32(math.log2(pls))
    scales, ress, sizes, offsets = [], [], [], []
    off = 0
    for lvl in range(n_levels):
        # exp2f(level * log2(pls)) * base - 1   evaluated in fp32
        scale = np.float32(np.exp2(np.float32(lvl) * log2_pls, dtype=np.float32) * np.float32(base) - np.float32(1.0))
        res = int(math.ceil(float(scale))) + 1
        dense = res ** 3
        entries = min(_pad_to(min(dense, 2 ** 40), 8), 1 << log2_size)
        scales.append(float(scale))
        ress.append(res)
        sizes.append(entries)
        offsets.append(off)
        off += entries
    return {
        "n_levels": n_levels, "n_feat": n_feat, "scale": scales, "res": ress,
        "size": sizes, "offset": offsets, "total": off,
    }


def _mul_u32(a: torch.Tensor, k: int) -> torch.Tensor:
    """(a * k) mod 2^32 for int64 tensors holding uint32 values, without int64 overflow."""
    lo = a & 0xFFFF
    hi = a >> 16
    return (lo * k + (((hi * k) & 0xFFFF) << 16)) & _U32


class HashGridRef(nn.Module):
    """Multiresolution hash grid, 3-D input, linear interpolation."""

    def __init__(self, cfg: dict, seed: int = 1337):
        super().__init__()
        self.geom = hashgrid_geometry(cfg)
        g = torch.Generator().manual_seed(seed)
        n = self.geom["total"] * self.geom["n_feat"]
        self.params = nn.Parameter((torch.rand(n, generator=g) * 2 - 1) * 1e-4)
        self.n_output_dims = self.geom["n_levels"] * self.geom["n_feat"]

    def corner_indices(self, u: torch.Tensor, lvl: int):
        """-> (idx[N,8] int64 into this level's table, w[N,8] fp32)."""
        geo = self.geom
        scale, res, size = geo["scale"][lvl], geo["res"][lvl], geo["size"][lvl]
        pos = (u.double() * scale + 0.5).float()            # == fmaf(scale, u, 0.5f)
        fl = torch.floor(pos)
        frac = pos - fl
        grid = fl.to(torch.int64) & _U32                    # (uint32)(int)floorf
        idxs, ws = [], []
        for corner in range(8):
            w = torch.ones_like(frac[:, 0])
            cs = []
            for d in range(3):
                bit = (corner >> d) & 1
                w = w * (frac[:, d] if bit else (1 - frac[:, d]))
                cs.append((grid[:, d] + bit) & _U32)
            stride, index = 1, torch.zeros_like(cs[0])
            d = 0
            while d < 3 and stride <= size:
                index = (index + _mul_u32(cs[d], stride & _U32)) & _U32
                stride *= res
                d += 1
            if size < stride:
                index = cs[0] ^ _mul_u32(cs[1], _PRIME_Y) ^ _mul_u32(cs[2], _PRIME_Z)
            idxs.append(index % size)
            ws.append(w)
        return torch.stack(idxs, 1), torch.stack(ws, 1)

    def forward(self, u: torch.Tensor) -> torch.Tensor:
        geo = self.geom
        nf = geo["n_feat"]
        table = self.params.view(-1, nf)
        outs = []
        for lvl in range(geo["n_levels"]):
            idx, w = self.corner_indices(u, lvl)
            vals = table[geo["offset"][lvl] + idx]          # [N,8,nf]
            outs.append((vals * w.unsqueeze(-1)).sum(1))
        return torch.cat(outs, dim=-1)


class MLPRef(nn.Module):
    """Bias-free ReLU MLP with tcnn padding rules.

    ``n_hidden_layers = h`` -> ``h + 1`` row-major ``[out, in]`` matrices; ReLU after all but the
    last.  Input is padded with ones to a multiple of 16 (FullyFusedMLP) / 8 (CutlassMLP);
    output padded likewise and sliced back.
    """

    def __init__(self, n_in: int, n_out: int, cfg: dict, seed: int = 1337):
        super().__init__()
        align = 16 if cfg.get("otype", "FullyFusedMLP") == "FullyFusedMLP" else 8
        width = int(cfg["n_neurons"])
        hidden = int(cfg["n_hidden_layers"])
        self.n_in, self.n_out = n_in, n_out
        self.in_pad, self.out_pad = _pad_to(n_in, align), _pad_to(n_out, align)
        dims = [self.in_pad] + [width] * hidden + [self.out_pad]
        self.shapes = [(dims[i + 1], dims[i]) for i in range(len(dims) - 1)]
        g = torch.Generator().manual_seed(seed)
        chunks = []
        for (o, i) in self.shapes:
            bound = math.sqrt(6.0 / (i + o))
            chunks.append((torch.rand(o * i, generator=g) * 2 - 1) * bound)
        self.params = nn.Parameter(torch.cat(chunks))
        self.n_output_dims = n_out

    def matrices(self):
        out, off = [], 0
        for (o, i) in self.shapes:
            out.append(self.params[off:off + o * i].view(o, i))
            off += o * i
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.in_pad != self.n_in:
            x = torch.cat([x, x.new_ones(x.shape[0], self.in_pad - self.n_in)], dim=-1)
        mats = self.matrices()
        for k, w in enumerate(mats):
            x = x @ w.t()
            if k + 1 < len(mats):
                x = F.relu(x)
        return x[:, : self.n_out]


def _channel_embed_mode(cfg: dict):
    ch = cfg.get("channel_embed") or {}
    if ch.get("is_embed", False) and ch.get("connection_type", None) in ("add", "concat"):
        raise NotImplementedError("channel_embed add/concat is outside the five BASELINE configs")


class AVRModelRef(nn.Module):
    """``/root/reference/model.py:63-235``."""

    def __init__(self, cfg: dict, seed: int = 1337):
        super().__init__()
        _channel_embed_mode(cfg)
        self._pos_encoding = HashGridRef(cfg["pos_encoding_sigma"], seed)
        self._dir_encoding = HashGridRef(cfg["dir_encoding_sig"], seed + 1)
        self._tx_encoding = HashGridRef(cfg["tx_encoding_sig"], seed + 2)
        self.signal_output_dim = int(cfg["signal_output_dim"])
        self._model_encoder_sigma = MLPRef(self._pos_encoding.n_output_dims, 128, cfg["sigma_encoder_network"], seed + 3)
        self._model_decoder_sigma = MLPRef(128, 1, cfg["sigma_decoder_network"], seed + 4)
        sig_in = 128 + self._dir_encoding.n_output_dims + self._tx_encoding.n_output_dims
        self._model_signal = MLPRef(sig_in, self.signal_output_dim, cfg["signal_network"], seed + 5)
        self.leaky_slope = 0.01          # model.py:233 uses F.leaky_relu's default, not cfg.leaky_relu

    def forward(self, pts, view, tx, ch_idx=None):
        bs, n_pts = pts.size(0), pts.size(1)
        pts = (pts.reshape(-1, 3) + 1) / 2               # model.py:187-189
        view = (view.reshape(-1, 3) + 1) / 2
        tx = (tx.reshape(-1, 3) + 1) / 2
        sigma_feat = self._model_encoder_sigma(self._pos_encoding(pts))            # :191,206
        attn = self._model_decoder_sigma(F.relu(sigma_feat))                        # :209-216
        sig_in = torch.cat([sigma_feat, self._dir_encoding(view), self._tx_encoding(tx)], dim=-1)   # :219-221
        signal = self._model_signal(sig_in)                                         # :231
        attn = torch.abs(F.leaky_relu(attn, self.leaky_slope)).view(bs, n_pts, 1)   # :233
        return attn, signal.view(bs, n_pts, self.signal_output_dim)


class AVRModelComplexRef(nn.Module):
    """``/root/reference/model.py:238-331`` (RAF).  Accepts and ignores ``ch_idx``."""

    def __init__(self, cfg: dict, seed: int = 1337):
        super().__init__()
        self.leaky_slope = float(cfg["leaky_relu"])
        self.signal_output_dim = int(cfg["signal_output_dim"])
        self._pos_encoding = HashGridRef(cfg["pos_encoding_sigma"], seed)
        self._pos_signal_encoding = HashGridRef(cfg["pos_encoding_sig"], seed + 1)
        self._tx_pos_encoding = HashGridRef(cfg["tx_pos_encoding_sigma"], seed + 2)
        self._tx_pos_signal_encoding = HashGridRef(cfg["tx_pos_encoding_sig"], seed + 3)
        self._dir_encoding = HashGridRef(cfg["dir_encoding_sig"], seed + 4)
        self._tx_dir_encoding = HashGridRef(cfg["tx_dir_encoding_sig"], seed + 5)
        n_enc = self._pos_encoding.n_output_dims
        self._model_encoder_sigma = MLPRef(2 * n_enc, 256, cfg["sigma_encoder_network"], seed + 6)
        self._model_decoder_sigma = MLPRef(256, 1, cfg["sigma_decoder_network"], seed + 7)
        self._model_signal = MLPRef(256 + 4 * n_enc, self.signal_output_dim, cfg["signal_network"], seed + 8)

    def forward(self, pts, view, tx, tx_view, ch_idx=None):
        bs, n_pts = pts.size(0), pts.size(1)
        pts = (pts.reshape(-1, 3) + 1) / 2               # model.py:308-311
        view = (view.reshape(-1, 3) + 1) / 2
        tx = (tx.reshape(-1, 3) + 1) / 2
        tx_view = (tx_view.reshape(-1, 3) + 1) / 2
        sigma_feat = self._model_encoder_sigma(
            torch.cat([self._pos_encoding(pts), self._tx_pos_encoding(tx)], -1))    # :313-318
        attn = self._model_decoder_sigma(F.relu(sigma_feat))                        # :319
        feat = torch.cat([F.relu(sigma_feat), self._dir_encoding(view), self._tx_dir_encoding(tx_view),
                          self._pos_signal_encoding(pts), self._tx_pos_signal_encoding(tx)], -1)     # :321-326
        signal = self._model_signal(feat)                                           # :327
        attn = torch.abs(F.leaky_relu(attn, self.leaky_slope)).view(bs, n_pts, 1)   # :329
        return attn, signal.reshape(bs, n_pts, self.signal_output_dim)


def trained_like_(model: nn.Module, seed: int = 7, table_std: float = 0.1) -> nn.Module:
    """Overwrite hash tables with N(0, table_std) so interior alphas are O(0.01-1)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, HashGridRef):
                mod.params.copy_(torch.randn(mod.params.shape, generator=g) * table_std)
    return model
