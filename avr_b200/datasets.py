"""On-disk dataset formats of the reference (``/root/reference/datasets_loader.py:10-220``).

``WaveLoader(base_folder, dataset_type, eval, seq_len, fs)`` reads what the reference's loader reads and yields what
it yields -- ``(spec[F] complex64, position_rx[3], position_tx[3][, rotation_tx[3]], ch_idx)`` -- so that recorded
scenes can be rendered / trained against with ``avr_b200.AVRRender`` without the reference's ``librosa`` dependency:

=============  ==========================================================================  =========================
dataset_type   files under ``base_folder``                                                 reference lines
=============  ==========================================================================  =========================
``MeshRIR``    ``pos_mic.npy [M,3]``, ``pos_src.npy [1,3]``, ``train|test/ir_<i>.npy``     datasets_loader.py:61-93
               (48 kHz IRs ``[1,L]``: decimated by ``48000 // fs``, window of ``seq_len``
               samples starting at ``9100 // decimation``)
``Simu``       ``*.npz`` with ``ir``, ``position_rx``, ``position_tx``; sorted, first 90 %  :95-119
               train / last 10 % test
``Real_env``   ``train_test_split.pkl`` ({"train": [...], "test": [...]}) naming ``.npz``   :121-152
               files as above plus an optional ``ch_idx``
``RAF``        ``train|test/<sample>/rir.wav`` (48 kHz), ``rx_pos.txt``, ``tx_pos.txt``     :154-205, :223-247
               (quaternion + position; y/z swapped); training items are jittered by
               N(0, 0.1^2) per coordinate on every access
=============  ==========================================================================  =========================

The spectrum is ``numpy.fft.rfft`` of the (float64-promoted) time window cast to complex64, exactly as upstream.
This module is host-side I/O: no CUDA here, and nothing on the render hot path imports it.
"""
from __future__ import annotations

import glob
import math
import os
import pickle

import numpy as np
import torch
from torch.utils.data import Dataset

DATASET_TYPES = ("MeshRIR", "RAF", "Simu", "Real_env")


def read_wav_mono(path: str) -> np.ndarray:
    """Samples of a PCM / float WAV file as float32 in [-1, 1), channels averaged -- what
    ``librosa.load(path, sr=None, mono=True)`` returns (datasets_loader.py:168), without librosa."""
    from scipy.io import wavfile
    _, x = wavfile.read(path)
    if x.dtype == np.int16:
        y = x.astype(np.float32) / 32768.0
    elif x.dtype == np.int32:
        y = (x.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif x.dtype == np.uint8:
        y = (x.astype(np.float32) - 128.0) / 128.0
    else:
        y = x.astype(np.float32)
    return y.mean(axis=1) if y.ndim == 2 else y


def quaternion_to_direction_vector(q) -> np.ndarray:
    """Horizontal facing direction of an [x, y, z, w] quaternion (datasets_loader.py:223-247): the forward vector's
    x / z components, normalised in the horizontal plane, negated; third component 0."""
    x, y, z, w = (float(v) for v in q)
    fwd_x = 2.0 * (x * z + w * y)
    fwd_z = 1.0 - 2.0 * (x * x + y * y)
    norm = math.sqrt(fwd_x * fwd_x + fwd_z * fwd_z)
    return np.array([-fwd_x / norm, -fwd_z / norm, 0.0])


def _numbers(path: str) -> np.ndarray:
    vals = []
    with open(path, "r") as fh:
        for line in fh:
            vals += [float(tok) for tok in line.split(",") if tok.strip()]
    return np.array(vals)


class WaveLoader(Dataset):
    """Drop-in for the reference ``WaveLoader`` (same constructor, attributes and items)."""

    def __init__(self, base_folder, dataset_type="MeshRIR", eval=False, seq_len=2048, fs=16000):
        if dataset_type not in DATASET_TYPES:
            raise ValueError("Unsupported dataset type")
        self.dataset_type, self.eval = dataset_type, eval
        self.wave_max, self.wave_min = float("-inf"), float("inf")
        self.position_max = np.full(3, -np.inf)
        self.position_min = np.full(3, np.inf)
        self.ch_idx_list = []
        spectra, rx, tx, rot = [], [], [], []
        reader = {"MeshRIR": self._records_meshrir, "RAF": self._records_raf, "Simu": self._records_simu,
                  "Real_env": self._records_real_env}[dataset_type]
        for ir, p_rx, p_tx, r_tx, ch in reader(base_folder, eval, seq_len, fs):
            self.wave_max = max(self.wave_max, ir.max())
            self.wave_min = min(self.wave_min, ir.min())
            self.position_max = np.maximum(self.position_max, p_rx)
            self.position_min = np.minimum(self.position_min, p_rx)
            spectra.append(np.fft.rfft(ir))
            rx.append(p_rx)
            tx.append(p_tx)
            if r_tx is not None:
                rot.append(r_tx)
            if ch is not None:
                self.ch_idx_list.append(ch)
        self.wave_chunks = torch.tensor(np.array(spectra), dtype=torch.complex64)
        self.positions_rx = torch.tensor(np.array(rx), dtype=torch.float32)
        self.positions_tx = torch.tensor(np.array(tx), dtype=torch.float32)
        self.rotations_tx = torch.tensor(np.array(rot), dtype=torch.float32) if rot else []

    # -- one generator per format: (time window, rx, tx, tx facing or None, channel or None) ------------------
    def _records_meshrir(self, base, eval, seq_len, fs):
        dec = 48000 // fs
        self.default_st_idx = start = int(9100 / dec)
        folder = os.path.join(base, "test" if eval else "train")
        mic, src = np.load(os.path.join(base, "pos_mic.npy")), np.load(os.path.join(base, "pos_src.npy"))[0]
        for name in sorted(f for f in os.listdir(folder) if f.endswith(".npy")):
            ir = np.load(os.path.join(folder, name))[0, ::dec][start:start + seq_len]
            yield ir, mic[int(name.split("_")[1].split(".")[0])], src, None, None

    def _records_simu(self, base, eval, seq_len, fs):
        names = sorted(f for f in os.listdir(base) if f.endswith(".npz"))
        cut = int(0.9 * len(names))
        for name in (names[cut:] if eval else names[:cut]):
            rec = np.load(os.path.join(base, name))
            yield rec["ir"][:seq_len], rec["position_rx"], rec["position_tx"], None, None

    def _records_real_env(self, base, eval, seq_len, fs):
        with open(os.path.join(base, "train_test_split.pkl"), "rb") as fh:
            split = pickle.load(fh)
        for path in split["test" if eval else "train"]:
            rec = np.load(path if os.path.isabs(path) else os.path.join(base, path))
            ch = rec["ch_idx"].item() if "ch_idx" in rec else None
            yield rec["ir"][:seq_len], rec["position_rx"], rec["position_tx"], None, ch

    def _records_raf(self, base, eval, seq_len, fs):
        dec = int(48000 / fs)
        for folder in sorted(glob.glob(os.path.join(base, "test" if eval else "train", "*"))):
            ir = read_wav_mono(os.path.join(folder, "rir.wav"))[:seq_len * dec:dec]
            p_rx = _numbers(os.path.join(folder, "rx_pos.txt"))[[0, 2, 1]]             # y / z swapped
            info = _numbers(os.path.join(folder, "tx_pos.txt"))
            yield ir, p_rx, info[4:][[0, 2, 1]], quaternion_to_direction_vector(info[:4]), None

    # -- Dataset protocol ---------------------------------------------------------------------------------------
    def __len__(self):
        return len(self.wave_chunks)

    def __getitem__(self, idx):
        rx, tx = self.positions_rx[idx], self.positions_tx[idx]
        ch = self.ch_idx_list[idx] if self.ch_idx_list else -1
        if self.dataset_type != "RAF":
            return self.wave_chunks[idx], rx, tx, ch
        if not self.eval:                                               # sigma = 0.1 position jitter, rx first
            rx = rx + torch.randn_like(rx) * 0.1
            tx = tx + torch.randn_like(tx) * 0.1
        return self.wave_chunks[idx], rx, tx, self.rotations_tx[idx], ch
