"""``avr_b200.WaveLoader`` on small fixtures written in the reference's four on-disk formats
(/root/reference/datasets_loader.py:61-220): known answers everywhere, and item-by-item equality with the UNMODIFIED
reference loader where the reference tree exists (``librosa`` stood in by a scipy WAV reader)."""
import os
import pickle

import numpy as np
import pytest
import torch
from scipy.io import wavfile

import avr_b200
from avr_b200.datasets import quaternion_to_direction_vector, read_wav_mono
from oracle.reference_shim import load_reference_datasets, reference_available


def _write_meshrir(root, n=6, length=12000):
    rng = np.random.default_rng(0)
    os.makedirs(root / "train"); os.makedirs(root / "test")
    np.save(root / "pos_mic.npy", rng.uniform(-0.5, 0.5, (n, 3)))
    np.save(root / "pos_src.npy", np.array([[2.0, 0.0, 0.1]]))
    irs = {}
    for i in range(n):
        ir = rng.standard_normal((1, length))
        irs[i] = ir
        np.save(root / ("test" if i % 3 == 0 else "train") / f"ir_{i}.npy", ir)
    return irs


def _write_simu(root, n=10, length=700, with_ch=False):
    rng = np.random.default_rng(1)
    os.makedirs(root, exist_ok=True)
    names = []
    for i in range(n):
        rec = dict(ir=rng.standard_normal(length).astype(np.float32), position_rx=rng.uniform(-4, 4, 3), position_tx=rng.uniform(-4, 4, 3))
        if with_ch:
            rec["ch_idx"] = np.array(i % 8)
        np.savez(root / f"sample_{i:03d}.npz", **rec)
        names.append(f"sample_{i:03d}.npz")
    return names


def _write_raf(root, n=3, length=4000):
    rng = np.random.default_rng(2)
    for split in ("train", "test"):
        for i in range(n):
            d = root / split / f"{i:04d}"
            os.makedirs(d)
            wavfile.write(d / "rir.wav", 48000, (rng.standard_normal(length) * 3000).astype(np.int16))
            (d / "rx_pos.txt").write_text(",".join(f"{v:.6f}" for v in rng.uniform(-3, 3, 3)))
            q = rng.standard_normal(4); q /= np.linalg.norm(q)
            (d / "tx_pos.txt").write_text(",".join(f"{v:.6f}" for v in q) + "\n" + ",".join(f"{v:.6f}" for v in rng.uniform(-3, 3, 3)))


def test_meshrir_format_known_answers(tmp_path):
    irs = _write_meshrir(tmp_path)
    ds = avr_b200.WaveLoader(str(tmp_path), "MeshRIR", eval=False, seq_len=512, fs=24000)
    assert len(ds) == 4 and ds.default_st_idx == 4550
    spec, rx, tx, ch = ds[0]                                           # ir_1.npy: sorted order, train split
    want = np.fft.rfft(irs[1][0, ::2][4550:4550 + 512])
    assert spec.dtype == torch.complex64 and spec.shape == (257,) and ch == -1
    assert torch.equal(spec, torch.tensor(want, dtype=torch.complex64))
    assert torch.equal(rx, torch.tensor(np.load(tmp_path / "pos_mic.npy")[1], dtype=torch.float32))
    assert torch.allclose(tx, torch.tensor([2.0, 0.0, 0.1]))
    assert len(avr_b200.WaveLoader(str(tmp_path), "MeshRIR", eval=True, seq_len=512, fs=24000)) == 2


def test_simu_and_real_env_formats(tmp_path):
    names = _write_simu(tmp_path / "simu")
    tr = avr_b200.WaveLoader(str(tmp_path / "simu"), "Simu", eval=False, seq_len=600)
    te = avr_b200.WaveLoader(str(tmp_path / "simu"), "Simu", eval=True, seq_len=600)
    assert len(tr) == 9 and len(te) == 1 and tr[0][0].shape == (301,) and tr[3][3] == -1
    names = _write_simu(tmp_path / "real", with_ch=True)
    with open(tmp_path / "real" / "train_test_split.pkl", "wb") as fh:
        pickle.dump({"train": names[:7], "test": [str(tmp_path / "real" / n) for n in names[7:]]}, fh)   # relative and absolute
    tr = avr_b200.WaveLoader(str(tmp_path / "real"), "Real_env", eval=False, seq_len=512)
    te = avr_b200.WaveLoader(str(tmp_path / "real"), "Real_env", eval=True, seq_len=512)
    assert len(tr) == 7 and len(te) == 3 and [tr[i][3] for i in range(7)] == [0, 1, 2, 3, 4, 5, 6] and te[0][3] == 7
    rec = np.load(tmp_path / "real" / names[2])
    assert torch.equal(tr[2][0], torch.tensor(np.fft.rfft(rec["ir"][:512]), dtype=torch.complex64))
    with pytest.raises(ValueError):
        avr_b200.WaveLoader(str(tmp_path), "Other")


def test_raf_format_jitter_and_quaternion(tmp_path):
    _write_raf(tmp_path)
    te = avr_b200.WaveLoader(str(tmp_path), "RAF", eval=True, seq_len=1000, fs=16000)
    spec, rx, tx, rot, ch = te[1]
    d = tmp_path / "test" / "0001"
    wav = read_wav_mono(str(d / "rir.wav"))
    assert wav.dtype == np.float32 and abs(wav).max() < 1
    assert torch.equal(spec, torch.tensor(np.fft.rfft(wav[:3000:3]), dtype=torch.complex64)) and spec.shape == (501,)
    p = np.array([float(v) for v in (d / "rx_pos.txt").read_text().split(",")])
    assert torch.allclose(rx, torch.tensor(p[[0, 2, 1]], dtype=torch.float32)) and ch == -1
    assert rot.shape == (3,) and float(rot[2]) == 0 and abs(float(rot.norm()) - 1) < 1e-6
    assert np.allclose(quaternion_to_direction_vector([0, 0, 0, 1]), [0, -1, 0])       # identity rotation faces -y
    tr = avr_b200.WaveLoader(str(tmp_path), "RAF", eval=False, seq_len=1000, fs=16000)
    torch.manual_seed(0)
    a = tr[0]
    torch.manual_seed(0)
    want_rx = tr.positions_rx[0] + torch.randn(3) * 0.1                  # rx is jittered first, then tx
    want_tx = tr.positions_tx[0] + torch.randn(3) * 0.1
    assert torch.equal(a[1], want_rx) and torch.equal(a[2], want_tx)
    assert not torch.equal(tr[0][1], tr[0][1])                           # fresh jitter on every access
    assert torch.equal(te[0][1], te[0][1])                               # none in eval


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
def test_items_equal_the_unmodified_reference_loader(tmp_path):
    ref = load_reference_datasets()
    _write_meshrir(tmp_path / "mesh")
    _write_simu(tmp_path / "simu")
    names = _write_simu(tmp_path / "real", with_ch=True)
    with open(tmp_path / "real" / "train_test_split.pkl", "wb") as fh:
        pickle.dump({"train": names[:7], "test": names[7:]}, fh)
    _write_raf(tmp_path / "raf")
    cases = [("mesh", "MeshRIR", 512, 24000), ("simu", "Simu", 600, 16000), ("real", "Real_env", 512, 16000), ("raf", "RAF", 1000, 16000)]
    for sub, kind, seq_len, fs in cases:
        for ev in (False, True):
            a = ref.WaveLoader(str(tmp_path / sub), dataset_type=kind, eval=ev, seq_len=seq_len, fs=fs)
            b = avr_b200.WaveLoader(str(tmp_path / sub), dataset_type=kind, eval=ev, seq_len=seq_len, fs=fs)
            assert len(a) == len(b) > 0
            assert a.wave_max == b.wave_max and a.wave_min == b.wave_min
            assert np.array_equal(a.position_max, b.position_max) and np.array_equal(a.position_min, b.position_min)
            for i in range(len(a)):
                torch.manual_seed(i)
                x = a[i]
                torch.manual_seed(i)
                y = b[i]
                assert len(x) == len(y)
                for u, v in zip(x, y):
                    assert (torch.equal(u, v) if torch.is_tensor(u) else u == v), (kind, ev, i)
