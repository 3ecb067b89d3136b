"""Import the UNMODIFIED reference CPU renderer (this container only).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``/root/reference`` does not exist on the
GPU box, so nothing that runs there may call ``load_reference_renderer()``; tests that use it
skip when the tree is absent.  No reference source is copied: the module is executed from
where it lies.
"""
from __future__ import annotations

import importlib.util
import os

REFERENCE_ROOT = os.environ.get("AVR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "renderer_cpu.py"))


def load_reference_renderer():
    """-> the ``renderer_cpu`` module object (``AVRRender``, ``ray_directions`` ...)."""
    if not reference_available():
        raise FileNotFoundError(f"no reference tree at {REFERENCE_ROOT}")
    spec = importlib.util.spec_from_file_location(
        "avr_reference_renderer_cpu", os.path.join(REFERENCE_ROOT, "renderer_cpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_model():
    """-> the reference ``model`` module (``AVRModel``, ``AVRModel_complex``, ``LayeredTCNNWithInjection``) executed
    from where it lies, with a stand-in ``tinycudann`` whose ``Encoding`` / ``Network`` are the oracle's
    ``HashGridRef`` / ``MLPRef`` (tiny-cuda-nn itself is absent: ``requirements.txt:11``).  What this pins
    is how ``model.py`` DRIVES tcnn -- concatenation order, which activations, the channel-embedding variants --
    not tcnn's arithmetic."""
    import sys
    import types

    from . import field_ref

    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py")):
        raise FileNotFoundError(f"no reference tree at {REFERENCE_ROOT}")
    counter = [0]

    def encoding(n_input_dims, encoding_config, dtype=None, seed=None):
        assert n_input_dims == 3
        counter[0] += 1
        return field_ref.HashGridRef(encoding_config, seed=1000 + counter[0])

    def network(n_input_dims, n_output_dims, network_config, seed=None):
        counter[0] += 1
        return field_ref.MLPRef(n_input_dims, n_output_dims, network_config, seed=1000 + counter[0])

    fake = types.ModuleType("tinycudann")
    fake.Encoding, fake.Network = encoding, network
    saved = sys.modules.get("tinycudann")
    sys.modules["tinycudann"] = fake
    try:
        spec = importlib.util.spec_from_file_location("avr_reference_model", os.path.join(REFERENCE_ROOT, "model.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            del sys.modules["tinycudann"]
        else:
            sys.modules["tinycudann"] = saved
    return mod


def load_reference_criterion():
    """-> the reference ``utils/criterion.py`` module executed from where it lies, with a stand-in ``auraloss`` whose
    ``freq.MultiResolutionSTFTLoss`` is the oracle's restatement (auraloss is absent and unpinned).  Pins everything
    ``Criterion.forward`` does itself; the multi-resolution STFT term stays parity-unpinned."""
    import sys
    import types

    from . import criterion_ref

    path = os.path.join(REFERENCE_ROOT, "utils", "criterion.py")
    if not os.path.isfile(path):
        raise FileNotFoundError(f"no reference tree at {REFERENCE_ROOT}")
    fake = types.ModuleType("auraloss")
    fake.freq = types.ModuleType("auraloss.freq")
    fake.freq.MultiResolutionSTFTLoss = criterion_ref.MultiResolutionSTFTLossRef
    saved = {k: sys.modules.get(k) for k in ("auraloss", "auraloss.freq")}
    sys.modules["auraloss"], sys.modules["auraloss.freq"] = fake, fake.freq
    try:
        spec = importlib.util.spec_from_file_location("avr_reference_criterion", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def load_reference_datasets():
    """-> the reference ``datasets_loader`` module executed from where it lies, with a stand-in ``librosa`` whose ``load``
    reads WAV files through scipy (librosa is absent here).  Pins the on-disk formats of ``avr_b200/datasets.py``."""
    import sys
    import types

    path = os.path.join(REFERENCE_ROOT, "datasets_loader.py")
    if not os.path.isfile(path):
        raise FileNotFoundError(f"no reference tree at {REFERENCE_ROOT}")

    def load(p, sr=None, mono=True):
        from scipy.io import wavfile
        import numpy as np
        rate, x = wavfile.read(p)
        assert sr is None and mono
        y = x.astype(np.float32) / 32768.0 if x.dtype == np.int16 else x.astype(np.float32)
        return (y.mean(axis=1) if y.ndim == 2 else y), rate

    fake = types.ModuleType("librosa")
    fake.load = load
    saved = sys.modules.get("librosa")
    sys.modules["librosa"] = fake
    try:
        spec = importlib.util.spec_from_file_location("avr_reference_datasets_loader", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop("librosa", None)
        else:
            sys.modules["librosa"] = saved
    return mod
