"""Shared helpers of the parity tests (test infrastructure; may import oracle/)."""
import os

import numpy as np
import torch

from avr_b200.configs import tiny_config
from oracle import field_ref, render_ref

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["avrmodel_sym", "avrmodel_box0_10", "avrmodel_complex", "stub_renderer_only"]

# mirrors oracle/make_golden.py::CASES (config kwargs only)
CASE_CFG = {
    "avrmodel_sym": ("AVRModel", dict()),
    "avrmodel_box0_10": ("AVRModel", dict(xyz_min=0, xyz_max=10, fs=4000)),
    "avrmodel_complex": ("AVRModel_complex", dict(xyz_min=-12, xyz_max=12, speed=346.8, pathloss=0.5, fs=8000)),
    "stub_renderer_only": ("stub", dict(n_azi=5, n_ele=4, n_samples=7, T=240)),
}


def load_golden(name):
    blob = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: torch.from_numpy(blob[k]) for k in blob.files}


def case_config(name):
    model_class, kw = CASE_CFG[name]
    return model_class, tiny_config("AVRModel" if model_class == "stub" else model_class, **kw)


def oracle_field(model_class, model_cfg, golden=None):
    cls = field_ref.AVRModelRef if model_class == "AVRModel" else field_ref.AVRModelComplexRef
    net = cls(model_cfg)
    if golden is not None:
        sd = {k[len("param/"):]: v for k, v in golden.items() if k.startswith("param/")}
        net.load_state_dict(sd)
    return net


def rel_l2(a, b):
    return render_ref.rel_l2(torch.as_tensor(a).cpu(), torch.as_tensor(b).cpu())
