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


def oracle_fp32_noise(ref_net, render_kwargs, rx, tx, G, dtx=None, **fwd_kwargs):
    """How far the fp32 oracle is from itself with the dense layers evaluated in float64: ``(rel_l2 of the IR,
    {param name: rel_l2 of its gradient})``.  Geometry, renderer and the hash-grid cell positions stay fp32 (the
    network inputs are NOT widened: at the fine levels of the real grids, resolution ~2^21, a float64 position lands
    elsewhere in the cell and would measure a different function, not rounding noise).  A single ReLU /
    |leaky_relu| decision that differs between the two evaluations moves a gradient by ~1e-4, so on some weight
    draws this noise of the checker reaches the 1e-4 parity bar; tests take the first seed where it does not."""
    import copy

    net64 = copy.deepcopy(ref_net).double()
    net32 = copy.deepcopy(ref_net)
    for p in list(net64.parameters()) + list(net32.parameters()):
        p.grad = None
    out32 = render_ref.RenderRef(net32, **render_kwargs)(rx, tx, dtx, **fwd_kwargs)
    (out32 * G).sum().backward()
    out64 = render_ref.RenderRef(net64, **render_kwargs)(rx, tx, dtx, **fwd_kwargs)
    (out64 * G.double()).sum().backward()
    g32 = dict(net32.named_parameters())
    return rel_l2(out32, out64), {n: rel_l2(g32[n].grad, p.grad) for n, p in net64.named_parameters()}
