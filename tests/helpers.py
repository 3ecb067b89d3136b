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


def oracle_fp32_noise(ref_net, render_kwargs, rx, tx, G, dtx=None, orders=(None, 1, 2), **fwd_kwargs):
    """How well the fp32 oracle determines the answer on this weight draw: ``(rel_l2 of the IR, {param name: rel_l2 of its
    gradient})``, the LARGEST distance of several fp32 evaluations of the oracle from its evaluation with float64 dense
    layers.  The fp32 evaluations differ only in the summation order of the dense layers (``field_ref.K_ORDER_SEED``:
    ``None`` = the plain oracle, then permuted reduction orders).  Geometry, renderer and the hash-grid cell positions
    stay fp32 in all of them (the network inputs are NOT widened: at the fine levels of the real grids, resolution ~2^21,
    a float64 position lands elsewhere in the cell and would measure a different function, not rounding noise).

    A single ReLU / |leaky_relu| decision that differs between two evaluations gates that unit's back-propagated gradient
    on or off; when the unit sits at a sample point that carries a visible share of the gradient (compositing weights and
    path loss concentrate it on few points) one such flip moves a hash-table gradient by 1e-4 .. 3e-3 (measured,
    profiles/r2/parity_flips.md).  The plain oracle alone under-reports this: it may happen to decide like float64 while
    any other rounding -- another GEMM blocking, the GPU's -- does not.  Hence several summation orders."""
    import copy

    net64 = copy.deepcopy(ref_net).double()
    for p in net64.parameters():
        p.grad = None
    out64 = render_ref.RenderRef(net64, **render_kwargs)(rx, tx, dtx, **fwd_kwargs)
    (out64 * G.double()).sum().backward()
    g64 = {n: p.grad for n, p in net64.named_parameters()}
    worst_out, worst = 0.0, {n: 0.0 for n in g64}
    saved = field_ref.K_ORDER_SEED
    try:
        for order in orders:
            field_ref.K_ORDER_SEED = order
            net32 = copy.deepcopy(ref_net)
            for p in net32.parameters():
                p.grad = None
            out32 = render_ref.RenderRef(net32, **render_kwargs)(rx, tx, dtx, **fwd_kwargs)
            (out32 * G).sum().backward()
            worst_out = max(worst_out, rel_l2(out32, out64))
            for n, p in net32.named_parameters():
                worst[n] = max(worst[n], rel_l2(p.grad, g64[n]))
    finally:
        field_ref.K_ORDER_SEED = saved
    return worst_out, worst


def oracle_gate_spread(ref_net, render_kwargs, rx, tx, G, dtx=None, tau=1e-6, **fwd_kwargs):
    """{param name: rel_l2} between the oracle's gradients with every ReLU / |leaky_relu| backward decision taken at
    ``v > +tau * mean|v|`` and at ``v > -tau * mean|v|`` (``field_ref.GATE_SHIFT``): ALL decisions that lie within fp32
    rounding noise of zero (an fp32 GEMM row is good to ~1e-6 of its scale) flipped at once.  Deterministic, unlike the
    summation-order probe, which only samples which of those decisions a particular rounding happens to flip."""
    import copy

    grads = []
    saved = field_ref.GATE_SHIFT
    try:
        for shift in (tau, -tau):
            field_ref.GATE_SHIFT = shift
            net = copy.deepcopy(ref_net)
            for p in net.parameters():
                p.grad = None
            out = render_ref.RenderRef(net, **render_kwargs)(rx, tx, dtx, **fwd_kwargs)
            (out * G).sum().backward()
            grads.append({n: p.grad for n, p in net.named_parameters()})
    finally:
        field_ref.GATE_SHIFT = saved
    return {n: rel_l2(grads[0][n], grads[1][n]) for n in grads[0]}


def oracle_conditioning(ref_net, render_kwargs, rx, tx, G, dtx=None, orders=(None, 1, 2), tau=1e-6, **fwd_kwargs):
    """-> (spread of the IR, {param: spread}): the larger of the summation-order probe and the gate-shift probe.  The
    parity bar of a test is ``max(1e-4, 2 * spread)``: where fp32 arithmetic itself leaves the answer open by more than
    the bar, no implementation -- the reference's included -- can be held to it."""
    n_out, a = oracle_fp32_noise(ref_net, render_kwargs, rx, tx, G, dtx=dtx, orders=orders, **fwd_kwargs)
    b = oracle_gate_spread(ref_net, render_kwargs, rx, tx, G, dtx=dtx, tau=tau, **fwd_kwargs)
    return n_out, {n: max(a[n], b[n]) for n in a}
