"""End-to-end parity of AVRRender on the GPU against the oracle and the committed golden vectors.

Bar (BASELINE.json north_star): rendered IR and parameter gradients within 1e-4 relative L2 (fp32);
sample positions / delay indices bit-exact (tests/test_gpu_kernels.py).
"""
import pytest
import torch

import avr_b200
from avr_b200.configs import get_config, tiny_config
from oracle import field_ref, render_ref
from tests.helpers import GOLDEN_CASES, case_config, load_golden, oracle_conditioning, oracle_field, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _native_from(ref_net, model_class, model_cfg):
    cls = avr_b200.AVRModel if model_class == "AVRModel" else avr_b200.AVRModel_complex
    net = cls(model_cfg)
    net.load_state_dict(ref_net.state_dict())
    return net.to(DEV)


def _check_grads(native, ref_net, tol=TOL):
    worst = 0.0
    ref_grads = dict(ref_net.named_parameters())
    for name, p in native.named_parameters():
        assert p.grad is not None, name
        err = rel_l2(p.grad, ref_grads[name].grad)
        worst = max(worst, err)
        assert err < tol, (name, err)
    return worst


@pytest.mark.parametrize("dense", ["tc", "tc-bf16x3", "simt"])
@pytest.mark.parametrize("name", GOLDEN_CASES[:3])
def test_fused_render_vs_golden(built_library, name, dense, monkeypatch):
    if dense == "tc-bf16x3":                                   # bf16 triples in the signal network's hidden layers too
        from avr_b200 import fused_tc
        monkeypatch.setattr(fused_tc, "SIG_HIDDEN_F16", False)
        dense = "tc"
    g = load_golden(name)
    model_class, cfg = case_config(name)
    ref_net = oracle_field(model_class, cfg["model"], g)
    native = _native_from(ref_net, model_class, cfg["model"])
    ren = avr_b200.AVRRender(native, **cfg["render"], dense=dense)
    dtx = g["dir_tx"].to(DEV) if "dir_tx" in g else None
    out = ren(g["rx"].to(DEV), g["tx"].to(DEV), dtx, azi_rand=g["azi_rand"])
    assert out.shape == g["out"].shape and out.dtype == torch.float32
    assert rel_l2(out, g["out"]) < TOL
    (out * g["G"].to(DEV)).sum().backward()
    for pname, p in native.named_parameters():
        assert rel_l2(p.grad, g["grad/" + pname]) < TOL, pname


def test_generic_network_path_vs_golden(built_library):
    g = load_golden("stub_renderer_only")
    _, cfg = case_config("stub_renderer_only")

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.attn = torch.nn.Parameter(g["param/attn"].clone())
            self.signal = torch.nn.Parameter(g["param/signal"].clone())

        def forward(self, pts, view, tx, ch_idx=None):
            assert pts.shape == (2, 22 * 7, 3) and view.shape == pts.shape and tx.shape == pts.shape
            return self.attn, self.signal

    net = Stub().to(DEV)
    ren = avr_b200.AVRRender(net, **cfg["render"])
    out = ren(g["rx"].to(DEV), g["tx"].to(DEV), azi_rand=g["azi_rand"])
    assert rel_l2(out, g["out"]) < 1e-5
    (out * g["G"].to(DEV)).sum().backward()
    assert rel_l2(net.attn.grad, g["grad/attn"]) < TOL
    assert rel_l2(net.signal.grad, g["grad/signal"]) < 1e-5


@pytest.mark.parametrize("dense", ["tc", "simt"])
@pytest.mark.parametrize("model_class,kw,bs", [
    ("AVRModel", dict(n_azi=12, n_ele=6, n_samples=24, T=400, width_sigma=64, width_signal=128), 3),
    ("AVRModel", dict(n_azi=9, n_ele=5, n_samples=40, T=320, xyz_min=0, xyz_max=10, fs=4000), 1),
    ("AVRModel_complex", dict(n_azi=10, n_ele=5, n_samples=16, T=480, fs=8000, xyz_min=-12, xyz_max=12), 2),
])
def test_fused_render_vs_oracle_seeded(built_library, model_class, kw, bs, dense):
    cfg = tiny_config(model_class, **kw)
    cls = field_ref.AVRModelRef if model_class == "AVRModel" else field_ref.AVRModelComplexRef
    ref_net = field_ref.trained_like_(cls(cfg["model"], seed=21), seed=22)
    native = _native_from(ref_net, model_class, cfg["model"])
    r = cfg["render"]
    gen = torch.Generator().manual_seed(5)
    c = (r["xyz_min"] + r["xyz_max"]) / 2
    rx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 2).float()
    tx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 2).float()
    dtx = torch.nn.functional.normalize(torch.randn(bs, 3, generator=gen), dim=-1) if model_class != "AVRModel" else None
    azi = torch.rand(r["n_azi"], generator=gen)
    G = torch.randn(bs, kw["T"] // 2 + 1, 2, generator=gen)
    ref_out = render_ref.RenderRef(ref_net, **r)(rx, tx, dtx, azi_rand=azi)
    (ref_out * G).sum().backward()
    ren = avr_b200.AVRRender(native, **r, dense=dense)
    out = ren(rx.to(DEV), tx.to(DEV), dtx.to(DEV) if dtx is not None else None, azi_rand=azi)
    assert float(ref_out.abs().max()) > 0
    assert rel_l2(out, ref_out) < TOL
    (out * G.to(DEV)).sum().backward()
    _check_grads(native, ref_net)


def _embed_cfg(conn, enc, dec, sig):
    return {"is_embed": True, "ch_num": 8, "connection_type": conn, "is_sigma_encoder": enc, "is_sigma_decoder": dec,
            "is_signal_network": sig, "emb_dim_sigma_encoder": 8, "emb_dim_sigma_decoder": 16, "emb_dim_signal_network": 24}


@pytest.mark.parametrize("conn,enc,dec,sig", [
    ("add", True, False, True),          # config_files/real_exp_ch_emb_add/*.yml
    ("add", True, True, True), ("add", False, True, False),
    ("concat", True, True, True), ("concat", False, False, True), ("concat", True, False, False), ("concat", False, True, False),
])
def test_channel_embedding_vs_oracle(built_library, conn, enc, dec, sig):
    """model.py:11-61 ('add': per-layer embedding row added before the ReLU) and :108-113,201-228 ('concat')."""
    cfg = tiny_config("AVRModel", n_azi=12, n_ele=6, n_samples=24, T=400, width_sigma=64, width_signal=128)
    cfg["model"]["channel_embed"] = _embed_cfg(conn, enc, dec, sig)
    r = cfg["render"]
    bs = 3
    gen = torch.Generator().manual_seed(5)
    rx = ((torch.rand(bs, 3, generator=gen) * 2 - 1) * 2).float()
    tx = ((torch.rand(bs, 3, generator=gen) * 2 - 1) * 2).float()
    ch = torch.tensor([5, 0, 5])                                   # a repeated channel: its row gradient is a sum
    azi = torch.rand(r["n_azi"], generator=gen)
    G = torch.randn(bs, 201, 2, generator=gen)
    # Random embedding rows can put a tiny problem on a kink (|leaky_relu(.)| at 0, a ReLU at 0): then fp32 evaluations that
    # differ only in summation order disagree with each other.  No seed is rejected: the bar is 1e-4, widened to twice the
    # oracle's own spread when that is larger (tests/helpers.py::oracle_conditioning).
    ref_net = field_ref.trained_like_(field_ref.AVRModelRef(cfg["model"], seed=21), seed=22)
    n_out, noise = oracle_conditioning(ref_net, r, rx, tx, G, ch_idx=ch, azi_rand=azi)
    tol = max(TOL, 2 * max([n_out] + list(noise.values())))
    print(f"\nchannel embedding {conn} enc={enc} dec={dec} sig={sig}: oracle spread {max(noise.values()):.1e} -> bar {tol:.1e}")
    native = _native_from(ref_net, "AVRModel", cfg["model"])
    assert [k for k, _ in native.named_parameters()] == [k for k, _ in ref_net.named_parameters()]
    ref_out = render_ref.RenderRef(ref_net, **r)(rx, tx, None, ch_idx=ch, azi_rand=azi)
    (ref_out * G).sum().backward()
    ren = avr_b200.AVRRender(native, **r)
    out = ren(rx.to(DEV), tx.to(DEV), ch_idx=ch.to(DEV), azi_rand=azi)
    assert float(ref_out.abs().max()) > 0 and rel_l2(out, ref_out) < tol
    (out * G.to(DEV)).sum().backward()
    _check_grads(native, ref_net, tol)
    emb = [p for n, p in native.named_parameters() if "embedding" in n]
    assert emb and all(float(p.grad[[1, 2, 3, 4, 6, 7]].abs().max()) == 0 and float(p.grad[5].abs().max()) > 0 for p in emb)
    # the standalone field (explicit points, torch-composed layers) agrees as well
    pts = torch.rand(2, 40, 3, generator=gen) * 2 - 1
    a0, s0 = ref_net(pts, pts.flip(1), pts.roll(1, 1), ch_idx=ch[:2])
    a1, s1 = native(pts.to(DEV), pts.flip(1).to(DEV), pts.roll(1, 1).to(DEV), ch_idx=ch[:2].to(DEV))
    assert rel_l2(a1, a0) < 1e-5 and rel_l2(s1, s0) < 1e-5
    if conn == "add":                                              # no ch_idx: the embedding is simply not added (model.py:58)
        out0 = ren(rx.to(DEV), tx.to(DEV), azi_rand=azi)
        ref0 = render_ref.RenderRef(ref_net, **r)(rx, tx, None, azi_rand=azi)
        assert rel_l2(out0, ref0) < TOL
    else:
        with pytest.raises(ValueError):
            ren(rx.to(DEV), tx.to(DEV), azi_rand=azi)


def test_backward_is_deterministic_and_chunking_is_transparent(built_library):
    cfg = tiny_config("AVRModel", n_azi=8, n_ele=4, n_samples=16, T=200)
    ref_net = field_ref.trained_like_(field_ref.AVRModelRef(cfg["model"], seed=1), seed=2)
    native = _native_from(ref_net, "AVRModel", cfg["model"])
    gen = torch.Generator().manual_seed(0)
    rx, tx = torch.randn(5, 3, generator=gen).to(DEV), torch.randn(5, 3, generator=gen).to(DEV)
    azi = torch.rand(8, generator=gen)
    runs = []
    for chunk in (8, 8, 2):
        ren = avr_b200.AVRRender(native, **cfg["render"], max_receivers_per_pass=chunk, grid_grad="deterministic")
        native.zero_grad(set_to_none=True)
        out = ren(rx, tx, azi_rand=azi)
        out.square().sum().backward()
        runs.append((out.detach().clone(), [p.grad.clone() for p in native.parameters()]))
    assert torch.equal(runs[0][0], runs[1][0])
    for a, b in zip(runs[0][1], runs[1][1]):
        assert torch.equal(a, b)                                   # bit-identical gradients run to run
    assert rel_l2(runs[2][0], runs[0][0]) < 1e-6
    for a, b in zip(runs[2][1], runs[0][1]):
        assert rel_l2(a, b) < 1e-5
    # the default fp32-atomic table gradients agree with the fixed-point ones to rounding
    ren = avr_b200.AVRRender(native, **cfg["render"], grid_grad="atomic")
    native.zero_grad(set_to_none=True)
    ren(rx, tx, azi_rand=azi).square().sum().backward()
    for p, b in zip(native.parameters(), runs[0][1]):
        assert rel_l2(p.grad, b) < 1e-5
    with pytest.raises(ValueError):
        avr_b200.AVRRender(native, **cfg["render"], grid_grad="sorted")


def test_standalone_field_matches_oracle(built_library):
    cfg = tiny_config("AVRModel")
    ref_net = field_ref.trained_like_(field_ref.AVRModelRef(cfg["model"], seed=3), seed=4)
    native = _native_from(ref_net, "AVRModel", cfg["model"])
    gen = torch.Generator().manual_seed(1)
    pts, view, tx = (torch.rand(2, 50, 3, generator=gen) * 2 - 1 for _ in range(3))
    a0, s0 = ref_net(pts, view, tx)
    (a0.sum() + (s0 * s0).sum()).backward()
    a1, s1 = native(pts.to(DEV), view.to(DEV), tx.to(DEV), ch_idx=torch.zeros(2, dtype=torch.long, device=DEV))
    assert rel_l2(a1, a0) < 1e-5 and rel_l2(s1, s0) < 1e-5
    (a1.sum() + (s1 * s1).sum()).backward()
    _check_grads(native, ref_net)


def test_no_grad_inference_and_optimizer_step(built_library):
    cfg = tiny_config("AVRModel")
    native = avr_b200.AVRModel(cfg["model"]).to(DEV)
    ren = avr_b200.AVRRender(native, **cfg["render"]).to(DEV)
    opt = torch.optim.Adam(ren.parameters(), lr=1e-3)
    rx, tx = torch.zeros(2, 3, device=DEV), torch.ones(2, 3, device=DEV)
    with torch.no_grad():
        o = ren(rx, tx)
    assert o.shape == (2, 101, 2) and not o.requires_grad
    before = [p.detach().clone() for p in ren.parameters()]
    loss = ren(rx, tx, ch_idx=torch.tensor([0, 1], device=DEV)).abs().mean()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(ren.parameters(), max_norm=1)
    opt.step()
    assert any(not torch.equal(a, b) for a, b in zip(before, ren.parameters()))
    assert set(ren.state_dict()) == {"network_fn." + k for k in native.state_dict()}


@pytest.mark.parametrize("name", ["simu", "raf_furnished"])
def test_inference_paths_agree(built_library, name):
    """torch.no_grad() skips every buffer only the backward pass reads (saved planes, ReLU bitmasks, bf16 copies) and
    ``graphed_inference`` replays that forward as one CUDA graph: all three give the same spectrum, bit for bit."""
    cfg = get_config(name)
    cfg["render"]["n_azi"], cfg["render"]["n_ele"] = 8, 4
    r = cfg["render"]
    cx = avr_b200.AVRModel_complex if cfg["model_class"] != "AVRModel" else avr_b200.AVRModel
    native = cx(cfg["model"]).to(DEV)
    with torch.no_grad():
        for m in native.modules():
            if isinstance(m, avr_b200.Encoding):
                m.params.normal_(0, 0.1)
    ren = avr_b200.AVRRender(native, **r)
    gen = torch.Generator().manual_seed(2)
    complex_field = cfg["model_class"] != "AVRModel"
    graphed = ren.graphed_inference(2, direction_tx=complex_field)
    for trial in range(3):
        rx = (torch.rand(2, 3, generator=gen) * 2 - 1).to(DEV)
        tx = (torch.rand(2, 3, generator=gen) * 2 - 1).to(DEV)
        dtx = torch.nn.functional.normalize(torch.randn(2, 3, generator=gen), dim=-1).to(DEV) if complex_field else None
        azi = torch.rand(8, generator=gen)
        o_train = ren(rx, tx, dtx, azi_rand=azi)
        with torch.no_grad():
            o_eval = ren(rx, tx, dtx, azi_rand=azi)
        o_graph = graphed(rx, tx, dtx, azi_rand=azi).clone()
        assert o_train.requires_grad and not o_eval.requires_grad
        assert float(o_train.abs().max()) > 0 and bool(torch.isfinite(o_train).all())
        assert torch.equal(o_train.detach(), o_eval) and torch.equal(o_graph, o_eval), trial
    with torch.no_grad():                                               # the graph reads the CURRENT parameters
        o, i = native._model_signal.shapes[-1]
        native._model_signal.params[-o * i:] *= 0.5                     # the (linear) output layer
        assert rel_l2(graphed(rx, tx, dtx, azi_rand=azi), 0.5 * o_eval) < 1e-5
    with pytest.raises(ValueError):
        graphed(rx[:1], tx[:1], dtx[:1] if dtx is not None else None)


def test_full_size_simu_properties(built_library):
    """BASELINE config[1] shape (R=2050, S=64, T=1600): size-independent properties instead of the oracle."""
    cfg = get_config("simu")
    native = avr_b200.AVRModel(cfg["model"]).to(DEV)
    with torch.no_grad():
        for m in native.modules():
            if isinstance(m, avr_b200.Encoding):
                m.params.normal_(0, 0.1)
    ren = avr_b200.AVRRender(native, **cfg["render"], grid_grad="deterministic")
    rx = torch.tensor([[1.0, -2.0, 0.5]], device=DEV)
    tx = torch.tensor([[-1.5, 1.0, 0.0]], device=DEV)
    azi = torch.rand(64)
    out = ren(rx, tx, azi_rand=azi)
    assert out.shape == (1, 801, 2) and bool(torch.isfinite(out).all()) and float(out.abs().max()) > 0
    # (1) linearity in the signal head: scaling the last signal matrix scales the IR
    o, i = native._model_signal.shapes[-1]
    with torch.no_grad():
        native._model_signal.params[-o * i:] *= 2.0
        out2 = ren(rx, tx, azi_rand=azi)
    assert rel_l2(out2, 2 * out) < 1e-5
    # (2) the DC bin of a real sequence is real (the per-sample phase is 1 at f = 0)
    assert abs(float(out[0, 0, 1])) < 1e-6 * float(out.abs().max())
    # (3) two passes are bit-identical, gradients included
    grads = []
    for _ in range(2):
        native.zero_grad(set_to_none=True)
        ren(rx, tx, azi_rand=azi).square().sum().backward()
        grads.append([p.grad.clone() for p in native.parameters()])
    for a, b in zip(*grads):
        assert torch.equal(a, b) and bool(torch.isfinite(a).all())


# (the real fields of the four BASELINE configs against the oracle, on reduced ray grids over ALL candidate seeds and at
# full size: tests/test_gpu_fullsize.py)


@pytest.mark.parametrize("name,bs", [("meshrir", 1), ("raf_furnished", 2), ("real_exp_ch_emb_1", 1)])
def test_full_size_other_configs(built_library, name, bs):
    """BASELINE configs[2..4] shapes at full size: finite, deterministic, linear in the signal head."""
    cfg = get_config(name)
    cls = avr_b200.AVRModel if cfg["model_class"] == "AVRModel" else avr_b200.AVRModel_complex
    native = cls(cfg["model"]).to(DEV)
    with torch.no_grad():
        for m in native.modules():
            if isinstance(m, avr_b200.Encoding):
                m.params.normal_(0, 0.1)
    r = cfg["render"]
    ren = avr_b200.AVRRender(native, **r, grid_grad="deterministic")
    gen = torch.Generator().manual_seed(3)
    c = (r["xyz_min"] + r["xyz_max"]) / 2
    rx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 1.5).to(DEV)
    tx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 1.5).to(DEV)
    dtx = torch.nn.functional.normalize(torch.randn(bs, 3, generator=gen), dim=-1).to(DEV) if cfg["model_class"] != "AVRModel" else None
    ch = torch.arange(bs, device=DEV) % 8
    azi = torch.rand(r["n_azi"], generator=gen)
    T = cfg["model"]["signal_output_dim"]
    grads = []
    for _ in range(2):
        native.zero_grad(set_to_none=True)
        out = ren(rx, tx, dtx, ch_idx=ch, azi_rand=azi)
        assert out.shape == (bs, T // 2 + 1, 2) and bool(torch.isfinite(out).all()) and float(out.abs().max()) > 0
        out.square().sum().backward()
        grads.append([p.grad.clone() for p in native.parameters()])
    for a, b in zip(*grads):
        assert torch.equal(a, b) and bool(torch.isfinite(a).all())
    o, i = native._model_signal.shapes[-1]
    with torch.no_grad():
        native._model_signal.params[-o * i:] *= -0.5
        out2 = ren(rx, tx, dtx, ch_idx=ch, azi_rand=azi)
    assert rel_l2(out2, -0.5 * out) < 1e-5
