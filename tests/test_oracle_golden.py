"""The oracle restatement reproduces the golden vectors captured from the unmodified reference."""
import pytest
import torch

from oracle import render_ref
from tests.helpers import GOLDEN_CASES, case_config, load_golden, oracle_field, rel_l2


class _Stub(torch.nn.Module):
    def __init__(self, attn, signal):
        super().__init__()
        self.attn = torch.nn.Parameter(attn.clone())
        self.signal = torch.nn.Parameter(signal.clone())
        self.signal_output_dim = signal.shape[-1]

    def forward(self, pts, view, tx, dir_tx=None):
        return self.attn, self.signal


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_golden(name):
    g = load_golden(name)
    model_class, cfg = case_config(name)
    if model_class == "stub":
        net = _Stub(g["param/attn"], g["param/signal"])
    else:
        net = oracle_field(model_class, cfg["model"], g)
    ren = render_ref.RenderRef(net, **cfg["render"])
    out = ren(g["rx"], g["tx"], g.get("dir_tx"), azi_rand=g["azi_rand"])
    assert torch.equal(render_ref.direction_table(cfg["render"]["n_azi"], cfg["render"]["n_ele"], g["azi_rand"]), g["dirs"])
    assert rel_l2(out, g["out"]) < 1e-6
    (out * g["G"]).sum().backward()
    for pname, p in net.named_parameters():
        assert rel_l2(p.grad, g["grad/" + pname]) < 1e-5, pname


@pytest.mark.parametrize("name", GOLDEN_CASES[:3])
def test_oracle_geometry_bit_exact(name):
    g = load_golden(name)
    _, cfg = case_config(name)
    r = cfg["render"]
    tab = render_ref.static_tables(r, cfg["model"]["signal_output_dim"])
    pts_n, view, tx_n, _ = render_ref.sample_geometry(g["rx"], g["tx"], g["dirs"], tab["d"], r)
    assert torch.equal(pts_n, g["net_pts"])
    step = max(1, view.shape[1] // 16)
    assert torch.equal(view[:, ::step], g["net_view"])
    assert torch.equal(tx_n[:, ::step], g["net_tx"])


def test_reordered_composite_equals_literal():
    g = load_golden("stub_renderer_only")
    _, cfg = case_config("stub_renderer_only")
    net = _Stub(g["param/attn"], g["param/signal"])
    ren = render_ref.RenderRef(net, **cfg["render"])
    ren.reordered = True
    out = ren(g["rx"], g["tx"], azi_rand=g["azi_rand"])
    assert rel_l2(out, g["out"]) < 2e-6
