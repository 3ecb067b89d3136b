"""Pin the oracle against the UNMODIFIED reference module (only where /root/reference exists)."""
import pytest
import torch

from avr_b200.configs import tiny_config
from oracle import field_ref, render_ref
from oracle.reference_shim import load_reference_renderer, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")


@pytest.mark.parametrize("model_class,box,fs", [("AVRModel", (-10, 10), 16000), ("AVRModel", (0, 10), 4000),
                                                 ("AVRModel_complex", (-12, 12), 8000)])
def test_render_bit_exact(model_class, box, fs):
    ref = load_reference_renderer()
    cfg = tiny_config(model_class, xyz_min=box[0], xyz_max=box[1], fs=fs)
    cls = field_ref.AVRModelRef if model_class == "AVRModel" else field_ref.AVRModelComplexRef
    net = field_ref.trained_like_(cls(cfg["model"]))
    g = torch.Generator().manual_seed(3)
    c = (box[0] + box[1]) / 2
    rx = c + (torch.rand(2, 3, generator=g) * 2 - 1) * 2
    tx = c + (torch.rand(2, 3, generator=g) * 2 - 1) * 2
    dtx = torch.nn.functional.normalize(torch.randn(2, 3, generator=g), dim=-1) if model_class != "AVRModel" else None
    r0, r1 = ref.AVRRender(net, **cfg["render"]), render_ref.RenderRef(net, **cfg["render"])
    with torch.no_grad():
        torch.manual_seed(9)
        o0 = r0(rx, tx, dtx) if dtx is not None else r0(rx, tx)
        torch.manual_seed(9)
        o1 = r1(rx, tx, dtx)
    assert float(o0.abs().max()) > 0
    assert torch.equal(o0, o1)


def test_direction_table_bit_exact():
    ref = load_reference_renderer()
    for n_azi, n_ele in ((6, 3), (64, 32), (80, 40), (36, 18)):
        torch.manual_seed(n_azi)
        d0, _, _ = ref.ray_directions(n_azi, n_ele)
        torch.manual_seed(n_azi)
        d1 = render_ref.direction_table(n_azi, n_ele)
        assert torch.equal(d0, d1)
        # generator state advanced identically (rand(n_azi) then rand(n_ele))
        torch.manual_seed(n_azi)
        ref.ray_directions(n_azi, n_ele)
        a = torch.rand(1)
        torch.manual_seed(n_azi)
        render_ref.direction_table(n_azi, n_ele)
        assert torch.equal(a, torch.rand(1))


def _embed(conn, enc, dec, sig):
    return {"is_embed": True, "ch_num": 8, "connection_type": conn, "is_sigma_encoder": enc, "is_sigma_decoder": dec,
            "is_signal_network": sig, "emb_dim_sigma_encoder": 8, "emb_dim_sigma_decoder": 16, "emb_dim_signal_network": 24}


@pytest.mark.parametrize("embed", [None, _embed("add", True, False, True), _embed("add", True, True, True),
                                   _embed("concat", True, True, True), _embed("concat", False, True, False),
                                   {"is_embed": True, "ch_num": 8}])          # avr_real_exp_ch_emb_1.yml: no connection_type
def test_field_restatement_matches_reference_model_py(embed):
    """``oracle/field_ref.py`` composes the tcnn pieces exactly like the reference's own ``model.py`` does (executed
    unmodified, with the oracle's HashGridRef / MLPRef standing in for the absent tiny-cuda-nn): same parameter names
    and shapes, bit-identical outputs and gradients -- channel-embedding variants (model.py:11-61,71-228) included."""
    from oracle.reference_shim import load_reference_model
    ref_model = load_reference_model()
    cfg = tiny_config("AVRModel")
    if embed is not None:
        cfg["model"]["channel_embed"] = embed
    theirs = ref_model.AVRModel(cfg["model"])
    ours = field_ref.AVRModelRef(cfg["model"], seed=5)
    assert {k: tuple(v.shape) for k, v in theirs.state_dict().items()} == {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    field_ref.trained_like_(ours, seed=6)
    theirs.load_state_dict(ours.state_dict())
    g = torch.Generator().manual_seed(1)
    pts, view, tx = (torch.rand(3, 30, 3, generator=g) * 2 - 1 for _ in range(3))
    ch = torch.tensor([2, 7, 2])
    a0, s0 = theirs(pts, view, tx, ch_idx=ch)
    a1, s1 = ours(pts, view, tx, ch_idx=ch)
    assert a0.shape == (3, 30, 1) and torch.equal(a0, a1) and torch.equal(s0, s1) and float(s0.abs().max()) > 0
    G = torch.randn(s0.shape, generator=g)
    (a0.sum() + (s0 * G).sum()).backward()
    (a1.sum() + (s1 * G).sum()).backward()
    mine = dict(ours.named_parameters())
    for name, p in theirs.named_parameters():
        assert torch.equal(p.grad, mine[name].grad), name


def test_complex_field_restatement_matches_reference_model_py():
    from oracle.reference_shim import load_reference_model
    ref_model = load_reference_model()
    cfg = tiny_config("AVRModel_complex")
    theirs = ref_model.AVRModel_complex(cfg["model"])
    ours = field_ref.trained_like_(field_ref.AVRModelComplexRef(cfg["model"], seed=5), seed=6)
    assert {k: tuple(v.shape) for k, v in theirs.state_dict().items()} == {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    theirs.load_state_dict(ours.state_dict())
    g = torch.Generator().manual_seed(2)
    pts, view, tx, tv = (torch.rand(2, 25, 3, generator=g) * 2 - 1 for _ in range(4))
    a0, s0 = theirs(pts, view, tx, tv)
    a1, s1 = ours(pts, view, tx, tv)
    assert torch.equal(a0, a1) and torch.equal(s0, s1) and float(s0.abs().max()) > 0


TRAIN_CFG = {"spec_loss_weight": 1, "amplitude_loss_weight": 0.5, "angle_loss_weight": 0.5, "time_loss_weight": 100,
             "energy_loss_weight": 5, "multistft_loss_weight": 1}                    # avr_meshrir.yml:35-40


@pytest.mark.parametrize("bs,T,das", [(4, 1600, False), (3, 400, False), (8, 1600, True)])
def test_criterion_restatement_matches_reference(bs, T, das):
    """oracle/criterion_ref.py vs the unmodified utils/criterion.py (stand-in auraloss): every output, and d/d pred."""
    from oracle import criterion_ref
    from oracle.reference_shim import load_reference_criterion
    cfg = dict(TRAIN_CFG)
    if das:
        cfg.update(das_reg_loss_weight=0.3, das_ce_loss_weight=0.2, beta=100.0)
    render = {"fs": 16000, "speed": 343.8}
    theirs = load_reference_criterion().Criterion(cfg, render)
    ours = criterion_ref.CriterionRef(cfg, render)
    g = torch.Generator().manual_seed(bs)
    F_ = T // 2 + 1
    pred = torch.complex(torch.randn(bs, F_, generator=g), torch.randn(bs, F_, generator=g)).requires_grad_()
    pred2 = pred.detach().clone().requires_grad_()
    ori = torch.complex(torch.randn(bs, F_, generator=g), torch.randn(bs, F_, generator=g))
    a, b = theirs(pred, ori), ours(pred2, ori)
    assert len(a) == len(b) == 10
    for x, y in zip(a, b):
        assert torch.allclose(x, y, rtol=1e-6, atol=1e-7)
    sum(a[:8]).backward()
    sum(b[:8]).backward()
    assert torch.allclose(pred.grad, pred2.grad, rtol=1e-5, atol=1e-8)
