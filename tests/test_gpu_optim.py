"""Fused clip + scrub + Adam against the reference sequence (avr_runner.py:192-200) run with torch on the CPU."""
import pytest
import torch

import avr_b200
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference_step(params, opt):
    torch.nn.utils.clip_grad_norm_(params, max_norm=1)
    for p in params:
        if p.grad is not None:
            with torch.no_grad():
                p.grad[p.grad != p.grad] = 0
                p.grad[torch.isinf(p.grad)] = 0
    opt.step()


@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_fused_adam_matches_reference_sequence(built_library, weight_decay):
    g = torch.Generator().manual_seed(0)
    shapes = [(1000,), (37, 5), (4099,), (3,)]
    ref = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    ours = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref]
    opt_ref = torch.optim.Adam(ref, lr=5e-4, weight_decay=weight_decay)
    opt = avr_b200.FusedAdam(ours, lr=5e-4, weight_decay=weight_decay, max_norm=1.0)
    for it in range(6):
        scale = [0.01, 3.0, 50.0, 1e-4, 1.0, 1.0][it]                  # below and above the clipping threshold
        grads = [torch.randn(s, generator=g) * scale for s in shapes]
        if it == 4:
            grads[1][3, 2] = float("nan")                                # poisons the clip coefficient -> zero step
        opt.zero_grad()
        for p, q, gr in zip(ref, ours, grads):
            p.grad = gr.clone()
            q.grad.copy_(gr.to(DEV))
        total = torch.linalg.vector_norm(torch.cat([x.reshape(-1) for x in grads]))
        _reference_step(ref, opt_ref)
        opt.step(write_back_grad=True)
        if it != 4:
            assert abs(float(opt.grad_norm()) - float(total)) < 1e-4 * float(total)
        for p, q in zip(ref, ours):
            assert rel_l2(q, p) < 1e-6, it
            assert torch.allclose(q.grad.cpu(), p.grad, rtol=1e-5, atol=1e-8), it
    # parameters are views of one flat buffer and stay usable as ordinary nn.Parameters
    assert ours[0].data_ptr() == opt.flat_params.data_ptr()
    assert ours[1].grad.data_ptr() == opt.arena.flat[1000:].data_ptr()


def test_fused_adam_trains_the_renderer(built_library):
    from avr_b200.configs import tiny_config
    cfg = tiny_config("AVRModel")
    net = avr_b200.AVRModel(cfg["model"])
    gen = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, avr_b200.Encoding):
                m.params.copy_(torch.randn(m.params.shape, generator=gen) * 0.1)
    net = net.to(DEV)
    ren = avr_b200.AVRRender(net, **cfg["render"])
    opt = avr_b200.FusedAdam(ren.parameters(), lr=1e-3)
    rx, tx = torch.zeros(2, 3, device=DEV), torch.ones(2, 3, device=DEV)
    azi = torch.rand(cfg["render"]["n_azi"], generator=gen)
    target = (torch.randn(2, 101, 2, generator=gen) * 1e-2).to(DEV)
    losses = []
    for _ in range(40):
        opt.zero_grad()
        loss = (ren(rx, tx, azi_rand=azi) - target).square().mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.9 * losses[0]
    assert set(ren.state_dict()) == {"network_fn." + k for k in net.state_dict()}


def test_reference_format_checkpoint_round_trip(built_library, tmp_path):
    """avr_runner.py:104-154: a checkpoint written the reference's way (stock ``torch.optim.Adam`` + cosine schedule over
    the oracle field, reference key names) resumes under ``AVRRender`` + ``FusedAdam`` and follows the same trajectory;
    the file written back loads into stock ``torch.optim.Adam`` again."""
    from avr_b200.configs import tiny_config
    from oracle import field_ref, render_ref
    cfg = tiny_config("AVRModel")
    gen = torch.Generator().manual_seed(4)
    rx, tx = torch.randn(2, 3, generator=gen), torch.randn(2, 3, generator=gen)
    azi = torch.rand(cfg["render"]["n_azi"], generator=gen)
    target = torch.randn(2, 101, 2, generator=gen) * 1e-3

    def make_opt(params):
        opt = torch.optim.Adam(params, lr=2e-3, weight_decay=1e-4, betas=(0.9, 0.999))
        return opt, torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=20.0, eta_min=1e-5)

    # the "reference" run on the CPU: oracle field inside the oracle renderer, stock optimiser, 3 steps, save
    ref_ren = render_ref.RenderRef(field_ref.trained_like_(field_ref.AVRModelRef(cfg["model"], seed=2), seed=3), **cfg["render"])
    ref_opt, ref_sched = make_opt(ref_ren.parameters())

    def ref_step():
        ref_opt.zero_grad()
        (ref_ren(rx, tx, None, azi_rand=azi) - target).square().sum().backward()
        torch.nn.utils.clip_grad_norm_(ref_ren.parameters(), max_norm=1)
        ref_opt.step()
        ref_sched.step()

    for _ in range(3):
        ref_step()
    (tmp_path / "ckpts").mkdir()
    path = str(tmp_path / "ckpts" / "000003.tar")
    torch.save({"current_iteration": 3, "audionerf_network_state_dict": ref_ren.state_dict(),
                "optimizer_state_dict": ref_opt.state_dict(), "scheduler_state_dict": ref_sched.state_dict()}, path)
    assert avr_b200.latest_checkpoint(str(tmp_path / "ckpts")) == path

    # resume natively
    net = avr_b200.AVRModel(cfg["model"]).to(DEV)
    ren = avr_b200.AVRRender(net, **cfg["render"])
    opt = avr_b200.FusedAdam(ren.parameters(), lr=2e-3, weight_decay=1e-4, max_norm=1.0)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=20.0, eta_min=1e-5)
    assert avr_b200.load_checkpoint(path, ren, opt, sched, map_location=DEV) == 3
    assert opt.step_count == 3 and abs(opt.lr - ref_opt.param_groups[0]["lr"]) < 1e-12
    for _ in range(2):
        ref_step()
        opt.zero_grad()
        (ren(rx.to(DEV), tx.to(DEV), azi_rand=azi) - target.to(DEV)).square().sum().backward()
        opt.step()
        sched.step()
    ref_params = dict(ref_ren.named_parameters())
    for name, p in ren.named_parameters():
        assert rel_l2(p, ref_params[name]) < 2e-5, name
    # and back: the file we write is a stock-Adam checkpoint again
    out = avr_b200.save_checkpoint(str(tmp_path / "ckpts" / "000005.tar"), ren, opt, sched, current_iteration=5)
    ref_ren2 = render_ref.RenderRef(field_ref.AVRModelRef(cfg["model"], seed=9), **cfg["render"])
    opt2, sched2 = make_opt(ref_ren2.parameters())
    ck = torch.load(out, map_location="cpu", weights_only=False)
    ref_ren2.load_state_dict(ck["audionerf_network_state_dict"])
    opt2.load_state_dict(ck["optimizer_state_dict"])
    sched2.load_state_dict(ck["scheduler_state_dict"])
    assert ck["current_iteration"] == 5 and int(opt2.state[next(iter(ref_ren2.parameters()))]["step"]) == 5
    assert abs(opt2.param_groups[0]["lr"] - ref_opt.param_groups[0]["lr"]) < 1e-12
