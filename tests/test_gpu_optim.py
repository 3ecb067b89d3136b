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
