"""Per-resolution check of the multi-resolution STFT gradient against torch autograd (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avr_b200
from avr_b200 import criterion as C
from oracle import criterion_ref
from oracle.render_ref import rel_l2
DEV = "cuda:0"
cfg = {"spec_loss_weight": 0, "amplitude_loss_weight": 0, "angle_loss_weight": 0, "time_loss_weight": 0, "energy_loss_weight": 0,
       "multistft_loss_weight": 1}
R = {"fs": 16000, "speed": 343.8}
for bs, T in [(1, 400), (3, 400), (4, 1600), (5, 800), (2, 2400)]:
    gen = torch.Generator().manual_seed(T + bs)
    env = torch.exp(-torch.arange(T) / (0.2 * T))
    ori = torch.fft.rfft(torch.randn(bs, T, generator=gen) * env).to(torch.complex64)
    pred0 = torch.fft.rfft(torch.randn(bs, T, generator=gen) * env).to(torch.complex64)
    for res in C.MRSTFT_RESOLUTIONS:
        saved = C.MRSTFT_RESOLUTIONS
        C.MRSTFT_RESOLUTIONS = (res,)
        ours = avr_b200.Criterion(cfg, R)
        ref = criterion_ref.CriterionRef(cfg, R)
        ref.mrstft = criterion_ref.MultiResolutionSTFTLossRef(w_lin_mag=1, fft_sizes=[res[0]], hop_sizes=[res[1]], win_lengths=[res[2]])
        p1 = pred0.clone().requires_grad_(); p2 = pred0.clone().to(DEV).requires_grad_()
        l1 = ref(p1, ori)[5]; l2 = ours(p2, ori.to(DEV))[5]
        l1.backward(); l2.backward()
        g1 = torch.fft.irfft(p1.grad); g2 = torch.fft.irfft(p2.grad.cpu())
        d = (g1 - g2).abs()
        print(bs, T, res, f"loss {float(l1):.6f} {float(l2):.6f} grad err {rel_l2(torch.view_as_real(p2.grad).cpu(), torch.view_as_real(p1.grad)):.2e}",
              "worst t", [int(i) for i in d.max(0).values.topk(4).indices])
        C.MRSTFT_RESOLUTIONS = saved
