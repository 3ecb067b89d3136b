"""Time avr_b200.Criterion (fwd+bwd) against the torch/auraloss-style formulation of the reference run on the same GPU
(oracle/criterion_ref.py moved to cuda) at the simu shape: bs=4, T=1600.  Run on the GPU box."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avr_b200
from avr_b200 import _lib
from oracle import criterion_ref
DEV = "cuda:0"
cfg = {"spec_loss_weight": 1, "amplitude_loss_weight": 0.5, "angle_loss_weight": 0.5, "time_loss_weight": 100,
       "energy_loss_weight": 5, "multistft_loss_weight": 1}
R = {"fs": 16000, "speed": 343.8}
bs, T = 4, 1600
g = torch.Generator().manual_seed(0)
env = torch.exp(-torch.arange(T) / (0.15 * T))
ori = torch.fft.rfft(torch.randn(bs, T, generator=g) * env).to(torch.complex64).to(DEV)
pred0 = torch.fft.rfft(torch.randn(bs, T, generator=g) * env).to(torch.complex64).to(DEV)
def run(crit):
    p = pred0.clone().requires_grad_()
    sum(crit(p, ori)[:8]).backward()
    return p.grad
def timeit(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ours = avr_b200.Criterion(cfg, R)
ref = criterion_ref.CriterionRef(cfg, R).to(DEV)
c0 = _lib.launch_count(); run(ours); launches = _lib.launch_count() - c0
print(json.dumps({"bs": bs, "T": T, "avr_b200_ms": timeit(lambda: run(ours)), "torch_formulation_ms": timeit(lambda: run(ref)),
                  "avr_b200_kernel_launches": launches}))
