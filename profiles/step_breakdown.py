"""Per-call CUDA-event timing of one simu render step (fwd+bwd, 4 receivers) in call order: every timed op of
avr_b200.ops with its work and rate.   python profiles/step_breakdown.py [config] [bs] [atomic|deterministic]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avr_b200
from avr_b200 import ops
from avr_b200.configs import get_config
DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "simu"
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = get_config(name)
cls = avr_b200.AVRModel if cfg["model_class"] == "AVRModel" else avr_b200.AVRModel_complex
net = cls(cfg["model"]).to(DEV)
with torch.no_grad():
    for n, p in net.named_parameters():
        if "encoding" in n:
            p.normal_(0, 0.1)
r = cfg["render"]
ren = avr_b200.AVRRender(net, **r, grid_grad=(sys.argv[3] if len(sys.argv) > 3 else None))
c = (r["xyz_min"] + r["xyz_max"]) / 2
gen = torch.Generator().manual_seed(0)
rx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 2).to(DEV); tx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * 2).to(DEV)
dtx = torch.nn.functional.normalize(torch.randn(bs, 3, generator=gen), dim=-1).to(DEV) if cfg["model_class"] != "AVRModel" else None
def step():
    net.zero_grad(set_to_none=True)
    out = ren(rx, tx, dtx)
    out.square().sum().backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
reps = 5
acc = None
for _ in range(reps):
    ops.PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    rows = [(n, w, u, s.elapsed_time(e), ex) for n, w, u, s, e, ex in prof]
    tot = e0.elapsed_time(e1)
    if acc is None:
        acc = [[n, w, u, 0.0, ex] for n, w, u, _, ex in rows]; total = 0.0
    for a, rrow in zip(acc, rows):
        a[3] += rrow[3] / reps
    total += tot / reps
print(f"{name} bs={bs}: step {total:.3f} ms; timed ops {sum(a[3] for a in acc):.3f} ms")
for n, w, u, ms, ex in acc:
    rate = w / ms / 1e9 if ms > 0 else 0
    extra = f"  pipe {ex / ms / 1e9:8.1f} TFLOP/s" if u == "flop" and ex else ""
    print(f"  {n:20s} {ms:7.3f} ms   work {w:.3e} {u:5s} -> {rate:9.1f} {'TFLOP/s' if u == 'flop' else 'GB/s'}{extra}")
