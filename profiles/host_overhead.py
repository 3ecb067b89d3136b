"""How long does the host take to enqueue one fused step vs how long the GPU takes to run it?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, avr_b200
from avr_b200.configs import get_config
cfg = get_config("simu")
dev = torch.device("cuda:0")
field = avr_b200.AVRModel(cfg["model"]).to(dev)
ren = avr_b200.AVRRender(field, **cfg["render"], max_receivers_per_pass=4)
arena = avr_b200.GradArena(ren.parameters())
rx = torch.rand(4, 3, device=dev) * 4 - 2
tx = torch.rand(4, 3, device=dev) * 4 - 2
def step():
    arena.zero_()
    out = ren(rx, tx)
    out.square().sum().backward()
for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0):.2f} ms, until GPU done {1e3*(t2-t0):.2f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); step(); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
