import sys, torch
sys.path.insert(0, '.')
import bench
dev = torch.device("cuda:0")
w = bench.Workload("simu", 4, "train", "deterministic", dev, 0, 1)
for _ in range(3): w.step(False)
for rnd in range(4):
    for host_io in (False, True):
        w.step(host_io)
        ms, _, _ = w.timed(host_io, 20)
        print("e2e  " if host_io else "value", round(ms / 20, 3), flush=True)
