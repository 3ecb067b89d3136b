import sys, torch
sys.path.insert(0, '.')
import bench
from avr_b200 import fused_tc
dev = torch.device("cuda:0")
for rnd in range(2):
    for fuse in (True, False):
        fused_tc.FUSE_SIGMA_CHAIN = fuse
        w = bench.Workload("real_exp_ch_emb_1", 8, "train", "deterministic", dev, 0, 1)
        for _ in range(3): w.step(False)
        ms, launches, _ = w.timed(False, 10)
        print("fused chain" if fuse else "layer by layer", round(ms / 10, 3), "ms/step", launches // 10, "launches", flush=True)
        w.close()
