"""A/B on one box: weight gradients of the signal network from the fp16-pair activation (converted in the kernel) against
the second, bf16 copy written by the forward layers (fused_tc.SIG_TN_FROM_F16)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from avr_b200 import fused_tc
dev = torch.device("cuda:0")
for rnd in range(3):
    for flag in (True, False):
        fused_tc.SIG_TN_FROM_F16 = flag
        w = bench.Workload("simu", 4, "train", "deterministic", dev, 0, 1)
        for _ in range(3): w.step(False)
        ms, launches, _ = w.timed(False, 20)
        print("converted in the kernel" if flag else "second copy from the forward layers", round(ms / 20, 3), "ms/step", flush=True)
        w.close()
