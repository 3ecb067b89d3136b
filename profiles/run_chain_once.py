"""One launch set of the fused sigma chain (simu shapes: 524 800 rows, 48 -> 128 x4 -> 128 x3 -> 16) for ncu and timing
experiments.  AVR_CHAIN_DEBUG (experiment builds only) switches parts of the kernel off."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avr_b200 import ops                                  # noqa: E402
from avr_b200.ops import PlanePair                        # noqa: E402

DEV = "cuda:0"
M = int(os.environ.get("CHAIN_ROWS", 524800))
g = torch.Generator().manual_seed(0)
K3 = ops.PLANES_BF16x3


def planes(x, kind=K3):
    return ops.planes_split(x.to(DEV), PlanePair.empty(x.shape[0], x.shape[1], DEV, kind=kind))


x0 = planes(torch.randn(M, 48, generator=g))
dims = [48, 128, 128, 128, 128, 128, 128, 128, 16]
ws = [planes(torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5) for i in range(8)]
layers = []
for i in range(8):
    L = {"w": ws[i], "relu": i < 7}
    if i == 7:
        L["out_f32"] = torch.empty(M, 16, device=DEV)
    else:
        L["save"] = PlanePair.empty(M, 128, DEV, kind=ops.PLANES_BF16x2 if i < 3 else K3)
        L["bits"] = ops.relu_bits_empty(M, 128, DEV)
        if i == 3:
            L["save_raw"] = PlanePair.empty(M, 128, DEV, kind=K3)
    layers.append(L)
reps = int(os.environ.get("CHAIN_REPS", 5))
for _ in range(2):
    ops.mlp_chain(x0, layers)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.mlp_chain(x0, layers)
e1.record()
torch.cuda.synchronize()
print(f"chain debug={os.environ.get('AVR_CHAIN_DEBUG', '0')} rows {M}: {e0.elapsed_time(e1) / reps:.3f} ms per launch")
