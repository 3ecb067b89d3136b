"""Timeline of one CTA of the fused chain kernel (experiment build: AVR_CHAIN_TRACE_PTR): clock64() stamps of the MMA thread
and of two epilogue warps, printed as per-layer deltas.  python -m avr_b200.build --experiments first."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
trace = torch.zeros(3 * 4096, dtype=torch.int64, device="cuda:0")
os.environ["AVR_CHAIN_TRACE_PTR"] = hex(trace.data_ptr())
os.environ["CHAIN_REPS"] = "1"
import runpy                                             # noqa: E402
trace.zero_()
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_chain_once.py"))
torch.cuda.synchronize()
t = trace.cpu().view(3, 4096)
# the script launches 2 warm-ups + 1 timed call; every call rewrites from index 0 -> the last call's stamps
mma, e0, e1 = t[0], t[1], t[2]
n_layers = 8
base = int(mma[0])
print("tile layer |  MMA: kb0 ready, kb0 issued, kb1 ready, kb1 issued  | epi warp2: acc seen h0, handed h0, acc seen h1, handed h1 | epi warp6 ...   (cycles since the tile's first event)")
for tile in range(4):
    for l in range(n_layers):
        nk = 2 if l > 0 else 1
        # MMA stamps: 2 per k-block; layer 0 has one k-block
        off = tile * (2 * (2 * n_layers - 1)) + (0 if l == 0 else 2 + (l - 1) * 4)
        m = [int(mma[off + i]) - base for i in range(2 * nk)]
        eo = (tile * n_layers + l) * 4
        a = [int(e0[eo + i]) - base for i in range(4)]
        b = [int(e1[eo + i]) - base for i in range(4)]
        print(f"{tile:3d} {l:3d}   | " + " ".join(f"{x:8d}" for x in m) + ("                  " if nk == 1 else "") + " | " + " ".join(f"{x:8d}" for x in a) + " | " + " ".join(f"{x:8d}" for x in b))
