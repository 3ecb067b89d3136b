"""One fp16-pair forward GEMM of the simu signal network (524800 x 512 x 512, ReLU) for ncu; also the bf16x3 one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avr_b200 import ops
from avr_b200.ops import PlanePair
DEV = "cuda:0"
M, N, K = 524800, 512, 512
g = torch.Generator(device=DEV).manual_seed(0)
A = torch.randn(M, K, device=DEV, generator=g).clamp_min(0)
W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
for kind in (ops.PLANES_F16x2, ops.PLANES_BF16x3):
    a = ops.planes_split(A, PlanePair.empty(M, K, DEV, kind=kind))
    w = ops.planes_split(W, PlanePair.empty(N, K, DEV, kind=kind))
    c = PlanePair.empty(M, N, DEV, kind=kind)
    for _ in range(3):
        ops.umma_nt(a, w, ops.UMMA_RELU, c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.umma_nt(a, w, ops.UMMA_RELU, c)
    e1.record(); torch.cuda.synchronize()
    print(f"kind {kind}: {e0.elapsed_time(e1) / 5:.3f} ms")
    del a, w, c
