"""The fp16-pair forward GEMMs of the simu signal network (524800 x 512 x 512, ReLU + bitmask; the same with a second,
bf16-pair copy of the output) for ncu."""
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
kind = ops.PLANES_F16x2
a = ops.planes_split(A, PlanePair.empty(M, K, DEV, kind=kind))
w = ops.planes_split(W, PlanePair.empty(N, K, DEV, kind=kind))
del A, W
c = PlanePair.empty(M, N, DEV, kind=kind)
c2 = PlanePair.empty(M, N, DEV, kind=ops.PLANES_BF16x2)
bits = torch.empty(M, N // 32, dtype=torch.int32, device=DEV)
for name, fn in (("relu+bits", lambda: ops.umma_nt(a, w, ops.UMMA_RELU, c, bits_out=bits)),
                 ("dual_copy", lambda: ops.umma_nt(a, w, ops.UMMA_RELU | ops.UMMA_DUAL_COPY, c, c2=c2, bits_out=bits))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 5:.3f} ms")
