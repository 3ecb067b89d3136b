"""A handful of tcgen05 GEMM launches at the simu layer shapes (target of `ncu --set full`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avr_b200 import ops
from avr_b200.ops import PlanePair
DEV = "cuda:0"
NP = 524800
def pp(r, c, n=2):
    x = PlanePair.empty(r, c, DEV, n=n); x.buf.normal_(); return x
# launch order: [0] fwd 512x512 (6 products)  [1] fwd 512, K=128  [2] bwd-data 512 (mask)  [3] dW 512x512 (TN)
#               [4] fwd 128x128 (6 products, B resident)  [5] dW 128x128 (TN)
a3, w3, c3 = pp(NP, 512, 3), pp(512, 512, 3), pp(NP, 512, 3)
bits = ops.relu_bits_empty(NP, 512, DEV)
ops.umma_nt(a3, w3, ops.UMMA_RELU, c3, bits_out=bits)
a1, w1k = pp(NP, 128, 3), pp(512, 128, 3)
ops.umma_nt(a1, w1k, ops.UMMA_RELU, c3, bits_out=bits)
g2, wt2, d2 = pp(NP, 512), pp(512, 512), pp(NP, 512)
ops.umma_nt(g2, wt2, ops.UMMA_MASK, d2, mask=bits)
dw = torch.empty(512, 512, device=DEV)
ws = torch.empty(max(4, ops.umma_tn_workspace_bytes(512, 512, NP) // 4), device=DEV)
ops.umma_tn(g2, a3, dw, ws)
del a3, c3, g2, d2
w1, c1 = pp(128, 128, 3), pp(NP, 128, 3)
bits1 = ops.relu_bits_empty(NP, 128, DEV)
ops.umma_nt(a1, w1, ops.UMMA_RELU, c1, bits_out=bits1)
g1 = pp(NP, 128)
dw1 = torch.empty(128, 128, device=DEV)
ops.umma_tn(g1, a1, dw1, ws)
torch.cuda.synchronize()
print("ok")
