"""Micro-benchmark of the tcgen05 plane-pair GEMMs at the simu layer shapes (run on the GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avr_b200 import ops
from avr_b200.ops import PlanePair

DEV = "cuda:0"
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

NP = 524800
res = []
for (M, N, K) in [(NP, 512, 512), (NP, 512, 208), (NP, 128, 128), (NP, 128, 48), (NP, 16, 128), (NP, 1600, 512), (NP, 208, 512)]:
    a = PlanePair.empty(M, K, DEV); a.buf.normal_()
    b = PlanePair.empty(N, K, DEV); b.buf.normal_()
    c = PlanePair.empty(M, N, DEV)
    ms = timeit(lambda: ops.umma_nt(a, b, ops.UMMA_RELU, c))
    res.append({"kind": "nt", "M": M, "N": N, "K": K, "ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9,
                "GBs": (M * K * 4 + M * N * 4) / ms / 1e6})
    del a, b, c
for (M, N, K) in [(512, 512, NP), (512, 208, NP), (128, 128, NP), (128, 48, NP), (16, 128, NP), (1600, 512, NP)]:
    a = PlanePair.empty(K, M, DEV); a.buf.normal_()
    b = PlanePair.empty(K, N, DEV); b.buf.normal_()
    c = torch.empty(M, N, device=DEV)
    ws = torch.empty(max(4, ops.umma_tn_workspace_bytes(M, N, K) // 4), device=DEV)
    ms = timeit(lambda: ops.umma_tn(a, b, c, ws))
    res.append({"kind": "tn", "M": M, "N": N, "K": K, "ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9,
                "GBs": (K * M * 4 + K * N * 4) / ms / 1e6})
    del a, b, c
for r in res:
    print(json.dumps(r))
