"""Where does the plane-output epilogue of the tcgen05 GEMM spend its time?  (run on the GPU box)

Times three layer shapes with parts of the epilogue switched off through AVR_UMMA_DEBUG (results are then
garbage; timing only): 128 = no wait on the previous bulk store, 256 = no bulk stores, 512 = no
conversion/staging, 1024 = no async-proxy fence.
"""
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
for (M, N, K, npl) in [(NP, 512, 208, 3), (NP, 128, 128, 3), (NP, 512, 64, 3), (NP, 512, 512, 3), (NP, 512, 128, 2), (NP, 512, 512, 2)]:
    a = PlanePair.empty(M, K, DEV, n=npl); a.buf.normal_()
    b = PlanePair.empty(N, K, DEV, n=npl); b.buf.normal_()
    c = PlanePair.empty(M, N, DEV, n=npl)
    row = {"M": M, "N": N, "K": K, "planes": npl}
    for name, dbg in [("full", 0), ("nowait", 128), ("nostore", 256 + 128), ("nostage", 512), ("nofence", 1024),
                      ("nostage_nostore", 512 + 256 + 128 + 1024)]:
        os.environ["AVR_UMMA_DEBUG"] = str(dbg)
        row[name] = round(timeit(lambda: ops.umma_nt(a, b, ops.UMMA_RELU, c)), 4)
    os.environ["AVR_UMMA_DEBUG"] = "0"
    print(json.dumps(row), flush=True)
    del a, b, c
