"""Where does the fp16-pair forward GEMM (524800 x 512 x 512, ReLU, 3 products) spend its time?  Experiment build only
(python -m avr_b200.build --experiments --force): tile shape forced through AVR_UMMA_F16_BN128, epilogue stages switched
off through AVR_UMMA_DEBUG (128 no wait on the previous bulk store, 256 no bulk stores, 512 no conversion / staging,
1024 no proxy fence) -- results are then garbage, timing only."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avr_b200 import ops
from avr_b200.ops import PlanePair
DEV = "cuda:0"
M, N = 524800, 512


def timeit(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for K in (512, 208):
    a = PlanePair.empty(M, K, DEV, kind=ops.PLANES_F16x2); a.buf.normal_()
    b = PlanePair.empty(N, K, DEV, kind=ops.PLANES_F16x2); b.buf.normal_()
    c = PlanePair.empty(M, N, DEV, kind=ops.PLANES_F16x2)
    c2 = PlanePair.empty(M, N, DEV, kind=ops.PLANES_BF16x2)
    bits = torch.empty(M, N // 32, dtype=torch.int32, device=DEV)
    for bn128 in ("", "1"):
        if bn128: os.environ["AVR_UMMA_F16_BN128"] = "1"
        else: os.environ.pop("AVR_UMMA_F16_BN128", None)
        for dual in (0, 1):
            row = {"K": K, "bn128": bool(bn128), "dual_copy": dual}
            for name, dbg in [("full", 0), ("nowait", 128), ("nostore", 256 + 128), ("nofence", 1024), ("nostage", 512),
                              ("no_epilogue", 512 + 256 + 128 + 1024)]:
                os.environ["AVR_UMMA_DEBUG"] = str(dbg)
                if dual:
                    fn = lambda: ops.umma_nt(a, b, ops.UMMA_RELU | ops.UMMA_DUAL_COPY, c, c2=c2, bits_out=bits)
                else:
                    fn = lambda: ops.umma_nt(a, b, ops.UMMA_RELU, c, bits_out=bits)
                row[name] = round(timeit(fn), 4)
            os.environ["AVR_UMMA_DEBUG"] = "0"
            print(json.dumps(row), flush=True)
    del a, b, c, c2
