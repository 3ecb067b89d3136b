"""Where does the pair-mode epilogue of the fp16-pair forward GEMM (524800 x 512 x 512) spend its time?  Experiment build:
AVR_UMMA_DEBUG bits 128 no wait on the staging buffer, 256 no bulk stores, 512 no conversion / staging, 1024 no proxy fence;
AVR_UMMA_EPI_SPLIT=1 one epilogue warp per lane group.  Interleaved rounds, minimum reported."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avr_b200 import ops
from avr_b200.ops import PlanePair
DEV = "cuda:0"
M, N, K = 524800, 512, 512


def timeit(fn, n=6):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


kind = ops.PLANES_F16x2
a = PlanePair.empty(M, K, DEV, kind=kind); a.buf.normal_()
b = PlanePair.empty(N, K, DEV, kind=kind); b.buf.normal_()
c = PlanePair.empty(M, N, DEV, kind=kind)
c2 = PlanePair.empty(M, N, DEV, kind=ops.PLANES_BF16x2)
bits = torch.empty(M, N // 32, dtype=torch.int32, device=DEV)
variants = {"full": {}, "nowait": {"AVR_UMMA_DEBUG": "128"}, "nostore": {"AVR_UMMA_DEBUG": str(256 + 128)},
            "nofence": {"AVR_UMMA_DEBUG": "1024"}, "nostage": {"AVR_UMMA_DEBUG": "512"},
            "nostage_nostore": {"AVR_UMMA_DEBUG": str(512 + 256 + 128 + 1024)}, "one_warp_per_group": {"AVR_UMMA_EPI_SPLIT": "1"},
            "one_buf": {"AVR_UMMA_EPI_BUFS": "1"}}
for dual in (0, 1):
    if dual:
        fn = lambda: ops.umma_nt(a, b, ops.UMMA_RELU | ops.UMMA_DUAL_COPY, c, c2=c2, bits_out=bits)
    else:
        fn = lambda: ops.umma_nt(a, b, ops.UMMA_RELU, c, bits_out=bits)
    best = {}
    for rnd in range(3):
        for name, env in variants.items():
            for k in ("AVR_UMMA_EPI_BUFS", "AVR_UMMA_DEBUG", "AVR_UMMA_EPI_SPLIT"): os.environ.pop(k, None)
            os.environ.update(env)
            best[name] = min(best.get(name, 1e9), timeit(fn))
    print(json.dumps({"dual_copy": dual, **{k: round(v, 4) for k, v in best.items()}}), flush=True)
