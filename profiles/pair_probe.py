"""CTA pairs (cta_group::2) and double-buffered epilogue staging of umma_gemm.cu: timing and bit-equality against the
single-CTA schedule.  Experiment build only (AVR_UMMA_CLUSTER = 0 off / 2 on for every eligible launch; AVR_UMMA_EPI_BUFS = 1
one staging buffer per epilogue warp).  Variants are timed in interleaved rounds; the minimum over the rounds is reported
(the boxes are power-capped and drift by +-10 % within seconds)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avr_b200 import ops
from avr_b200.ops import PlanePair
DEV = "cuda:0"
M, N = 524800, 512


def timeit(fn, n=6):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator(device=DEV).manual_seed(0)
for K, kind, okind in ((512, ops.PLANES_F16x2, ops.PLANES_F16x2), (208, ops.PLANES_BF16x3, ops.PLANES_F16x2),
                       (208, ops.PLANES_F16x2, ops.PLANES_F16x2), (512, ops.PLANES_BF16x2, ops.PLANES_BF16x2)):
    A = torch.randn(M, K, device=DEV, generator=g).clamp_min(0)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    a = ops.planes_split(A, PlanePair.empty(M, K, DEV, kind=kind))
    b = ops.planes_split(W, PlanePair.empty(N, K, DEV, kind=kind))
    del A, W
    c = PlanePair.empty(M, N, DEV, kind=okind)
    c2 = PlanePair.empty(M, N, DEV, kind=ops.PLANES_BF16x2)
    bits = torch.empty(M, N // 32, dtype=torch.int32, device=DEV)
    mask = torch.randint(-2**31, 2**31 - 1, (M, N // 32), dtype=torch.int32, device=DEV)
    variants = {"single": {"AVR_UMMA_CLUSTER": "0", "AVR_UMMA_EPI_BUFS": "1"}, "single_2buf": {"AVR_UMMA_CLUSTER": "0"},
                "pair_1buf": {"AVR_UMMA_CLUSTER": "2", "AVR_UMMA_EPI_BUFS": "1"}, "pair": {"AVR_UMMA_CLUSTER": "2"},
                "pair_noepi": {"AVR_UMMA_CLUSTER": "2", "AVR_UMMA_DEBUG": str(512 + 256 + 128 + 1024)}}
    for dual in (0, 1, 2):
        if dual == 2 and kind != ops.PLANES_BF16x2: continue
        if dual == 1:
            fn = lambda: ops.umma_nt(a, b, ops.UMMA_RELU | ops.UMMA_DUAL_COPY, c, c2=c2, bits_out=bits)
        elif dual == 2:
            fn = lambda: ops.umma_nt(a, b, ops.UMMA_MASK, c, mask=mask)          # backward-data shape of the step
        else:
            fn = lambda: ops.umma_nt(a, b, ops.UMMA_RELU, c, bits_out=bits)
        row = {"K": K, "kind": kind, "mode": ["relu+bits", "dual_copy", "masked"][dual]}
        best, ref = {}, None
        for rnd in range(3):
            for name, env in variants.items():
                for k in ("AVR_UMMA_CLUSTER", "AVR_UMMA_EPI_BUFS", "AVR_UMMA_DEBUG"): os.environ.pop(k, None)
                os.environ.update(env)
                if rnd == 0 and "DEBUG" not in "".join(env):
                    c.buf.zero_(); c2.buf.zero_(); bits.zero_()
                    fn()
                    out = (c.buf.clone(), c2.buf.clone() if dual == 1 else None, bits.clone())
                    if ref is None: ref = out
                    else: row["equal_" + name] = bool(torch.equal(out[0], ref[0]) and torch.equal(out[2], ref[2]) and
                                                      (dual != 1 or torch.equal(out[1], ref[1])))
                    del out
                t = timeit(fn)
                best[name] = min(best.get(name, 1e9), t)
        row.update({k: round(v, 4) for k, v in best.items()})
        print(json.dumps(row), flush=True)
        del ref
    del a, b, c, c2, bits, mask
