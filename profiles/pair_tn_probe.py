"""Weight-gradient (MN-major) GEMMs as CTA pairs: timing and bit-equality against the single-CTA schedule.  Experiment build
(AVR_UMMA_CLUSTER=0 switches pair mode off)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avr_b200 import ops
from avr_b200.ops import PlanePair
DEV = "cuda:0"
K = 524800


def timeit(fn, n=6):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator(device=DEV).manual_seed(0)
for M, N, n_planes in ((512, 512, 2), (512, 208, 2), (256, 128, 3), (512, 512, 3)):
    dY = torch.randn(K, M, device=DEV, generator=g)
    X = torch.randn(K, N, device=DEV, generator=g).clamp_min(0)
    a = ops.planes_split(dY, PlanePair.empty(K, M, DEV, n=n_planes))
    b = ops.planes_split(X, PlanePair.empty(K, N, DEV, n=n_planes))
    del dY, X
    ws = torch.empty(max(4, ops.umma_tn_workspace_bytes(M, N, K) // 4), device=DEV)
    c = {}
    best = {}
    for rnd in range(3):
        for name, env in (("single", "0"), ("pair", None)):
            if env is None: os.environ.pop("AVR_UMMA_CLUSTER", None)
            else: os.environ["AVR_UMMA_CLUSTER"] = env
            out = torch.zeros(M, N, device=DEV)
            fn = lambda: ops.umma_tn(a, b, out, ws)
            best[name] = min(best.get(name, 1e9), timeit(fn))
            c[name] = out.clone()
    print(json.dumps({"M": M, "N": N, "planes": n_planes, "single_ms": round(best["single"], 4), "pair_ms": round(best["pair"], 4),
                      "bit_equal": bool(torch.equal(c["single"], c["pair"]))}), flush=True)
    del a, b, ws
