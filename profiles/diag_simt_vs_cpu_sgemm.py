"""Is the library's fp32 SIMT GEMM the SAME floating-point function as the CPU oracle's sgemm (torch CPU matmul) for the layer
shapes of the sigma networks?  Counts bit-identical outputs; also a sequential-FMA chain and a permuted-order evaluation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avr_b200 import ops
DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
for (M, N, K) in ((8192, 128, 48), (8192, 128, 128), (8192, 16, 128), (8192, 512, 208), (8192, 512, 512)):
    x = torch.randn(M, K, generator=g).clamp_min(0) * 0.3
    w = torch.randn(N, K, generator=g) / K ** 0.5
    y_cpu = x @ w.t()                                              # what oracle/field_ref.py does
    y = torch.empty(M, N, device=DEV)
    ops.linear_fwd(x.to(DEV), w.to(DEV), y)
    y = y.cpu()
    seq = torch.zeros(M, N)                                         # one sequential fp32 FMA chain over k (float64 fma emulation is
    acc = torch.zeros(M, N, dtype=torch.float64)                    # not exact; use the GPU's own addcmul in fp32 as the chain)
    xs, ws = x.to(DEV), w.to(DEV)
    chain = torch.zeros(M, N, device=DEV)
    for k in range(K):
        chain = torch.addcmul(chain, xs[:, k:k + 1], ws[:, k].unsqueeze(0))   # fused multiply-add in fp32 on the GPU
    chain = chain.cpu()
    perm = torch.randperm(K, generator=g)
    y_perm = x[:, perm] @ w[:, perm].t()
    ref = x.double() @ w.double().t()
    scale = ref.abs().mean()
    print(f"M{M} N{N} K{K}: simt == cpu sgemm on {float((y == y_cpu).float().mean()) * 100:.2f} % of the outputs, simt == sequential FMA chain "
          f"{float((y == chain).float().mean()) * 100:.2f} %, cpu sgemm == its own permuted-k evaluation {float((y_cpu == y_perm).float().mean()) * 100:.2f} %; "
          f"|err| / mean|y|: simt {float((y - ref).abs().mean() / scale):.1e}, cpu {float((y_cpu - ref).abs().mean() / scale):.1e}", flush=True)
