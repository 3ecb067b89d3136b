"""tcgen05 / TMEM / TMA dense layers (3 x bf16 error-compensated products) against float64."""
import pytest
import torch

from avr_b200 import ops
from avr_b200.ops import PlanePair
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 2e-5        # ~2^-17 per product, random signs


def _pp(x, ld=None, n=2):
    out = PlanePair.empty(x.shape[0], x.shape[1], DEV, ld, n)
    return ops.planes_split(x.to(DEV).contiguous(), out)


def test_planes_roundtrip(built_library):
    x = torch.randn(300, 72, generator=torch.Generator().manual_seed(0)) * 3
    pp = _pp(x, ld=80)
    back = ops.planes_merge(pp).cpu()
    assert float((back - x).abs().max() / x.abs().max()) < 2 ** -16
    t = PlanePair.empty(72, 300, DEV, 304)
    ops.planes_split(x.to(DEV), t, transpose=True, relu=True)
    assert rel_l2(ops.planes_merge(t), x.clamp_min(0).t()) < 1e-5


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1000, 128, 48), (4099, 512, 512), (300, 16, 128), (777, 208, 128),
                                   (513, 48, 128), (2050, 1600, 512), (256, 80, 208), (130, 512, 208)])
def test_umma_nt_plain(built_library, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    a, b = _pp(A), _pp(B)
    c = PlanePair.empty(M, N, DEV)
    ops.umma_nt(a, b, 0, c)
    ref = A.double() @ B.double().t()
    assert rel_l2(ops.planes_merge(c), ref) < TOL
    c32 = torch.empty(M, N, device=DEV)
    ops.umma_nt(a, b, ops.UMMA_OUT_F32, c_f32=c32)
    assert rel_l2(c32, ref) < TOL


def _pack_bits(pos):
    """bool [M, N] -> int32 bitmask [M, words] (bit j%32 of word j//32), as the forward epilogue writes it."""
    M, N = pos.shape
    words = ((N + 31) // 32 + 3) // 4 * 4
    padded = torch.zeros(M, words * 32, dtype=torch.int64)
    padded[:, :N] = pos.long()
    w = (padded.view(M, words, 32) << torch.arange(32)).sum(-1)
    return (w - ((w >> 31) << 32)).to(torch.int32).to(DEV)


def test_umma_nt_epilogues(built_library):
    g = torch.Generator().manual_seed(5)
    M, N, K = 1500, 256, 128
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    act, old = torch.randn(M, N, generator=g), torch.randn(M, N, generator=g)
    a, b, mask = _pp(A), _pp(B), _pack_bits(act > 0)
    ref = A.double() @ B.double().t()
    c = PlanePair.empty(M, N, DEV)
    bits = ops.relu_bits_empty(M, N, DEV)
    ops.umma_nt(a, b, ops.UMMA_RELU, c, bits_out=bits)
    assert rel_l2(ops.planes_merge(c), ref.clamp_min(0)) < TOL
    assert torch.equal(bits[:, : N // 32].cpu(), _pack_bits(ops.planes_merge(c).cpu() > 0)[:, : N // 32].cpu())
    c, c2 = PlanePair.empty(M, N, DEV), PlanePair.empty(M, N, DEV)
    ops.umma_nt(a, b, ops.UMMA_DUAL_RELU, c, c2)
    assert rel_l2(ops.planes_merge(c), ref) < TOL and rel_l2(ops.planes_merge(c2), ref.clamp_min(0)) < TOL
    c = _pp(old)
    ops.umma_nt(a, b, ops.UMMA_MASK | ops.UMMA_ACCUM, c, mask=mask)
    assert rel_l2(ops.planes_merge(c), ref * (act > 0) + old.double()) < TOL
    # column windows of wider buffers (how concatenated network inputs are addressed)
    wide = PlanePair.empty(M, 416, DEV)
    wide.buf.zero_()
    ops.umma_nt(a, b, 0, wide.window(128, N))
    m = ops.planes_merge(wide)
    assert rel_l2(m[:, 128:128 + N], ref) < TOL and float(m[:, :128].abs().max()) == 0 and float(m[:, 384:].abs().max()) == 0


@pytest.mark.parametrize("M,N,K", [(1000, 128, 48), (4099, 512, 512), (300, 16, 128), (777, 208, 128), (130, 512, 208)])
def test_umma_nt_six_products_fp32_grade(built_library, M, N, K):
    """3 planes x 3 planes (forward pass): operands carry 24 bits; what remains is the tensor core's own
    accumulation rounding (one truncation per tcgen05.mma into TMEM: ~7e-9 * K relative, measured)."""
    g = torch.Generator().manual_seed(M * N + K)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    a, b = _pp(A, n=3), _pp(B, n=3)
    assert float((ops.planes_merge(a).cpu() - A).abs().max()) <= 2 ** -23 * float(A.abs().max())
    ref = A.double() @ B.double().t()
    c = PlanePair.empty(M, N, DEV, n=3)
    ops.umma_nt(a, b, ops.UMMA_RELU, c)
    bound = 1e-8 * K + 5e-7
    err = rel_l2(ops.planes_merge(c), ref.clamp_min(0))
    assert err < bound, (err, bound)
    c2 = PlanePair.empty(M, N, DEV, n=2)
    ops.umma_nt(_pp(A), _pp(B), ops.UMMA_RELU, c2)
    assert err < rel_l2(ops.planes_merge(c2), ref.clamp_min(0))          # strictly better than the 3-product mode
    c32 = torch.empty(M, N, device=DEV)
    ops.umma_nt(a, b, ops.UMMA_OUT_F32, c_f32=c32)
    assert rel_l2(c32, ref) < bound


@pytest.mark.parametrize("M,N,K", [(256, 2400, 2402), (256, 1608, 1600), (100, 808, 640)])
def test_umma_nt_split_k_long_reduction(built_library, M, N, K):
    """The DFT / adjoint-DFT shapes: one accumulator chain of K/16*6 MMAs is truncated once per MMA (a bias of
    ~6e-9*K); the split-K mode sums 256-deep slices in fp32 instead and must be markedly closer to float64."""
    g = torch.Generator().manual_seed(K)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    a, b = _pp(A, n=3), _pp(B, n=3)
    ref = ops.planes_merge(a).double() @ ops.planes_merge(b).double().t()
    c0, c1, c2 = (torch.empty(M, N, device=DEV) for _ in range(3))
    ops.umma_nt(a, b, ops.UMMA_OUT_F32, c_f32=c0)
    ops.umma_nt(a, b, ops.UMMA_OUT_F32, c_f32=c1, split_k=True)
    ops.umma_nt(a, b, ops.UMMA_OUT_F32, c_f32=c2, split_k=True)
    e0, e1 = rel_l2(c0, ref), rel_l2(c1, ref)
    assert e1 < 2.5e-6 and e1 < 0.4 * e0, (e0, e1)
    assert torch.equal(c1, c2)


def _pp16(x, ld=None):
    out = PlanePair.empty(x.shape[0], x.shape[1], DEV, ld, kind=ops.PLANES_F16x2)
    return ops.planes_split(x.to(DEV).contiguous(), out)


def test_f16_pair_roundtrip(built_library):
    """x = hi + lo' * 2^-11: 24 bits in two fp16 planes inside fp16's normal range; absolute error <= 1.5e-11 below it."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(500, 72, generator=g) * 10 ** (torch.rand(500, 72, generator=g) * 6 - 3)     # 1e-3 .. 1e3
    pp = _pp16(x, ld=80)
    assert pp.kind == ops.PLANES_F16x2 and pp.buf.dtype == torch.float16
    back = ops.planes_merge(pp).cpu()
    assert bool(((back - x).abs() <= 2 ** -22 * x.abs() + 3e-11).all())
    tiny = torch.randn(64, 8, generator=g) * 1e-6
    assert float((ops.planes_merge(_pp16(tiny)).cpu() - tiny).abs().max()) <= 3e-11
    t = PlanePair.empty(72, 500, DEV, 504, kind=ops.PLANES_F16x2)
    ops.planes_split(x.to(DEV), t, transpose=True, relu=True)
    assert rel_l2(ops.planes_merge(t), x.clamp_min(0).t()) < 2e-7


@pytest.mark.parametrize("M,N,K", [(1000, 128, 48), (4099, 512, 512), (300, 16, 128), (777, 208, 128), (130, 512, 208),
                                   (2050, 1600, 512)])
def test_umma_nt_f16_pairs_fp32_grade(built_library, M, N, K):
    """fp16 (hi, lo') pairs: three products (hi*hi; hi*lo' + lo'*hi through the scaled second accumulator) are as
    accurate as the six products of bf16 triples -- and as an fp32 GEMM."""
    g = torch.Generator().manual_seed(M * N + K)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    a, b = _pp16(A), _pp16(B)
    ref = ops.planes_merge(a).double() @ ops.planes_merge(b).double().t()
    assert rel_l2(ops.planes_merge(a), A) < 1e-7
    c = PlanePair.empty(M, N, DEV, kind=ops.PLANES_F16x2)
    bits = ops.relu_bits_empty(M, N, DEV)
    ops.umma_nt(a, b, ops.UMMA_RELU, c, bits_out=bits)
    got = ops.planes_merge(c)
    bound = 2e-9 * K + 4e-7
    assert rel_l2(got, ref.clamp_min(0)) < bound, (rel_l2(got, ref.clamp_min(0)), bound)
    assert torch.equal(_unpack_bits(bits, N), got > 0)
    fp32 = (A.to(DEV) @ B.to(DEV).t())
    assert rel_l2(got, ref.clamp_min(0)) < 3 * rel_l2(fp32.clamp_min(0), (A.double() @ B.double().t()).clamp_min(0)) + 1e-7
    c32 = torch.empty(M, N, device=DEV)
    ops.umma_nt(a, b, ops.UMMA_OUT_F32, c_f32=c32)
    assert rel_l2(c32, ref) < bound
    if N % 8 == 0 and N >= 128:
        c1, c2 = PlanePair.empty(M, N, DEV, kind=ops.PLANES_F16x2), PlanePair.empty(M, N, DEV, kind=ops.PLANES_F16x2)
        ops.umma_nt(a, b, ops.UMMA_DUAL_RELU, c1, c2)                        # raw + ReLU outputs
        assert rel_l2(ops.planes_merge(c1), ref) < bound and torch.equal(ops.planes_merge(c2), got)
        # the same activation twice: fp16 pair for the next forward layer, bf16 (hi, mid) pair for the weight gradient
        c3, c4 = PlanePair.empty(M, N, DEV, kind=ops.PLANES_F16x2), PlanePair.empty(M, N, DEV, ld=N + 8)
        ops.umma_nt(a, b, ops.UMMA_RELU | ops.UMMA_DUAL_COPY, c3, c4, bits_out=bits)
        assert torch.equal(ops.planes_merge(c3), got) and rel_l2(ops.planes_merge(c4), got) < 2 ** -16
        # bf16 triples in, fp16 pair + bf16 pair out (first hidden layer of the signal network)
        a3, b3 = _pp(A, n=3), _pp(B, n=3)
        ops.umma_nt(a3, b3, ops.UMMA_RELU | ops.UMMA_DUAL_COPY, c3, c4)
        ref3 = (ops.planes_merge(a3).double() @ ops.planes_merge(b3).double().t()).clamp_min(0)
        assert rel_l2(ops.planes_merge(c3), ref3) < bound and rel_l2(ops.planes_merge(c4), ref3) < 2 ** -16


def _unpack_bits(bits, N):
    sh = torch.arange(32, device=bits.device, dtype=torch.int32)
    return ((bits[:, :(N + 31) // 32].unsqueeze(-1) >> sh) & 1).reshape(bits.shape[0], -1)[:, :N].bool()


@pytest.mark.parametrize("M,N,K,mode", [(8192, 512, 512, "relu"), (8192, 128, 128, "dual"), (5000, 512, 208, "bias"),
                                         (40000, 16, 128, "f32")])
def test_umma_near_zero_guard(built_library, M, N, K, mode):
    """Forward layers: elements too close to zero for the tensor core's accumulation error (one truncation per MMA,
    ~6e-9*K of the row scale, biased) are listed by the epilogue and re-evaluated with fp32 FMAs, so that sign decisions
    (ReLU masks, the |leaky_relu| kink of the density head) are fp32-grade.  The listed set is a function of the values
    only: outputs and counts are bit-identical run to run."""
    g = torch.Generator().manual_seed(M + N + K)
    A, B = torch.randn(M, K, generator=g).clamp_min(0), torch.randn(N, K, generator=g) / K ** 0.5
    a, b = _pp(A, n=3), _pp(B, n=3)
    ref = ops.planes_merge(a).double() @ ops.planes_merge(b).double().t()
    kw, geom = {}, None
    if mode == "bias":
        from avr_b200.configs import tiny_config
        geom = ops.make_geom(tiny_config("AVRModel", n_azi=6, n_ele=4, n_samples=10)["render"], M // (26 * 10), 400)
        M = geom.bs * geom.R * geom.S
        a, ref = a.row_window(0, M), ref[:M]
        t_rcv = torch.randn(geom.bs, N, generator=g).to(DEV)
        ref = ref + t_rcv.double().repeat_interleave(geom.R * geom.S, 0)
        kw = dict(bias_rcv=t_rcv, geom=geom)
    scale = ref.abs().mean(1, keepdim=True)
    tau = max(ops.NearZeroGuard.TAU_MIN, ops.NearZeroGuard.TAU_PER_K * K)

    def run(guard):
        bits = ops.relu_bits_empty(M, N, DEV).zero_()
        c, c2 = PlanePair.zeros(M, N, DEV, n=3), PlanePair.zeros(M, N, DEV, n=3)
        c32 = torch.zeros(M, N, device=DEV)
        if mode == "f32":
            ops.umma_nt(a, b, ops.UMMA_OUT_F32, c_f32=c32, guard=guard)
            return c32, None, None
        if mode == "dual":
            ops.umma_nt(a, b, ops.UMMA_DUAL_RELU, c, c2, bits_out=bits, guard=guard)
            return ops.planes_merge(c), ops.planes_merge(c2), _unpack_bits(bits, N)
        ops.umma_nt(a, b, ops.UMMA_RELU, c, bits_out=bits, guard=guard, **kw)
        return None, ops.planes_merge(c), _unpack_bits(bits, N)

    guard = ops.NearZeroGuard(DEV)
    raw, pos, bits = run(guard)
    n_listed = guard.counts()[0]
    listed = guard.list[:n_listed].long()
    assert 0 < n_listed < 1e-3 * M * N
    near = torch.zeros(M, N, dtype=torch.bool, device=DEV)
    near[listed[:, 0], listed[:, 1]] = True
    # everything well inside the threshold is listed, nothing well outside (the scale is a 32-column estimate)
    assert bool(near[ref.abs() < 0.3 * tau * scale].all()) and not bool(near[ref.abs() > 3 * tau * scale].any())
    # listed elements carry fp32-FMA values: error ~1e-7 of the row scale instead of ~6e-9*K
    for val, want in ((raw, ref), (pos, ref.clamp_min(0))):
        if val is not None:
            assert float(((val.double() - want).abs() / scale)[near].max()) < 6e-7
    if bits is not None:
        assert torch.equal(bits, pos > 0)
        assert int((bits != (ref > 0)).sum()) <= 2                 # fp32-grade decisions (float64 disagrees ~1e-7 of the time)
    # unguarded: same values elsewhere; guarded runs are reproducible
    raw0, pos0, bits0 = run(None)
    for x, x0 in ((raw, raw0), (pos, pos0)):
        if x is not None:
            assert torch.equal(x[~near], x0[~near])
    guard2 = ops.NearZeroGuard(DEV)
    raw2, pos2, bits2 = run(guard2)
    assert guard2.counts()[0] == n_listed
    assert all(x is None or torch.equal(x, y) for x, y in ((raw, raw2), (pos, pos2), (bits, bits2)))


@pytest.mark.parametrize("M,N,K", [(128, 128, 4096), (512, 512, 20000), (16, 128, 5000), (128, 48, 3001), (512, 208, 7777),
                                   (1600, 512, 4100), (128, 80, 64), (512, 464, 9000), (256, 160, 4200)])   # 160-column tiles: halves of 80
def test_umma_tn_weight_grad(built_library, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    dY, X = torch.randn(K, M, generator=g), torch.randn(K, N, generator=g)
    a, b = _pp(dY), _pp(X)
    ws = torch.empty(max(4, ops.umma_tn_workspace_bytes(M, N, K) // 4), device=DEV)
    outs = []
    for _ in range(2):
        c = torch.zeros(M, N + 8, device=DEV)
        ops.umma_tn(a, b, c[:, :N], ws)
        outs.append(c.cpu())
    assert torch.equal(outs[0], outs[1])                     # deterministic split-K
    assert rel_l2(outs[0][:, :N], dY.double().t() @ X.double()) < TOL
    assert float(outs[0][:, N:].abs().max()) == 0


@pytest.mark.parametrize("M,N,K", [(512, 512, 40000), (512, 208, 9001), (128, 80, 3000), (512, 512, 128 * 148 * 3 + 77),
                                   (512, 464, 8320), (256, 160, 5000)])
def test_umma_tn_fp16_pair_activation_converted_in_kernel(built_library, M, N, K):
    """Weight gradient whose activation operand is an fp16 (hi, lo') pair: the kernel's epilogue warps turn every tile into a
    bf16 (hi, mid) pair in shared memory (umma_gemm.cu, conv_b).  Bit-identical to handing it the bf16 pair that
    planes_split writes for the same values."""
    g = torch.Generator().manual_seed(M + N + K)
    dY = torch.randn(K, M, generator=g) * torch.rand(K, 1, generator=g) ** 8          # heavy-tailed rows, like real gradients
    X = torch.randn(K, N, generator=g).clamp_min(0) * 3
    a = _pp(dY)
    x16 = ops.planes_split(X.to(DEV), PlanePair.empty(K, N, DEV, kind=ops.PLANES_F16x2))
    xbf = ops.planes_split(ops.planes_merge(x16), PlanePair.empty(K, N, DEV, kind=ops.PLANES_BF16x2))
    ws = torch.empty(max(4, ops.umma_tn_workspace_bytes(M, N, K) // 4), device=DEV)
    c16, cbf = torch.zeros(M, N, device=DEV), torch.zeros(M, N, device=DEV)
    ops.umma_tn(a, x16, c16, ws)
    ops.umma_tn(a, xbf, cbf, ws)
    assert torch.equal(c16, cbf)
    assert torch.equal(ops.planes_merge(x16), ops.planes_merge(ops.planes_split(X.to(DEV), PlanePair.empty(K, N, DEV, kind=ops.PLANES_F16x2))))   # operand untouched
    assert rel_l2(c16, ops.planes_merge(a).double().t() @ ops.planes_merge(x16).double()) < TOL


@pytest.mark.parametrize("M,N,K", [(16, 128, 5000), (128, 128, 20000), (128, 48, 3001), (512, 208, 2000), (128, 256, 9000)])
def test_umma_tn_six_products_ill_conditioned(built_library, M, N, K):
    """Weight gradients of the density path: the terms of sum_k dY[k,m] X[k,n] cancel to a few percent of their
    magnitude, so the 16-bit (three-product) mode is ~30x too coarse there; 24-bit operands / six products are not."""
    g = torch.Generator().manual_seed(M + N + K)
    X = 1 + 0.5 * torch.randn(K, N, generator=g)
    dY = torch.randn(K, M, generator=g)
    dY -= dY.mean(0, keepdim=True)                                              # sums against X cancel ~ sqrt(K)-fold
    a3, b3 = _pp(dY, n=3), _pp(X, n=3)
    ref = ops.planes_merge(a3).double().t() @ ops.planes_merge(b3).double()
    ws = torch.empty(max(4, ops.umma_tn_workspace_bytes(M, N, K) // 4), device=DEV)
    c3, c3b, c2 = (torch.zeros(M, N, device=DEV) for _ in range(3))
    ops.umma_tn(a3, b3, c3, ws)
    ops.umma_tn(a3, b3, c3b, ws)
    ops.umma_tn(_pp(dY), _pp(X), c2, ws)
    e3, e2 = rel_l2(c3, ref), rel_l2(c2, ref)
    assert torch.equal(c3, c3b)
    assert e3 < 1e-5 and e3 < 0.2 * e2, (e3, e2)


@pytest.mark.parametrize("bs,R,S,T,W", [(2, 37, 5, 200, 64), (1, 300, 7, 400, 512), (3, 66, 4, 240, 136)])
@pytest.mark.parametrize("act_kind", [ops.PLANES_BF16x2, ops.PLANES_F16x2])
def test_collapse_matches_literal_output_layer(built_library, bs, R, S, T, W, act_kind):
    """y, d_H, d_w, d_W_out of the collapsed output layer == literal GEMM + masked ray reduction (float64)."""
    g = torch.Generator().manual_seed(R + T)
    n = bs * R * S
    H = torch.randn(n, W, generator=g).clamp_min(0)                      # post-ReLU hidden activation
    Wout = torch.randn(T, W, generator=g) / W ** 0.5
    w = torch.rand(bs, R, S, generator=g)
    delay = torch.randint(0, T // 3, (bs, R, S), generator=g, dtype=torch.int32) + torch.arange(S, dtype=torch.int32) * 11
    delay[0, :, 0] = 7                                                    # one (b,s) where every ray shares a delay
    dy = torch.randn(bs, S, T, generator=g)
    geom = ops.RenderGeom(bs, R, S, T, -10.0, 20.0, 16000.0, 343.8)
    hp = PlanePair.empty(n, W, DEV, kind=act_kind)
    ops.planes_split(H.to(DEV), hp)
    Hq = ops.planes_merge(hp).cpu().double()                              # the values the kernels actually see
    Hq.requires_grad_()
    Wd = Wout.double().requires_grad_()
    wd = w.double().requires_grad_()
    sig = (Hq @ Wd.t()).view(bs, R, S, T)
    mask = (torch.arange(T)[None, None, None, :] >= delay[..., None]).double()
    y_ref = (sig * mask * wd[..., None]).sum(1)
    (y_ref * dy.double()).sum().backward()
    sort = ops.delay_sort(geom, delay.to(DEV), w.to(DEV))
    sd = sort[1].cpu()
    assert bool((sd[..., 1:] >= sd[..., :-1]).all())                      # sortedness
    assert torch.equal(torch.sort(sort[0].cpu(), dim=-1).values, torch.arange(R, dtype=torch.int32).expand(bs, S, R))
    y, prefix = ops.collapse_fwd(geom, hp, sort, Wout.to(DEV), tspan=T)
    assert rel_l2(y, y_ref) < 2e-6
    dH = PlanePair.empty(n, W, DEV)
    dW = torch.zeros(T, W + 8, device=DEV)
    d_w = ops.collapse_bwd(geom, hp, sort, Wout.to(DEV), dy.to(DEV), T, prefix, dH, dW[:, :W])
    assert rel_l2(d_w, wd.grad) < 2e-6
    assert rel_l2(ops.planes_merge(dH), Hq.grad * (Hq.detach() > 0)) < 2e-5
    assert rel_l2(dW[:, :W], Wd.grad) < 2e-6 and float(dW[:, W:].abs().max()) == 0
    y2, prefix2 = ops.collapse_fwd(geom, hp, sort, Wout.to(DEV), tspan=3)     # violated bound must poison, not corrupt
    assert bool(torch.isnan(y2).any())


def test_umma_bias_epilogue_and_block_sum(built_library):
    """Per-ray / per-receiver rows added in the GEMM epilogue and the adjoint block sum."""
    g = torch.Generator().manual_seed(9)
    bs, R, S, K, N = 3, 10, 64, 128, 256                       # S = 64: warps share their ray; also a ragged case below
    for S_ in (S, 24):
        M = bs * R * S_
        geom = ops.RenderGeom(bs, R, S_, 200, -10.0, 20.0, 16000.0, 343.8)
        A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
        t_ray, t_rcv = torch.randn(R, N, generator=g), torch.randn(bs, N, generator=g)
        rows = torch.arange(M)
        ref = A.double() @ B.double().t() + t_ray[(rows // S_) % R].double() + t_rcv[rows // (R * S_)].double()
        c = PlanePair.empty(M, N, DEV, n=3)
        ops.umma_nt(_pp(A, n=3), _pp(B, n=3), ops.UMMA_RELU, c, bias_ray=t_ray.to(DEV), bias_rcv=t_rcv.to(DEV), geom=geom)
        assert rel_l2(ops.planes_merge(c), ref.clamp_min(0)) < 2e-6
        x = torch.randn(M, 72, generator=g)
        part = ops.rows_block_sum(geom, x[:, :64].contiguous().to(DEV))
        assert rel_l2(part, x[:, :64].double().view(bs * R, S_, 64).sum(1)) < 1e-6
        part_p = ops.rows_block_sum(geom, _pp(x[:, :64].contiguous()))
        assert rel_l2(part_p, x[:, :64].double().view(bs * R, S_, 64).sum(1)) < 1e-5


@pytest.mark.parametrize("kind,okind,N,K,mode", [
    (ops.PLANES_F16x2, ops.PLANES_F16x2, 512, 512, "relu_bits"),          # fp16-pair hidden layer
    (ops.PLANES_F16x2, ops.PLANES_F16x2, 256, 208, "dual_copy"),          # ... with the bf16 copy for the weight gradient
    (ops.PLANES_BF16x3, ops.PLANES_F16x2, 512, 208, "dual_copy"),         # first signal layer (six products)
    (ops.PLANES_BF16x2, ops.PLANES_BF16x2, 512, 512, "masked"),           # backward-data product
    (ops.PLANES_BF16x2, ops.PLANES_BF16x2, 208, 512, "plain"),            # a single column tile whose B is not resident
])
def test_umma_nt_cta_pairs_equal_single_cta(built_library, kind, okind, N, K, mode):
    """Products over >= 2 * 128 * (number of SMs) rows run as CTA pairs (tcgen05 cta_group::2, 2x1x1 clusters; umma_gemm.cu,
    pair mode): an ODD number of row tiles with a ragged last one here.  The same product computed in row slices too short
    for pair mode takes the single-CTA schedule -- results must agree bit for bit (and with float64 to the kind's accuracy)."""
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    M = 2 * 128 * sms + 128 * 3 + 37
    g = torch.Generator(device=DEV).manual_seed(N + K)
    A = torch.randn(M, K, device=DEV, generator=g).clamp_min(-0.5)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    a = ops.planes_split(A, PlanePair.empty(M, K, DEV, kind=kind))
    b = ops.planes_split(W, PlanePair.empty(N, K, DEV, kind=kind))
    words = ops.relu_bits_empty(1, N, DEV).shape[1]
    mask = torch.randint(-2 ** 31, 2 ** 31 - 1, (M, words), dtype=torch.int32, device=DEV) if mode == "masked" else None

    def run(a_rows, c, c2, bits, mask_rows):
        if mode == "relu_bits":
            ops.umma_nt(a_rows, b, ops.UMMA_RELU, c, bits_out=bits)
        elif mode == "dual_copy":
            ops.umma_nt(a_rows, b, ops.UMMA_RELU | ops.UMMA_DUAL_COPY, c, c2=c2, bits_out=bits)
        elif mode == "masked":
            ops.umma_nt(a_rows, b, ops.UMMA_MASK, c, mask=mask_rows)
        else:
            ops.umma_nt(a_rows, b, 0, c)

    c = PlanePair.empty(M, N, DEV, kind=okind)
    c2 = PlanePair.empty(M, N, DEV, kind=ops.PLANES_BF16x2)
    bits = ops.relu_bits_empty(M, N, DEV).zero_()
    c.buf.zero_(); c2.buf.zero_()
    run(a, c, c2, bits, mask)                                            # pair schedule
    s = PlanePair.empty(M, N, DEV, kind=okind)
    s2 = PlanePair.empty(M, N, DEV, kind=ops.PLANES_BF16x2)
    sbits = ops.relu_bits_empty(M, N, DEV).zero_()
    s.buf.zero_(); s2.buf.zero_()
    step = 16384                                                         # < 2 * 128 * SMs rows: single-CTA schedule
    for r0 in range(0, M, step):
        n = min(step, M - r0)
        run(a.row_window(r0, n), s.row_window(r0, n), s2.row_window(r0, n), sbits[r0:r0 + n],
            mask[r0:r0 + n] if mask is not None else None)
    assert torch.equal(c.buf, s.buf)
    if mode == "dual_copy":
        assert torch.equal(c2.buf, s2.buf)
    if mode in ("relu_bits", "dual_copy"):
        assert torch.equal(bits, sbits)
    ref = ops.planes_merge(a).double() @ ops.planes_merge(b).double().t()
    if mode in ("relu_bits", "dual_copy"):
        ref = ref.clamp_min(0)
    if mode == "masked":
        ref = ref * _unpack_bits(mask, N).to(ref.dtype)
    bound = TOL if kind == ops.PLANES_BF16x2 else 2e-6
    assert rel_l2(ops.planes_merge(c), ref) < bound
