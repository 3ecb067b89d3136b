"""The fused chain of 128-wide dense layers (``avr_mlp_chain_fwd``, csrc/mlp_chain.cu) against the layer-by-layer
tensor-core kernel (``avr_umma_gemm_nt``): same six-product / two-accumulator arithmetic, so EVERY output -- saved planes,
raw planes, ReLU bitmasks, the fp32 head -- must be bit-identical; and against float64 for good measure."""
import pytest
import torch

from avr_b200 import ops
from avr_b200.ops import PlanePair
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
K3 = ops.PLANES_BF16x3


def _planes(x, kind=K3):
    return ops.planes_split(x.to(DEV), PlanePair.empty(x.shape[0], x.shape[1], DEV, kind=kind))


def _reference_chain(x0p, mats, spec, M):
    """Layer by layer through umma_nt, mirroring what fused_tc.py did before the chain kernel existed."""
    outs, h = [], x0p
    for w, s in zip(mats, spec):
        wp = _planes(w)
        o = {}
        if s.get("out_f32"):
            o["out_f32"] = torch.zeros(M, w.shape[0], device=DEV)
            ops.umma_nt(h, wp, ops.UMMA_OUT_F32, c_f32=o["out_f32"])
        elif s.get("raw"):
            o["save_raw"] = PlanePair.zeros(M, w.shape[0], DEV, kind=K3)
            o["save"] = PlanePair.zeros(M, w.shape[0], DEV, kind=K3)
            o["bits"] = torch.zeros_like(ops.relu_bits_empty(M, w.shape[0], DEV))
            ops.umma_nt(h, wp, ops.UMMA_DUAL_RELU, o["save_raw"], o["save"], bits_out=o["bits"])
            h = o["save"]
        else:
            y = PlanePair.zeros(M, w.shape[0], DEV, kind=K3)
            o["bits"] = torch.zeros_like(ops.relu_bits_empty(M, w.shape[0], DEV))
            ops.umma_nt(h, wp, ops.UMMA_RELU if s["relu"] else 0, y, bits_out=o["bits"])
            o["full"] = y
            h = y
        outs.append(o)
    return outs


@pytest.mark.parametrize("M,k0,n_hidden,head", [(300, 48, 4, 16), (128 * 149 + 37, 48, 4, 16), (1000, 128, 3, 16), (517, 80, 2, 64),
                                                (129, 16, 1, 128)])
def test_chain_is_bit_identical_to_the_layer_by_layer_kernel(built_library, M, k0, n_hidden, head):
    g = torch.Generator().manual_seed(M + k0)
    x0 = torch.randn(M, k0, generator=g)
    dims = [k0] + [128] * n_hidden + [head]
    mats = [torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5 for i in range(len(dims) - 1)]
    # hidden layers: ReLU + saved planes + bitmask; the LAST hidden layer also keeps its raw output (sigma_feat); the head
    # is fp32 when 16 wide, else a linear plane output
    spec = [{"relu": True} for _ in range(n_hidden)]
    spec[-1]["raw"] = True
    spec.append({"out_f32": True} if head == 16 else {"relu": False})
    x0p = _planes(x0)
    ref = _reference_chain(x0p, mats, spec, M)
    layers, mine = [], []
    for li, (w, s) in enumerate(zip(mats, spec)):
        L = {"w": _planes(w), "relu": bool(s.get("relu"))}
        o = {}
        if s.get("out_f32"):
            o["out_f32"] = torch.zeros(M, w.shape[0], device=DEV)
            L["out_f32"] = o["out_f32"]
        else:
            kind = ops.PLANES_BF16x2 if (li % 2 == 1 and not s.get("raw")) else K3       # 2-plane saves: the first two planes
            o["save"] = PlanePair.zeros(M, w.shape[0], DEV, kind=kind)
            o["bits"] = torch.zeros_like(ops.relu_bits_empty(M, w.shape[0], DEV))
            L["save"], L["bits"] = o["save"], o["bits"]
            if s.get("raw"):
                o["save_raw"] = PlanePair.zeros(M, w.shape[0], DEV, kind=K3)
                L["save_raw"] = o["save_raw"]
        layers.append(L)
        mine.append(o)
    ops.mlp_chain(x0p, layers)
    torch.cuda.synchronize()
    # float64 chain on the values the kernels see
    h = ops.planes_merge(x0p).double().cpu()
    for li, (w, s, a, b) in enumerate(zip(mats, spec, mine, ref)):
        wq = ops.planes_merge(_planes(w)).double().cpu()
        pre = h @ wq.t()
        if s.get("out_f32"):
            assert torch.equal(a["out_f32"], b["out_f32"]), li
            assert rel_l2(a["out_f32"], pre) < 2e-6
            continue
        words = (w.shape[0] + 31) // 32
        assert torch.equal(a["bits"][:, :words], b["bits"][:, :words]), li
        full = b["full"] if "full" in b else b["save"]
        n = a["save"].n
        assert torch.equal(a["save"].buf[:n], full.buf[:n]), li                       # the saved planes, bit for bit
        if s.get("raw"):
            assert torch.equal(a["save_raw"].buf, b["save_raw"].buf), li
            assert rel_l2(ops.planes_merge(a["save_raw"]), pre) < 2e-6
        h = ops.planes_merge(full).double().cpu()                                      # next layer sees the stored values
        assert rel_l2(h, pre.clamp_min(0) if s["relu"] else pre) < 2e-6


@pytest.mark.parametrize("bs,R,S", [(3, 37, 6), (2, 50, 64)])
def test_chain_with_per_receiver_bias_rows(built_library, bs, R, S):
    """Channel embeddings in 'add' mode (model.py:44-47): a per-receiver fp32 row is added to a hidden layer's output before
    the ReLU.  The chain (``bias`` / ``bias_group_rows``) against the layer-by-layer kernel (``bias_rcv`` + geometry): bit for
    bit; receivers change in the middle of 128-row tiles (R * S is not a multiple of 128)."""
    M = bs * R * S
    geom = ops.RenderGeom(bs, R, S, 200, -10.0, 20.0, 16000.0, 343.8)
    g = torch.Generator().manual_seed(M)
    x0 = torch.randn(M, 48, generator=g)
    dims = [48, 128, 128, 128, 16]
    mats = [torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5 for i in range(4)]
    biases = [torch.randn(bs, 128, generator=g).to(DEV), None, torch.randn(bs, 128, generator=g).to(DEV)]   # layers 0 and 2
    x0p = _planes(x0)
    wps = [_planes(w) for w in mats]
    # layer by layer
    h, ref = x0p, []
    for li in range(3):
        y = PlanePair.zeros(M, 128, DEV, kind=K3)
        bits = torch.zeros_like(ops.relu_bits_empty(M, 128, DEV))
        if biases[li] is not None:
            ops.umma_nt(h, wps[li], ops.UMMA_RELU, y, bits_out=bits, bias_rcv=biases[li], geom=geom)
        else:
            ops.umma_nt(h, wps[li], ops.UMMA_RELU, y, bits_out=bits)
        ref.append((y, bits))
        h = y
    head_ref = torch.zeros(M, 16, device=DEV)
    ops.umma_nt(h, wps[3], ops.UMMA_OUT_F32, c_f32=head_ref)
    # one launch
    layers, mine = [], []
    for li in range(3):
        y = PlanePair.zeros(M, 128, DEV, kind=K3)
        bits = torch.zeros_like(ops.relu_bits_empty(M, 128, DEV))
        L = dict(w=wps[li], relu=True, save=y, bits=bits)
        if biases[li] is not None:
            L.update(bias=biases[li], bias_group_rows=R * S)
        layers.append(L)
        mine.append((y, bits))
    head = torch.zeros(M, 16, device=DEV)
    layers.append(dict(w=wps[3], relu=False, out_f32=head))
    ops.mlp_chain(x0p, layers)
    for li, ((y, bits), (yr, br)) in enumerate(zip(mine, ref)):
        assert torch.equal(y.buf, yr.buf), li
        assert torch.equal(bits[:, :4], br[:, :4]), li
    assert torch.equal(head, head_ref)
    # and the bias really is the receiver's row: float64 of the first layer
    pre = ops.planes_merge(x0p).double().cpu() @ ops.planes_merge(wps[0]).double().cpu().t()
    pre = pre + biases[0].double().cpu().repeat_interleave(R * S, dim=0)
    assert rel_l2(ops.planes_merge(mine[0][0]).cpu(), pre.clamp_min(0)) < 2e-6


@pytest.mark.parametrize("M", [300, 128 * 150 + 5])
def test_backward_data_chain_is_bit_identical(built_library, M):
    """The backward-data pass of sigma decoder -> sigma encoder as one chain: ReLU bitmasks multiplied into every output,
    24-bit (six-product) layers along the decoder, the accumulate into d_feat (two consumers of sigma_feat), then 16-bit
    (three-product) layers along the encoder, a 48-wide last layer -- against umma_nt with UMMA_MASK / UMMA_ACCUM."""
    g = torch.Generator().manual_seed(M)
    K2 = ops.PLANES_BF16x2
    g0 = _planes(torch.randn(M, 16, generator=g))                                     # d(density head), 3 planes
    dims = [16, 128, 128, 128, 128, 48]
    kinds = [K3, K3, K3, K2, K2]                                                       # transposed weights: decoder 24 bits, encoder 16
    wts = [_planes(torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5, kinds[i]) for i in range(5)]
    masks = [torch.randint(-2 ** 31, 2 ** 31 - 1, (M, 4), generator=g, dtype=torch.int64).to(torch.int32).to(DEV) for _ in range(4)]
    feat_prev = _planes(torch.randn(M, 128, generator=g), K2)                          # d_feat from the signal path
    out_kinds = [K3, K3, K2, K2, K2]

    def run_reference():
        acc = PlanePair(feat_prev.buf.clone())
        outs, h = [], g0
        for li in range(5):
            if li == 2:
                ops.umma_nt(h, wts[li], ops.UMMA_MASK | ops.UMMA_ACCUM, acc, mask=masks[li])
                y = acc
            else:
                y = PlanePair.zeros(M, dims[li + 1], DEV, kind=out_kinds[li])
                ops.umma_nt(h, wts[li], ops.UMMA_MASK if li < 4 else 0, y, mask=masks[li] if li < 4 else None)
            outs.append(y)
            h = y
        return outs

    ref = run_reference()
    acc = PlanePair(feat_prev.buf.clone())
    mine = [PlanePair.zeros(M, dims[li + 1], DEV, kind=out_kinds[li]) if li != 2 else acc for li in range(5)]
    layers = [dict(w=wts[li], save=mine[li], mask=masks[li] if li < 4 else None, accumulate=(li == 2)) for li in range(5)]
    ops.mlp_chain(g0, layers)
    torch.cuda.synchronize()
    for li, (a, b) in enumerate(zip(mine, ref)):
        assert torch.equal(a.buf[:, :, :dims[li + 1]], b.buf[:, :, :dims[li + 1]]), li
    assert float(ops.planes_merge(mine[-1]).abs().max()) > 0


def test_chain_argument_checks(built_library):
    from avr_b200 import _lib
    x0 = _planes(torch.randn(64, 48))
    w_bad = _planes(torch.randn(128, 64))                                              # k_in does not match
    with pytest.raises(_lib.AVRLibraryError):
        ops.mlp_chain(x0, [{"w": w_bad, "relu": True}])
    w0, w1 = _planes(torch.randn(64, 48)), _planes(torch.randn(16, 64))               # a 64-wide hidden layer
    with pytest.raises(_lib.AVRLibraryError):
        ops.mlp_chain(x0, [{"w": w0, "relu": True}, {"w": w1, "relu": False, "out_f32": torch.zeros(64, 16, device=DEV)}])
