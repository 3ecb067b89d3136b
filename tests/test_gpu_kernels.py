"""Kernel-level parity on the GPU, through the C-ABI: geometry, hash grid, dense layers, compositing."""
import numpy as np
import pytest
import torch

from avr_b200 import ops, tables
from avr_b200._lib import GEMM_ACCUM, GEMM_MASK, GEMM_RELU, GEMM_RELU_A, GEMM_RELU_B, I_CONTIG, K_CONTIG
from avr_b200.configs import get_config, tiny_config
from avr_b200.model import hashgrid_geometry
from oracle import field_ref, render_ref
from tests.helpers import GOLDEN_CASES, case_config, load_golden, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(render, bs, seed, half):
    g = torch.Generator().manual_seed(seed)
    c = (render["xyz_min"] + render["xyz_max"]) / 2
    rx = (c + (torch.rand(bs, 3, generator=g) * 2 - 1) * half).float()
    tx = (c + (torch.rand(bs, 3, generator=g) * 2 - 1) * half).float()
    azi = torch.rand(render["n_azi"], generator=g)
    return rx, tx, azi


# ---------------------------------------------------------------------------------------------- geometry
@pytest.mark.parametrize("name,bs", [("simu", 2), ("real_exp_ch_emb_1", 1), ("meshrir", 1), ("raf_furnished", 3)])
def test_sample_points_bit_exact_full_configs(built_library, name, bs):
    cfg = get_config(name)
    r, T = cfg["render"], cfg["model"]["signal_output_dim"]
    rx, tx, azi = _inputs(r, bs, 5, 3.0)
    dirs = render_ref.direction_table(r["n_azi"], r["n_ele"], azi)
    tab = render_ref.static_tables(r, T)
    pts_n, view, tx_n, _ = render_ref.sample_geometry(rx, tx, dirs, tab["d"], r)
    delay = render_ref.source_delay(pts_n, tx_n, r, T, r["n_samples"])
    geom = ops.make_geom(r, bs, T)
    p, v, t, d = ops.sample_points(geom, rx.to(DEV), tx.to(DEV), dirs.to(DEV), tab["d"].to(DEV))
    assert torch.equal(p.cpu(), pts_n)                   # bit-exact sample positions
    assert torch.equal(v.cpu(), view)
    assert torch.equal(t.cpu(), tx_n)
    assert torch.equal(d.cpu().float(), delay)           # bit-exact delay indices
    assert int(d.min()) >= 0 and int(d.max()) <= T - 1


@pytest.mark.parametrize("name", GOLDEN_CASES[:3])
def test_sample_points_vs_golden(built_library, name):
    g = load_golden(name)
    _, cfg = case_config(name)
    r, T = cfg["render"], cfg["model"]["signal_output_dim"]
    geom = ops.make_geom(r, g["rx"].shape[0], T)
    d = tables.RenderTables(r, T, DEV).dev["d"]
    p, v, t, _ = ops.sample_points(geom, g["rx"].to(DEV), g["tx"].to(DEV), g["dirs"].to(DEV), d)
    assert torch.equal(p.cpu(), g["net_pts"])
    step = max(1, v.shape[1] // 16)
    assert torch.equal(v.cpu()[:, ::step], g["net_view"])
    assert torch.equal(t.cpu()[:, ::step], g["net_tx"])


def test_empty_batch_is_a_noop(built_library):
    r = tiny_config()["render"]
    geom = ops.make_geom(r, 0, 200)
    dirs = render_ref.direction_table(r["n_azi"], r["n_ele"], torch.zeros(r["n_azi"])).to(DEV)
    d = torch.linspace(0, 6, r["n_samples"], device=DEV)
    p, v, t, dl = ops.sample_points(geom, torch.zeros(0, 3, device=DEV), torch.zeros(0, 3, device=DEV), dirs, d)
    assert p.shape[0] == 0 and dl.numel() == 0


# ---------------------------------------------------------------------------------------------- hash grid
GRID_CFGS = {
    "tiny": {"base_resolution": 4, "log2_hashmap_size": 10, "n_features_per_level": 2, "n_levels": 6, "otype": "HashGrid"},
    "simu": get_config("simu")["model"]["pos_encoding_sigma"],
    "mesh_dir": get_config("meshrir")["model"]["dir_encoding_sig"],
}


@pytest.mark.parametrize("stride", ["uint32", "exact"])
@pytest.mark.parametrize("which,n", [("tiny", 1000), ("simu", 4099), ("mesh_dir", 513)])
def test_grid_encode_fwd_bwd_vs_oracle(built_library, which, n, stride):
    """Both grid_index arithmetics (tiny-cuda-nn's wrapping uint32 stride -- levels 12..14 / 12..16 of the 2^18 / 2^20
    grids not hashed -- and the exact stride) against the oracle's restatement of each."""
    cfg = dict(GRID_CFGS[which], index_stride=stride)
    enc = field_ref.HashGridRef(cfg, seed=3)
    with torch.no_grad():
        enc.params.copy_(torch.randn(enc.params.shape, generator=torch.Generator().manual_seed(1)) * 0.1)
    g = torch.Generator().manual_seed(2)
    u = torch.rand(n, 3, generator=g)
    u[:7] = torch.tensor([[0., 0., 0.], [1., 1., 1.], [0.5, 0.5, 0.5], [1., 0., 1.], [-0.25, 0.3, 1.2],
                          [0.999999, 1e-7, 0.25], [1.5, -0.5, 0.75]])            # edges + out-of-cube wrap-around
    ref = enc(u)
    G = torch.randn(ref.shape, generator=g)
    (ref * G).sum().backward()
    meta = ops.make_grid_meta(hashgrid_geometry(cfg))
    W = ref.shape[1]
    out = torch.full((n, W + 8), -7.0, device=DEV)
    ops.grid_encode_fwd(meta, u.to(DEV), enc.params.detach().to(DEV), out, col0=4, n_ones=3)
    got = out.cpu()
    assert torch.allclose(got[:, 4:4 + W], ref.detach(), rtol=0, atol=2e-7 * float(ref.abs().max()) + 1e-9)
    assert torch.all(got[:, :4] == -7.0) and torch.all(got[:, 4 + W:4 + W + 3] == 1.0) and torch.all(got[:, 4 + W + 3:] == -7.0)
    # backward: deterministic fixed-point accumulation
    dbuf = torch.zeros(n, W + 8, device=DEV)
    dbuf[:, 4:4 + W] = G.to(DEV)
    grads = []
    for _ in range(2):
        acc = ops.GridGradAccumulator(meta, DEV, n)
        acc.observe(dbuf, 4, W)
        acc.add_points(u.to(DEV), dbuf, col0=4)
        grads.append(acc.finalize().cpu())
    assert torch.equal(grads[0], grads[1])                                        # bit-identical reruns
    assert rel_l2(grads[0], enc.params.grad) < 1e-5
    # backward, fp32 vector reductions straight into the gradient (AVR_GRID_GRAD_F32)
    acc = ops.GridGradAccumulator(meta, DEV, n, mode="atomic")
    acc.observe(dbuf, 4, W)
    acc.add_points(u.to(DEV), dbuf, col0=4)
    fast = acc.finalize().cpu()
    assert rel_l2(fast, enc.params.grad) < 1e-5 and rel_l2(fast, grads[0]) < 1e-5
    again = torch.ones_like(grads[0]).to(DEV)
    acc.finalize(again, accumulate=True)
    assert torch.equal(again.cpu(), fast + 1.0)


def test_grid_grad_zero_and_nan_inputs(built_library):
    cfg = GRID_CFGS["tiny"]
    meta = ops.make_grid_meta(hashgrid_geometry(cfg))
    u = torch.rand(64, 3, device=DEV)
    d = torch.zeros(64, 12, device=DEV)
    acc = ops.GridGradAccumulator(meta, DEV, 64)
    acc.observe(d, 0, 12)
    acc.add_points(u, d)
    assert float(acc.finalize().abs().max()) == 0.0
    d[3, 5] = float("nan")
    acc = ops.GridGradAccumulator(meta, DEV, 64)
    acc.observe(d, 0, 12)
    acc.add_points(u, d)
    assert bool(torch.isnan(acc.finalize()).all())                                # poisoned gradient is reported, not hidden


def test_raygen_encode_matches_explicit_points(built_library):
    cfg = tiny_config(n_azi=7, n_ele=5, n_samples=9)
    r, T = cfg["render"], 200
    bs = 3
    rx, tx, azi = _inputs(r, bs, 8, 3.0)
    dirs = render_ref.direction_table(r["n_azi"], r["n_ele"], azi)
    tab = render_ref.static_tables(r, T)
    pts_n, _, tx_n, _ = render_ref.sample_geometry(rx, tx, dirs, tab["d"], r)
    delay = render_ref.source_delay(pts_n, tx_n, r, T, r["n_samples"])
    gcfg = cfg["model"]["pos_encoding_sigma"]
    enc = field_ref.HashGridRef(gcfg, seed=5)
    with torch.no_grad():
        enc.params.normal_(0, 0.1, generator=torch.Generator().manual_seed(1))
    ref = enc(((pts_n.reshape(-1, 3) + 1) / 2))
    geom = ops.make_geom(r, bs, T)
    meta = ops.make_grid_meta(hashgrid_geometry(gcfg))
    n, W = ref.shape
    out = torch.zeros(n, 16, device=DEV)
    dl = torch.empty(bs, geom.R, geom.S, dtype=torch.int32, device=DEV)
    ops.raygen_encode_fwd(geom, meta, rx.to(DEV), tx.to(DEV), dirs.to(DEV), tab["d"].to(DEV), enc.params.detach().to(DEV),
                          out, col0=0, n_ones=4, delay=dl)
    assert torch.allclose(out[:, :W].cpu(), ref.detach(), rtol=0, atol=1e-7)
    assert torch.all(out[:, W:] == 1.0)
    assert torch.equal(dl.cpu().float(), delay)
    G = torch.randn(n, W, generator=torch.Generator().manual_seed(4))
    (ref * G).sum().backward()
    dbuf = torch.zeros(n, 16, device=DEV)
    dbuf[:, :W] = G.to(DEV)
    acc = ops.GridGradAccumulator(meta, DEV, n)
    acc.observe(dbuf, 0, W)
    acc.add_rays(geom, rx.to(DEV), dirs.to(DEV), tab["d"].to(DEV), dbuf, 0)
    plain = acc.finalize().cpu()
    assert rel_l2(plain, enc.params.grad) < 1e-5
    # run merging (avr_raygen_encode_bwd's sample_step): consecutive samples of a ray that add to the same coarse-level
    # entry are summed with warp shuffles and issue one reduction.  Integer addends: bit-identical to the plain scatter.
    step = float(r["far"] - r["near"]) / (r["n_samples"] - 1) / float(r["xyz_max"] - r["xyz_min"])
    assert step * meta.scale[0] < 0.5                                             # at least the coarsest level merges
    acc = ops.GridGradAccumulator(meta, DEV, n)
    acc.observe(dbuf, 0, W)
    acc.add_rays(geom, rx.to(DEV), dirs.to(DEV), tab["d"].to(DEV), dbuf, 0, sample_step=step)
    assert torch.equal(acc.finalize().cpu(), plain)
    acc = ops.GridGradAccumulator(meta, DEV, n, mode="atomic")
    acc.add_rays(geom, rx.to(DEV), dirs.to(DEV), tab["d"].to(DEV), dbuf, 0, sample_step=step)
    assert rel_l2(acc.finalize(), enc.params.grad) < 1e-5


# ---------------------------------------------------------------------------------------------- dense layers
@pytest.mark.parametrize("M,N,K", [(300, 128, 48), (1000, 16, 128), (257, 200, 512), (64, 1604, 1600), (5, 4, 4)])
def test_gemm_nt_flags(built_library, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    C0, aux = torch.randn(M, N, generator=g), torch.randn(M, N, generator=g)
    Ad, Bd, auxd = A.to(DEV), B.to(DEV), aux.to(DEV)
    ref = A.double() @ B.double().t()
    for flags, fn in ((0, lambda p: p), (GEMM_RELU, lambda p: p.clamp_min(0)),
                      (GEMM_MASK | GEMM_ACCUM, lambda p: p * (aux > 0) + C0),
                      (GEMM_RELU_A | GEMM_RELU, lambda p: (A.clamp_min(0).double() @ B.double().t()).clamp_min(0))):
        Cd = C0.clone().to(DEV)
        ops.gemm(K_CONTIG, K_CONTIG, M, N, K, Ad, K, Bd, K, Cd, N, flags, auxd, N)
        assert rel_l2(Cd, fn(ref)) < 2e-6, flags


def test_gemm_column_views_and_other_layouts(built_library):
    g = torch.Generator().manual_seed(0)
    n, m, k = 777, 64, 40
    X = torch.randn(n, 96, generator=g).to(DEV)          # use columns 8..48
    W = torch.randn(m, 56, generator=g).to(DEV)          # use columns 16..56
    dY = torch.randn(n, m, generator=g).to(DEV)
    x, w = X[:, 8:48], W[:, 16:56]
    y = torch.zeros(n, m, device=DEV)
    ops.linear_fwd(x, w, y)
    assert rel_l2(y, x.double().cpu() @ w.double().cpu().t()) < 2e-6
    dx = torch.zeros(n, 96, device=DEV)
    ops.linear_bwd_data(dY, w, dx[:, 8:48], mask_src=x)
    ref = (dY.double().cpu() @ w.double().cpu()) * (x.cpu() > 0)
    assert rel_l2(dx[:, 8:48], ref) < 2e-6 and float(dx[:, :8].abs().max()) == 0 and float(dx[:, 48:].abs().max()) == 0
    dw = torch.zeros(m, 56, device=DEV)
    ws = torch.empty(max(4, ops.gemm_workspace_bytes(m, k, n) // 4), device=DEV)
    ops.linear_bwd_weight(dY, x, dw[:, 16:56], ws)
    assert rel_l2(dw[:, 16:56], dY.double().cpu().t() @ x.double().cpu()) < 2e-6 and float(dw[:, :16].abs().max()) == 0
    dw2 = torch.zeros(m, k, device=DEV)
    ops.linear_bwd_weight(dY, x, dw2, ws, relu_in=True)
    assert rel_l2(dw2, dY.double().cpu().t() @ x.double().cpu().clamp_min(0)) < 2e-6


def test_weight_grad_split_k_is_deterministic(built_library):
    g = torch.Generator().manual_seed(1)
    n, m, k = 40000, 128, 48
    dY, X = torch.randn(n, m, generator=g).to(DEV), torch.randn(n, k, generator=g).to(DEV)
    assert ops.gemm_workspace_bytes(m, k, n) > 0
    ws = torch.empty(ops.gemm_workspace_bytes(m, k, n) // 4, device=DEV)
    outs = []
    for _ in range(2):
        dw = torch.empty(m, k, device=DEV)
        ops.linear_bwd_weight(dY, X, dw, ws)
        outs.append(dw.cpu())
    assert torch.equal(outs[0], outs[1])
    assert rel_l2(outs[0], dY.double().cpu().t() @ X.double().cpu()) < 2e-6


# ---------------------------------------------------------------------------------------------- compositing
def _render_case(bs=2, n_azi=5, n_ele=4, S=7, T=240, seed=0, attn_scale=0.8):
    cfg = tiny_config(n_azi=n_azi, n_ele=n_ele, n_samples=S, T=T)
    r = cfg["render"]
    rx, tx, azi = _inputs(r, bs, seed, 3.0)
    dirs = render_ref.direction_table(n_azi, n_ele, azi)
    tab = render_ref.static_tables(r, T)
    pts_n, _, tx_n, _ = render_ref.sample_geometry(rx, tx, dirs, tab["d"], r)
    delay = render_ref.source_delay(pts_n, tx_n, r, T, S)
    g = torch.Generator().manual_seed(seed + 1)
    R = n_azi * n_ele + 2
    attn = torch.rand(bs, R, S, generator=g) * attn_scale
    sig = torch.randn(bs, R, S, T, generator=g)
    return r, tab, delay, attn, sig


@pytest.mark.parametrize("S", [7, 33, 64, 80])
def test_ray_weights_fwd_bwd(built_library, S):
    bs, R = 2, 11
    g = torch.Generator().manual_seed(S)
    raw = (torch.randn(bs, R, S, generator=g) * 0.7).requires_grad_()
    raw.data[0, 0, :3] = torch.tensor([0.0, -0.5, 4.0])
    d = torch.linspace(0, 6, S)
    delta = torch.cat([d[1:] - d[:-1], torch.tensor([1e10])])
    for slope in (0.01, 0.03):
        raw.grad = None
        attn = torch.abs(torch.nn.functional.leaky_relu(raw, slope))
        w, alpha, trans = render_ref.ray_weights(attn, delta)
        Gw = torch.randn(w.shape, generator=g)
        (w * Gw).sum().backward()
        geom = ops.RenderGeom(bs, R, S, 200, -10.0, 20.0, 16000.0, 343.8)
        rawd = torch.zeros(bs * R * S, 16, device=DEV)
        rawd[:, 0] = raw.detach().reshape(-1).to(DEV)
        wd, attnd = ops.ray_weights_fwd(geom, rawd, 16, delta.to(DEV), slope, want_attn=True)
        assert torch.equal(attnd.cpu(), attn.detach())
        assert torch.allclose(wd.cpu(), w.detach(), rtol=2e-6, atol=1e-7)
        d_raw = torch.zeros_like(rawd)
        ops.ray_weights_bwd(geom, rawd, 16, delta.to(DEV), slope, Gw.to(DEV), d_raw, 16)
        assert rel_l2(d_raw[:, 0].reshape(bs, R, S), raw.grad) < 1e-5
        assert float(d_raw[:, 1:].abs().max()) == 0


@pytest.mark.parametrize("shape", [dict(), dict(T=200, S=8, n_azi=9, n_ele=7), dict(T=1600, S=5, n_azi=6, n_ele=5, bs=1),
                                   dict(T=2400, S=3, n_azi=4, n_ele=3, bs=1)])
def test_composite_and_spectrum_vs_oracle(built_library, shape):
    r, tab, delay, attn, sig = _render_case(**shape)
    bs, R, S, T = sig.shape
    attn.requires_grad_(); sig.requires_grad_()
    ref = render_ref.composite(attn, sig, delay, tab)
    G = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3))
    (ref * G).sum().backward()
    geom = ops.make_geom(r, bs, T)
    tabs = tables.RenderTables(r, T, DEV).dev
    w_ref, _, _ = render_ref.ray_weights(attn.detach(), tab["delta"])
    wd, _ = ops.ray_weights_fwd(geom, attn.detach().to(DEV), 1, tabs["delta"], -1.0)
    dl = delay.to(torch.int32).to(DEV)
    sd = sig.detach().to(DEV)
    y = ops.composite_fwd(geom, sd, wd, dl)
    t = torch.arange(T)
    y_ref = torch.sum(sig.detach() * (w_ref[..., None] * (t >= delay.unsqueeze(-1))), dim=1)
    assert rel_l2(y, y_ref) < 2e-6
    out = ops.spectrum_fwd(geom, y, tabs)
    assert rel_l2(out, ref) < 1e-5
    d_y = ops.spectrum_bwd(geom, G.to(DEV), tabs)
    d_sig, d_w = ops.composite_bwd(geom, sd, wd, dl, d_y)
    assert rel_l2(d_sig, sig.grad) < 1e-5
    d_attn = torch.empty(bs, R, S, device=DEV)
    ops.ray_weights_bwd(geom, attn.detach().to(DEV), 1, tabs["delta"], -1.0, d_w, d_attn, 1)
    assert rel_l2(d_attn, attn.grad) < 1e-4


def test_rows_broadcast_and_reduce(built_library):
    geom = ops.RenderGeom(3, 10, 6, 200, -10.0, 20.0, 16000.0, 343.8)
    n = 3 * 10 * 6
    g = torch.Generator().manual_seed(0)
    for per_receiver, rows in ((False, 10), (True, 3)):
        src = torch.randn(rows, 12, generator=g)
        dst = torch.zeros(n, 32, device=DEV)
        ops.rows_broadcast(geom, src.to(DEV), per_receiver, dst, 8)
        idx = torch.arange(n)
        row = idx // 60 if per_receiver else (idx // 6) % 10
        assert torch.equal(dst[:, 8:20].cpu(), src[row]) and float(dst[:, :8].abs().max()) == 0
        d = torch.randn(n, 32, generator=g)
        red = ops.rows_reduce(geom, d.to(DEV), 8, 12, per_receiver)
        ref = torch.zeros(rows, 12, dtype=torch.float64).index_add_(0, row, d[:, 8:20].double())
        assert rel_l2(red, ref) < 1e-6


@pytest.mark.parametrize("w,col0,nplanes", [(40, 8, 3), (40, 48, 2), (16, 0, 3), (12, 8, 3)])
def test_rows_broadcast_and_reduce_plane_sets(built_library, w, col0, nplanes):
    """bf16 plane-set windows (the signal-network input): 16-byte vector path (w, col0 multiples of 8) and scalar path."""
    from avr_b200.ops import PlanePair
    geom = ops.RenderGeom(3, 37, 6, 200, -10.0, 20.0, 16000.0, 343.8)
    n = 3 * 37 * 6
    g = torch.Generator().manual_seed(w + col0)
    idx = torch.arange(n)
    for per_receiver, rows in ((False, 37), (True, 3)):
        row = idx // (37 * 6) if per_receiver else (idx // 6) % 37
        src = torch.randn(rows, w, generator=g)
        dst = PlanePair.empty(n, 96, DEV, n=nplanes)
        dst.buf.zero_()
        ops.rows_broadcast(geom, src.to(DEV), per_receiver, dst, col0)
        got = dst.buf.float().sum(0).cpu()                       # planes add up to the fp32 value
        tol = 2.0 ** -16 if nplanes == 2 else 2.0 ** -23
        assert float((got[:, col0:col0 + w] - src[row]).abs().max()) <= tol * float(src.abs().max())
        assert float(got[:, :col0].abs().max() if col0 else 0.0) == 0 and float(got[:, col0 + w:].abs().max()) == 0
        d = PlanePair.empty(n, 96, DEV, n=nplanes)
        d.buf.normal_(generator=torch.Generator(device=DEV).manual_seed(1))
        red = ops.rows_reduce(geom, d, col0, w, per_receiver)
        two = d.buf[:2].float().sum(0).cpu()                     # readers of plane sets use the first two planes
        ref = torch.zeros(rows, w, dtype=torch.float64).index_add_(0, row, two[:, col0:col0 + w].double())
        assert rel_l2(red, ref) < 1e-6
        assert torch.equal(red, ops.rows_reduce(geom, d, col0, w, per_receiver))
