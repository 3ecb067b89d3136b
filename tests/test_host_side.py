"""CPU tests of the host-side logic: tables, configs, C-ABI surface, loud failure, gradient arena."""
import os
import re

import pytest
import torch

import avr_b200
from avr_b200 import _lib, tables
from avr_b200.configs import CONFIGS, get_config
from oracle import field_ref, render_ref
from oracle.reference_shim import REFERENCE_ROOT, reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported_and_bound(built_library):
    header = open(os.path.join(ROOT, "include", "avr_b200.h")).read()
    declared = set(re.findall(r"AVR_API\s+[\w\s\*]+?\b(avr_\w+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(built_library, name)
    assert built_library.avr_abi_version() == 2


def test_renderer_refuses_cpu_tensors():
    from avr_b200.configs import tiny_config
    cfg = tiny_config()
    ren = avr_b200.AVRRender(avr_b200.AVRModel(cfg["model"]), **cfg["render"])
    with pytest.raises(_lib.AVRLibraryError):
        ren(torch.zeros(1, 3), torch.zeros(1, 3))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_tables_bit_exact_vs_oracle(name):
    cfg = get_config(name)
    r, T = cfg["render"], cfg["model"]["signal_output_dim"]
    ref = render_ref.static_tables(r, T)
    tab = tables.RenderTables(r, T, "cpu")
    assert torch.equal(tab.host["d"], ref["d"])
    assert torch.equal(tab.host["tau"], ref["tau"])
    assert torch.equal(tab.host["shift"], ref["shift"])
    assert torch.equal(tab.host["pl"], ref["pl"])
    assert torch.equal(tab.dev["delta"], ref["delta"])
    assert torch.equal(torch.view_as_complex(tab.dev["phase"]), ref["phase"].to(torch.complex64))
    # gain = tail mask * shifted path loss, exactly the product the reference applies in two steps
    t = torch.arange(T)
    tail = ((torch.arange(T - 1, -1, -1)[None, :] - ref["shift"][:, None]) > 0).float()
    pl_all = torch.stack([ref["pl"][int(i):int(i) + T] for i in ref["shift"]])
    assert torch.equal(tab.dev["gain"], tail * pl_all)
    torch.manual_seed(4)
    a = tables.direction_table(r["n_azi"], r["n_ele"])
    torch.manual_seed(4)
    b = render_ref.direction_table(r["n_azi"], r["n_ele"])
    assert torch.equal(a, b)


def test_dft_matrix_matches_rfft():
    T = 200
    m = tables.dft_matrix(T).double()
    x = torch.randn(5, T, dtype=torch.float64)
    spec = torch.fft.rfft(x, dim=-1)
    got = x @ m
    F = T // 2 + 1
    assert torch.allclose(got[:, 0:2 * F:2], spec.real, atol=1e-5)
    assert torch.allclose(got[:, 1:2 * F:2], spec.imag, atol=1e-5)
    assert m.shape[1] % 4 == 0 and float(m[:, 2 * F:].abs().max() if m.shape[1] > 2 * F else 0) == 0


@pytest.mark.parametrize("name", list(CONFIGS))
def test_grid_geometry_matches_oracle_and_survey(name):
    from avr_b200.model import hashgrid_geometry
    for key, sub in get_config(name)["model"].items():
        if isinstance(sub, dict) and sub.get("otype") == "HashGrid":
            a, b = hashgrid_geometry(sub), field_ref.hashgrid_geometry(sub)
            assert a == b
            if sub["log2_hashmap_size"] == 18:
                assert a["total"] * 2 == 9510912          # SURVEY 8 "derived sizes"
                assert a["size"][:4] == [4096, 32768, 262144, 262144]
            else:
                assert a["total"] * 2 == 36249600


@pytest.mark.parametrize("log2_size,wrapped", [(18, (12, 13, 14)), (20, (12, 13, 14, 15, 16))])
def test_grid_index_stride_modes_known_answers(log2_size, wrapped):
    """tiny-cuda-nn keeps the dense-index stride of ``grid_index`` in a uint32: at res = 2^16 .. hashmap size the second
    ``stride *= res`` wraps to 0 and the level is NOT hashed (index = (x + y*res mod 2^32) % size).  Hand-computed
    expectations for both modes of the oracle, level by level (the GPU kernels are checked against the oracle in
    tests/test_gpu_kernels.py::test_grid_encode_fwd_bwd_vs_oracle)."""
    base = {"base_resolution": 16, "log2_hashmap_size": log2_size, "n_features_per_level": 2, "n_levels": 20, "otype": "HashGrid"}
    u = torch.tensor([[0.3, 0.6, 0.9], [0.123, 0.987, 0.5]])
    size = 1 << log2_size
    for mode in ("uint32", "exact"):
        enc = field_ref.HashGridRef(dict(base, index_stride=mode))
        for lvl in range(20):
            res, scale = enc.geom["res"][lvl], enc.geom["scale"][lvl]
            idx, _ = enc.corner_indices(u, lvl)
            g = torch.floor((u.double() * scale + 0.5).float()).to(torch.int64)
            x, y, z = g[:, 0], g[:, 1], g[:, 2]
            h = ((x ^ ((y * 2654435761) & 0xFFFFFFFF) ^ ((z * 805459861) & 0xFFFFFFFF)) & 0xFFFFFFFF) % size
            dense = res ** 3 <= size
            if dense:
                want = x + y * res + z * res * res
            elif mode == "uint32" and lvl in wrapped:
                want = ((x + y * res) & 0xFFFFFFFF) % size
            else:
                want = h
            assert torch.equal(idx[:, 0], want), (mode, lvl)
        assert [l for l in range(20) if 2 ** 16 <= enc.geom["res"][l] <= size] == list(wrapped)


def test_parameter_counts_match_survey():
    for name, total in (("simu", 30089216), ("meshrir", 57237504), ("raf_furnished", 58994688)):
        cfg = get_config(name)
        cls = avr_b200.AVRModel if cfg["model_class"] == "AVRModel" else avr_b200.AVRModel_complex
        net = cls(cfg["model"])
        assert sum(p.numel() for p in net.parameters()) == total, name
        names = {k for k, _ in net.named_parameters()}
        assert "_pos_encoding.params" in names and "_model_signal.params" in names


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
def test_configs_match_reference_yaml():
    import yaml
    files = {"simu": "avr_simu.yml", "meshrir": "avr_meshrir.yml", "raf_furnished": "avr_raf_furnished.yml",
             "real_exp_ch_emb_1": "avr_real_exp_ch_emb_1.yml"}
    for name, fn in files.items():
        y = yaml.safe_load(open(os.path.join(REFERENCE_ROOT, "config_files", fn)))
        ours = get_config(name)
        assert ours["render"] == y["render"], name
        assert ours["model"] == y["model"], name
        assert ours["dataset_type"] == y["path"]["dataset_type"]


def test_state_dict_interchange_with_oracle():
    from avr_b200.configs import tiny_config
    for mc, ours, ref in (("AVRModel", avr_b200.AVRModel, field_ref.AVRModelRef),
                          ("AVRModel_complex", avr_b200.AVRModel_complex, field_ref.AVRModelComplexRef)):
        cfg = tiny_config(mc)
        a, b = ours(cfg["model"]), ref(cfg["model"])
        a.load_state_dict(b.state_dict())
        for (ka, va), (kb, vb) in zip(sorted(a.state_dict().items()), sorted(b.state_dict().items())):
            assert ka == kb and torch.equal(va, vb)


@pytest.mark.parametrize("conn", ["add", "concat"])
def test_channel_embedding_parameters_and_plan(conn):
    """model.py:71-181: parameter names/shapes of the 'add' / 'concat' variants, and what the fused step is handed."""
    from avr_b200.configs import tiny_config
    cfg = tiny_config("AVRModel")
    cfg["model"]["channel_embed"] = {"is_embed": True, "ch_num": 8, "connection_type": conn, "is_sigma_encoder": True,
                                     "is_sigma_decoder": False, "is_signal_network": True, "emb_dim_sigma_encoder": 8,
                                     "emb_dim_signal_network": 24}
    a, b = avr_b200.AVRModel(cfg["model"]), field_ref.AVRModelRef(cfg["model"])
    assert [(k, tuple(v.shape)) for k, v in a.state_dict().items()] == [(k, tuple(v.shape)) for k, v in b.state_dict().items()]
    a.load_state_dict(b.state_dict())
    assert (a.encoder_mode, a.decoder_mode, a.signal_mode) == (("injection", "none", "injection") if conn == "add"
                                                              else ("concat", "none", "concat"))
    ch = torch.tensor([3, 3, 0])
    plan = a.fused_plan(ch)
    if conn == "add":
        n_enc, n_sig = cfg["model"]["sigma_encoder_network"]["n_hidden_layers"], cfg["model"]["signal_network"]["n_hidden_layers"]
        assert plan["extras"] == [("bias", "enc", i) for i in range(n_enc)] + [("bias", "sig", i) for i in range(n_sig)]
        assert torch.equal(plan["extra_tensors"][0], a._model_encoder_sigma.layer_embeddings[0][ch])
        assert a._model_encoder_sigma.params.numel() == sum(o * i for o, i in a._model_encoder_sigma.shapes)
        assert a.fused_plan(None)["extras"] == []                        # no ch_idx: nothing is injected (model.py:58)
    else:
        assert plan["extras"] == [("rows", "encoder_channel_embedding"), ("rows", "signal_channel_embedding")]
        assert [k for _, k in plan["x0"]] == ["point", "receiver_rows"] and plan["tail"][-1][1] == "receiver_rows"
        assert a._model_encoder_sigma.n_input_dims == a._pos_encoding.n_output_dims + 8
        with pytest.raises(ValueError):
            a.fused_plan(None)


def test_grad_arena_views_and_rebind():
    p1, p2 = torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(3, 3))
    arena = avr_b200.GradArena([p1, p2])
    (p1.sum() * 2 + (p2 * 3).sum()).backward()
    assert p1.grad.data_ptr() == arena.flat.data_ptr()
    assert torch.equal(arena.flat[:5], torch.full((5,), 2.0))
    assert torch.equal(arena.flat[8:17], torch.full((9,), 3.0))
    arena.zero_()
    assert float(arena.flat.abs().sum()) == 0 and p2.grad.data_ptr() == arena.flat[8:].data_ptr()


def test_plane_set_kinds_and_sample_step(built_library):
    """Host-side description of the tensor-core operands (include/avr_b200.h, plane-set kinds) and of the scatter's
    run-merging hint."""
    from avr_b200 import ops
    header = open(os.path.join(ROOT, "include", "avr_b200.h")).read()
    kinds = dict(re.findall(r"(AVR_PLANES_\w+)\s*=\s*(\d+)", header))
    assert (int(kinds["AVR_PLANES_BF16x2"]), int(kinds["AVR_PLANES_BF16x3"]), int(kinds["AVR_PLANES_F16x2"])) == \
        (ops.PLANES_BF16x2, ops.PLANES_BF16x3, ops.PLANES_F16x2)
    flags = dict(re.findall(r"(AVR_UMMA_\w+)\s*=\s*(\d+)", header))
    assert int(flags["AVR_UMMA_DUAL_COPY"]) == ops.UMMA_DUAL_COPY and int(flags["AVR_UMMA_BIAS"]) == ops.UMMA_BIAS
    a = ops.PlanePair.empty(10, 20, "cpu", n=3)
    b = ops.PlanePair.empty(10, 20, "cpu", kind=ops.PLANES_F16x2)
    assert (a.kind, a.n, a.f16, a.ld) == (3, 3, False, 24) and (b.kind, b.n, b.f16) == (ops.PLANES_F16x2, 2, True)
    assert b.window(8, 8).cols == 8 and b.window(8, 8).kind == ops.PLANES_F16x2
    with pytest.raises(AssertionError):
        ops.PlanePair(torch.empty(3, 4, 8, dtype=torch.float16))           # fp16 sets are pairs
    cfg = get_config("simu")["render"]
    t = tables.RenderTables(cfg, 1600, "cpu")
    step = (cfg["far"] - cfg["near"]) / (cfg["n_samples"] - 1) / (cfg["xyz_max"] - cfg["xyz_min"])
    assert abs(t.sample_step - step) < 1e-12 and t.dev["sample_step"] == t.sample_step
    # levels whose cells hold >= 2 consecutive samples (scale * step < 0.5) get the merged scatter: 0..2 at simu
    geo = field_ref.hashgrid_geometry(get_config("simu")["model"]["pos_encoding_sigma"])
    assert sum(1 for s in geo["scale"] if s * step < 0.5) == 3
