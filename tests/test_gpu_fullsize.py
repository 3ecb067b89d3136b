"""Parity at the REAL sizes of the BASELINE configs ("no BASELINE config is parity-checked at its real size").

* all four BASELINE configs at their real ray counts (simu R = 2050, S = 64, T = 1600; MeshRIR R = 3202, T = 2400; ...):
  the shipped tensor-core path against the CPU oracle (renderer_cpu.py restatement + fp32 field), rendered IR and EVERY
  parameter gradient within 1e-4 rel-L2, both table-gradient modes -- this is what reaches ``delay_sort`` at
  R = 2050 / 3202, ``prefix_walk`` with its full ``tspan``, the 128 x 256 fp16-pair tiles at M = 524 800 and the
  run-merged scatter on real run lengths.  The independent fp32 SIMT path of this library runs beside it.
* the real-field tests on reduced ray grids loop over ALL candidate weight seeds and report, per seed, the oracle's own
  fp32 noise and the tc / simt distances; nothing is selected silently.

Every test appends its numbers to ``gpurun_out/parity_fullsize.jsonl`` (when that directory exists) -- the source of
``profiles/r2/parity_fullsize.md``.
"""
import json
import os
import time

import pytest
import torch

import avr_b200
from avr_b200.configs import get_config
from oracle import field_ref, render_ref
from tests.helpers import oracle_conditioning, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record(**kw):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_fullsize.jsonl"), "a") as fh:
            fh.write(json.dumps(kw) + "\n")


def _inputs(cfg, bs, seed=11, spread=1.5):
    r, mc = cfg["render"], cfg["model_class"]
    gen = torch.Generator().manual_seed(seed)
    c = (r["xyz_min"] + r["xyz_max"]) / 2
    rx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * spread).float()
    tx = (c + (torch.rand(bs, 3, generator=gen) * 2 - 1) * spread).float()
    dtx = torch.nn.functional.normalize(torch.randn(bs, 3, generator=gen), dim=-1) if mc != "AVRModel" else None
    azi = torch.rand(r["n_azi"], generator=gen)
    T = cfg["model"]["signal_output_dim"]
    G = torch.randn(bs, T // 2 + 1, 2, generator=gen)
    return rx, tx, dtx, azi, G


def _native(cfg, state_dict):
    cls = avr_b200.AVRModel if cfg["model_class"] == "AVRModel" else avr_b200.AVRModel_complex
    net = cls(cfg["model"])
    net.load_state_dict(state_dict)
    return net.to(DEV)


def _run(native, cfg, rx, tx, dtx, azi, G, **ren_kw):
    native.zero_grad(set_to_none=True)
    ren = avr_b200.AVRRender(native, **cfg["render"], **ren_kw)
    out = ren(rx.to(DEV), tx.to(DEV), dtx.to(DEV) if dtx is not None else None, azi_rand=azi)
    (out * G.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    return out.detach().cpu(), {n: p.grad.detach().cpu().clone() for n, p in native.named_parameters()}


def _errs(out, grads, ref_out, ref_grads):
    return rel_l2(out, ref_out), {n: rel_l2(g, ref_grads[n]) for n, g in grads.items()}


FLIP_NOISE = 5e-3     # hash-table gradients of the fp32 SIMT path at full size, see test_full_size_vs_cpu_oracle


@pytest.mark.parametrize("name,bs", [("simu", 1), ("simu", 2), ("meshrir", 1), ("raf_furnished", 2), ("real_exp_ch_emb_1", 1)])
def test_full_size_vs_cpu_oracle(built_library, name, bs):
    """Every BASELINE config at its REAL ray count against the CPU oracle (3-10 s of host time per receiver on the GPU
    box): the shipped tensor-core path must hold IR and EVERY parameter gradient to 1e-4, in both accumulation modes of
    the hash-table gradients.

    The independent fp32 SIMT path (exact FMA GEMMs, literal ``sig[bs,R,S,T]`` tensor, ``composite_*`` kernels) is run on
    the same inputs: its IR and its dense-layer gradients must agree as well; its HASH-TABLE gradients are reported and
    held to a looser bound only.  At 4-8e7 ReLU units per receiver a few dozen pre-activations lie within fp32 rounding
    noise of zero; the SIMT GEMM (one sequential fp32 accumulation chain, error ~1e-6 of the row scale) decides some of
    them differently from the oracle's blocked CPU GEMM, and every flipped unit gates its back-propagated gradient on or
    off (measured: 1.8e-4 .. 1.5e-3 on the tables, first run of this test).  The tensor-core path's two-accumulator
    products (2e-7) flip fewer decisions than either fp32 GEMM, which is why it is the one that meets the bar."""
    cfg = get_config(name)
    cls = field_ref.AVRModelRef if cfg["model_class"] == "AVRModel" else field_ref.AVRModelComplexRef
    ref_net = field_ref.trained_like_(cls(cfg["model"], seed=41), seed=42)
    rx, tx, dtx, azi, G = _inputs(cfg, bs)
    t0 = time.time()
    ref_out = render_ref.RenderRef(ref_net, **cfg["render"])(rx, tx, dtx, azi_rand=azi)
    (ref_out * G).sum().backward()
    cpu_s = time.time() - t0
    ref_out = ref_out.detach()
    ref_grads = {n: p.grad for n, p in ref_net.named_parameters()}
    native = _native(cfg, ref_net.state_dict())
    res, raw = {}, {}
    for key, kw in (("tc", dict(dense="tc", grid_grad="deterministic")), ("tc_atomic", dict(dense="tc", grid_grad="atomic")),
                    ("simt", dict(dense="simt", grid_grad="deterministic"))):
        out, grads = _run(native, cfg, rx, tx, dtx, azi, G, **kw)
        raw[key] = (out, grads)
        res[key] = _errs(out, grads, ref_out, ref_grads)
    worst = {d: max(res[d][1].values()) for d in res}
    tc_vs_simt = _errs(raw["tc"][0], raw["tc"][1], raw["simt"][0], raw["simt"][1])
    r = cfg["render"]
    print(f"\n{name} full size (R={r['n_azi'] * r['n_ele'] + 2}, S={r['n_samples']}, T={cfg['model']['signal_output_dim']}, bs={bs}), "
          f"CPU oracle fwd+bwd {cpu_s:.1f} s: tc IR {res['tc'][0]:.2e} worst grad {worst['tc']:.2e}; tc[atomic] worst grad "
          f"{worst['tc_atomic']:.2e}; simt IR {res['simt'][0]:.2e} worst grad {worst['simt']:.2e}; tc vs simt worst grad "
          f"{max(tc_vs_simt[1].values()):.2e}")
    bar, noise = TOL, None
    if res["tc"][0] >= TOL or worst["tc"] >= TOL or worst["tc_atomic"] >= TOL:
        # an ill-conditioned draw: measure how far fp32 arithmetic itself leaves the answer open (oracle_conditioning) --
        # only then, it multiplies the host time -- and hold the product to max(1e-4, 2 x that spread)
        n_out, n_g = oracle_conditioning(ref_net, cfg["render"], rx, tx, G, dtx=dtx, azi_rand=azi)
        noise = max([n_out] + list(n_g.values()))
        bar = max(TOL, 2 * noise)
    _record(test="full_size_vs_cpu_oracle", config=name, bs=bs, cpu_oracle_seconds=cpu_s, oracle_fp32_noise=noise,
            **{k: {"ir": v[0], "grads": v[1]} for k, v in res.items()},
            tc_vs_simt={"ir": tc_vs_simt[0], "grads": tc_vs_simt[1]})
    assert float(ref_out.abs().max()) > 0
    for key in ("tc", "tc_atomic"):
        assert res[key][0] < bar, (key, res[key][0])
        for n, e in res[key][1].items():
            assert e < bar, (key, n, e, bar)
    assert res["simt"][0] < 1e-5 and tc_vs_simt[0] < 1e-5
    for n, e in res["simt"][1].items():
        assert e < (FLIP_NOISE if "encoding" in n else max(bar, 2e-4)), ("simt", n, e)


@pytest.mark.parametrize("name,n_azi,n_ele", [("simu", 16, 8), ("meshrir", 10, 6), ("raf_furnished", 12, 6),
                                              ("real_exp_ch_emb_1", 16, 8)])
def test_real_fields_all_seeds_reported(built_library, name, n_azi, n_ele):
    """The real fields on a reduced ray grid, EVERY candidate seed.  The oracle's own spread (tests/helpers.py::
    oracle_conditioning: fp32 evaluations in other summation orders against the float64-dense evaluation, and the
    gradients with every ReLU decision inside fp32 rounding noise of zero flipped) says how well fp32 arithmetic
    determines the answer on that draw: where it is <= 5e-5 the bar is 1e-4; elsewhere (a decision within 1e-6 of zero at
    a point that carries a visible share of the gradient) the bar is max(1e-4, 2 x spread).  Per seed: spread, tc and
    simt distances -- printed and recorded; no seed is skipped."""
    cfg = get_config(name)
    cfg["render"]["n_azi"], cfg["render"]["n_ele"] = n_azi, n_ele
    mc = cfg["model_class"]
    cls = field_ref.AVRModelRef if mc == "AVRModel" else field_ref.AVRModelComplexRef
    rx, tx, dtx, azi, G = _inputs(cfg, 2)
    rows, rejected = [], 0
    for seed in range(41, 51, 2):
        ref_net = field_ref.trained_like_(cls(cfg["model"], seed=seed), seed=seed + 1)
        n_out, n_g = oracle_conditioning(ref_net, cfg["render"], rx, tx, G, dtx=dtx, azi_rand=azi)
        noise = max([n_out] + list(n_g.values()))
        well = 2 * noise <= TOL                                   # <=> this seed is held to the plain 1e-4 bar below
        rejected += 0 if well else 1
        ref_out = render_ref.RenderRef(ref_net, **cfg["render"])(rx, tx, dtx, azi_rand=azi)
        (ref_out * G).sum().backward()
        ref_grads = {n: p.grad for n, p in ref_net.named_parameters()}
        native = _native(cfg, ref_net.state_dict())
        e = {}
        for dense in ("tc", "simt"):
            out, grads = _run(native, cfg, rx, tx, dtx, azi, G, dense=dense)
            e_out, e_g = _errs(out, grads, ref_out.detach(), ref_grads)
            e[dense] = (e_out, max(e_g.values()), max(e_g, key=e_g.get))
        rows.append({"seed": seed, "oracle_noise": noise, "well_conditioned": well, "tc_ir": e["tc"][0], "tc_worst_grad": e["tc"][1],
                     "tc_worst_param": e["tc"][2], "simt_ir": e["simt"][0], "simt_worst_grad": e["simt"][1]})
        print(f"\n{name} seed {seed}: oracle noise {noise:.1e} ({'well' if well else 'ILL'}-conditioned)  tc IR {e['tc'][0]:.1e} grad "
              f"{e['tc'][1]:.1e} ({e['tc'][2]})  simt IR {e['simt'][0]:.1e} grad {e['simt'][1]:.1e}")
        bar = TOL if well else max(TOL, 2 * noise)
        assert e["tc"][0] < bar and e["tc"][1] < bar, (seed, e["tc"], bar)
    print(f"{name}: {rejected} of {len(rows)} seeds ill-conditioned for fp32 arithmetic")
    _record(test="real_fields_all_seeds", config=name, rays=n_azi * n_ele + 2, ill_conditioned=rejected, seeds=rows)
    assert rejected < len(rows), "no seed on which the oracle can check at 1e-4"
