#!/usr/bin/env python
"""Headline benchmark: rendered IRs/sec (forward + backward) of the AVR render hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config simu] [--bs 4]

A *step* is one forward + backward pass of ``AVRRender`` (ray generation, hash-grid encode, MLPs,
compositing, spectrum; loss = sum of squares of the ``[bs,F,2]`` IR spectrum) over one synthetic batch of
``bs`` receivers per GPU at the ``avr_simu.yml`` shape (BASELINE.json configs[1]), plus -- for N > 1 -- the
gradient all-reduce of the data-parallel step (weak scaling: ``bs`` receivers per GPU).  No optimizer
update is part of the metric ("receivers*steps / time for forward+backward").

Two numbers per run: ``value`` (inputs resident in HBM) and ``e2e`` (host buffers: pinned H2D of the
receiver/transmitter positions and a D2H read of the rendered spectra every step).  ``--impl reference``
times the CPU oracle (``oracle/``: restatement of renderer_cpu.py + the tcnn field in fp32 torch) on the
host cores instead -- the reference is a Python repo whose own runner needs tiny-cuda-nn, so its CPU
path is the oracle port.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "rendered IRs/sec (fwd+bwd)"
UNIT = "IR/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="simu")
    ap.add_argument("--bs", type=int, default=4, help="receivers per GPU per step")
    ap.add_argument("--mode", default="train", choices=["train", "infer"],
                    help="train: fwd+bwd (+ gradient all-reduce); infer: no_grad forward only (BASELINE configs[3])")
    ap.add_argument("--overlap", action="store_true",
                    help="per-tensor gradient all-reduces issued inside the backward pass (GradArena.attach) instead of "
                         "one all-reduce of the flat arena after it; measured slower at 8 GPUs (DESIGN 7), off by default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm": p["hbm_gbs"], "tensor": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured"}
    return {"hbm": 6650.0, "tensor": 1400.0, "src": "fallback"}


def synthetic_inputs(render, bs, seed):
    """rays_o, position_tx ~ U(c-h, c+h)^3 with h = half-box - far (all samples in-box)."""
    lo, hi = float(render["xyz_min"]), float(render["xyz_max"])
    c, h = (lo + hi) / 2, (hi - lo) / 2 - float(render["far"])
    if h <= 0:
        c, h = (lo + hi) / 2, 2.0
    g = torch.Generator().manual_seed(seed)
    rx = c + (torch.rand(bs, 3, generator=g) * 2 - 1) * h
    tx = c + (torch.rand(bs, 3, generator=g) * 2 - 1) * h
    ang = torch.rand(bs, generator=g) * 6.283185307179586
    dtx = torch.stack([torch.cos(ang), torch.sin(ang), torch.zeros(bs)], 1)
    return rx.float(), tx.float(), dtx.float()


# ------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline of the native line, and the whole --impl reference arm)
# ------------------------------------------------------------------------------------------------
def cpu_sample_config(cfg, ray_div):
    """A bounded sample of the workload: same samples/ray, IR length and networks, 1/ray_div of the rays."""
    import copy
    c = copy.deepcopy(cfg)
    c["render"]["n_azi"] = max(2, cfg["render"]["n_azi"] // ray_div[0])
    c["render"]["n_ele"] = max(2, cfg["render"]["n_ele"] // ray_div[1])
    return c


def cpu_oracle_step_fn(cfg, bs, seed=0):
    from oracle import field_ref, render_ref
    cls = field_ref.AVRModelRef if cfg["model_class"] == "AVRModel" else field_ref.AVRModelComplexRef
    net = field_ref.trained_like_(cls(cfg["model"]))
    ren = render_ref.RenderRef(net, **cfg["render"])
    rx, tx, dtx = synthetic_inputs(cfg["render"], bs, seed)
    dtx = dtx if cfg["model_class"] != "AVRModel" else None

    def step():
        net.zero_grad(set_to_none=True)
        out = ren(rx, tx, dtx)
        out.square().sum().backward()
        return float(out.detach()[0, 0, 0])
    return step


def time_cpu_oracle(cfg, steps, warmup, ray_div=(4, 2)):
    """-> (IR/s extrapolated to the full ray count, seconds per sample step, description)."""
    torch.set_num_threads(os.cpu_count() or 1)
    full_rays = cfg["render"]["n_azi"] * cfg["render"]["n_ele"] + 2
    sample = cpu_sample_config(cfg, ray_div)
    rays = sample["render"]["n_azi"] * sample["render"]["n_ele"] + 2
    step = cpu_oracle_step_fn(sample, 1)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, steps)
    ir_per_s = 1.0 / (dt * full_rays / rays)
    desc = (f"1 receiver, {rays} of {full_rays} rays (n_azi={sample['render']['n_azi']}, n_ele={sample['render']['n_ele']}), "
            f"S={cfg['render']['n_samples']}, T={cfg['model']['signal_output_dim']}, fwd+bwd through oracle/ (fp32 torch CPU), "
            f"{dt:.2f} s per sample step, scaled linearly in rays")
    return ir_per_s, dt, desc


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    val, dt, desc = time_cpu_oracle(cfg, args.steps, args.warmup, ray_div=(4, 2))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args, 1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(cfg, args, world):
    r = cfg["render"]
    R = r["n_azi"] * r["n_ele"] + 2
    which = {"simu": 1, "raf_furnished": 2, "meshrir": 3, "real_exp_ch_emb_1": 4}.get(args.config, 1)
    what = "fwd+bwd" if args.mode == "train" else "inference (no_grad forward)"
    return {"workload": f"avr_{args.config}.yml render step (BASELINE configs[{which}]): {what}, {args.bs} receivers/GPU",
            "rays": R, "samples_per_ray": r["n_samples"], "ir_len": cfg["model"]["signal_output_dim"],
            "receivers_per_gpu": args.bs, "global_receivers": args.bs * world, "field": cfg["model_class"],
            "parallelism": f"dp{world}", "weights": "random init (hash tables N(0,0.1))",
            "l2": "per-step working set (activations + signal tensor, >10 GB) >> 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.path = None, f"/tmp/avr_clocks_{os.getpid()}.csv"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def run_native(args, cfg):
    import torch.distributed as dist

    import avr_b200
    from avr_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)                        # NCCL prints its version banner on stdout when the communicator is
        os.dup2(2, 1)                            # created: send it to stderr, stdout carries the one JSON line
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _lib.load()

    cls = avr_b200.AVRModel if cfg["model_class"] == "AVRModel" else avr_b200.AVRModel_complex
    field = cls(cfg["model"], seed=1337)                       # same init on every rank (DDP broadcast semantics)
    with torch.no_grad():
        g = torch.Generator().manual_seed(7)
        for m in field.modules():
            if isinstance(m, avr_b200.Encoding):
                m.params.copy_(torch.randn(m.params.shape, generator=g) * 0.1)
    field = field.to(dev)
    ren = avr_b200.AVRRender(field, **cfg["render"], max_receivers_per_pass=args.bs)
    arena = avr_b200.GradArena(ren.parameters())
    overlap = world > 1 and args.overlap
    if overlap:
        arena.attach(ren)
    complex_field = cfg["model_class"] != "AVRModel"

    rx_h, tx_h, dtx_h = synthetic_inputs(cfg["render"], args.bs, 100 + rank)
    rx_h, tx_h, dtx_h = rx_h.pin_memory(), tx_h.pin_memory(), dtx_h.pin_memory()
    rx_d, tx_d, dtx_d = rx_h.to(dev), tx_h.to(dev), dtx_h.to(dev)
    F = cfg["model"]["signal_output_dim"] // 2 + 1
    out_h = torch.empty(args.bs, F, 2).pin_memory()

    def step(host_io: bool):
        arena.zero_()
        if host_io:
            rx, tx = rx_h.to(dev, non_blocking=True), tx_h.to(dev, non_blocking=True)
            dtx = dtx_h.to(dev, non_blocking=True) if complex_field else None
        else:
            rx, tx, dtx = rx_d, tx_d, (dtx_d if complex_field else None)
        if args.mode == "infer":
            with torch.no_grad():
                out = ren(rx, tx, dtx)
        else:
            out = ren(rx, tx, dtx)
            out.square().sum().backward()
            if not overlap:
                arena.all_reduce_mean()
        if host_io:
            out_h.copy_(out.detach(), non_blocking=True)

    def timed(host_io, steps, profile):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.PROFILE = [] if profile else None
        _lib.launch_count(reset=True)
        e0.record()
        for _ in range(steps):
            step(host_io)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        launches = _lib.launch_count()
        prof, ops.PROFILE = ops.PROFILE, None
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, launches, prof

    for _ in range(max(3, args.warmup)):
        step(False)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches, prof = timed(False, args.steps, profile=True)
    clocks = sampler.stop() if sampler else None
    step(True)
    ms_e2e, _, _ = timed(True, args.steps, profile=False)

    total_ir = args.bs * world * args.steps
    value = total_ir / (ms * 1e-3)
    e2e = total_ir / (ms_e2e * 1e-3)

    if rank == 0:
        pk = peaks()
        per = {}
        for name, work, unit, s, e, executed in prof:
            d = per.setdefault(name, {"ms": 0.0, "work": 0.0, "n": 0, "unit": unit, "executed": 0.0})
            d["ms"] += s.elapsed_time(e); d["work"] += work; d["n"] += 1; d["executed"] += executed
        kernels = {}
        for name, d in per.items():
            rate = d["work"] / (d["ms"] * 1e-3) if d["ms"] > 0 else 0.0
            if d["unit"] == "flop":
                kernels[name] = {"bound": "tensor", "achieved": rate / 1e12, "peak": pk["tensor"], "unit": "TFLOP/s",
                                 "frac": rate / 1e12 / pk["tensor"], "launches_per_step": d["n"] / args.steps,
                                 "ms_per_step": d["ms"] / args.steps, "share_of_step": d["ms"] / ms,
                                 "tensor_pipe_tflops": d["executed"] / (d["ms"] * 1e-3) / 1e12,
                                 "tensor_pipe_frac": d["executed"] / (d["ms"] * 1e-3) / 1e12 / pk["tensor"]}
            else:
                kernels[name] = {"bound": "hbm", "achieved": rate / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                 "frac": rate / 1e9 / pk["hbm"], "launches_per_step": d["n"] / args.steps,
                                 "ms_per_step": d["ms"] / args.steps, "share_of_step": d["ms"] / ms}
        if os.environ.get("AVR_BENCH_DETAIL"):
            per_step = len(prof) // args.steps
            for name, work, unit, s0, e0, _ex in prof[-per_step:]:
                t = s0.elapsed_time(e0)
                rate = work / (t * 1e-3) / (1e12 if unit == "flop" else 1e9)
                sys.stderr.write(f"DETAIL {name:22s} {t:8.3f} ms  work {work:.3e} {unit}  {rate:8.1f} {'TFLOP/s' if unit == 'flop' else 'GB/s'}\n")
        dominant = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
        roof = dict(kernels[dominant]) if dominant else {}
        traffic, traffic_src = None, None
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "umma_traffic.json")
        if dominant == "umma_gemm" and args.config == "simu" and args.bs == 4 and os.path.exists(tpath):
            # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture of this very
            # command (profilers are never run inside a timed bench); null for any other workload
            with open(tpath) as fh:
                tj = json.load(fh)
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        roof.update({"kernel": dominant, "traffic": traffic, "traffic_unit": "DRAM bytes per launch", "traffic_source": traffic_src,
                     "peak_source": pk["src"],
                     "note": "achieved = algorithmic flops (2MNK of the fp32-grade product) or bytes (SURVEY 8d) of the timed "
                             "launches / their CUDA-event time inside the timed region; each algorithmic product costs 3 "
                             "(backward bf16 pairs, forward fp16 pairs) or 6 (forward bf16 triples) tcgen05 products, see tensor_pipe_*"})
        cpu = None
        if not args.no_cpu_baseline:
            v, dt, desc = time_cpu_oracle(cfg, steps=1, warmup=1, ray_div=(4, 2))
            cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": desc}
        line = {
            "metric": METRIC if args.mode == "train" else "rendered IRs/sec (inference)", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(cfg, args, world), "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(rx_h.numel() * 4 * (3 if complex_field else 2) +
                                              (cfg["render"]["n_azi"] * cfg["render"]["n_ele"] + 2) * 12),
                    "d2h_bytes_per_step": int(out_h.numel() * 4)},
            "gpu_launches": int(launches), "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
            "grad_allreduce_bytes": int(arena.numel() * 4) if world > 1 and args.mode == "train" else 0,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    from avr_b200.configs import get_config
    cfg = get_config(args.config)
    if args.impl == "reference":
        run_reference(args, cfg)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device; there is no CPU fallback. "
                         "Use --impl reference for the CPU oracle timing.")
    run_native(args, cfg)


if __name__ == "__main__":
    main()
