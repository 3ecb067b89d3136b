#!/usr/bin/env python
"""Headline benchmark: rendered IRs/sec (forward + backward) of the AVR render hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config simu] [--bs 4]
                    [--mode train|infer] [--receivers M] [--no-other-configs] [--no-cpu-baseline]

A *step* is one forward + backward pass of ``AVRRender`` (ray generation, hash-grid encode, MLPs,
compositing, spectrum; loss = sum of squares of the ``[bs,F,2]`` IR spectrum) over one synthetic batch of
``bs`` receivers per GPU at the ``avr_simu.yml`` shape (BASELINE.json configs[1]), plus -- for N > 1 -- the
gradient exchange of the data-parallel step (weak scaling: ``bs`` receivers per GPU).  No optimizer
update is part of the metric ("receivers*steps / time for forward+backward").

What the JSON line holds, and how each number was taken:

* ``value``       K steps, inputs resident in HBM, NO per-kernel events, one CUDA-event pair around the K steps, max over ranks.
                  Timed after the W warm-up steps plus ~1.5 s of further untimed steps (``warmup_note``): the power-capped
                  GPUs run the first tenths of a second after idle 2-3 % faster than the steady state.
* ``e2e``         the same K steps through ``AVRRender.forward`` with pinned HOST inputs: every step copies its positions
                  (on a copy stream, one step ahead: two device buffers; the per-step direction table inside forward) and
                  copies its spectra back to a pinned buffer right after the forward pass; the steps are not synchronised
                  one by one (throughput, not latency), the region ends with a full device sync.
* ``kernels`` / ``roofline``   a THIRD pass with CUDA events around every library call (``ops.PROFILE``).
* ``grid_grad_atomic``         the same workload with fp32-atomic table gradients instead of the deterministic default.
* ``other_configs``            the other BASELINE shapes (configs[2..4]) measured in this very run, 5 steps each, through the
                               same code path and -- under torchrun -- with the same gradient exchange.
* ``cpu_baseline`` / ``--impl reference``   the CPU oracle (``oracle/``: restatement of renderer_cpu.py + the tcnn field in
                  fp32 torch; the reference's own runner needs tiny-cuda-nn) on the host cores: ONE receiver per step at
                  the FULL ray count, measured seconds, nothing extrapolated; ``config`` says so.
* ``grad_checksums``           (N > 1) float64 sum of every rank's gradient arena after the exchange; the run fails unless
                  they are equal.
"""
from __future__ import annotations

import argparse
import gc
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
WHICH = {"simu": 1, "raf_furnished": 2, "meshrir": 3, "real_exp_ch_emb_1": 4}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="simu")
    ap.add_argument("--bs", type=int, default=4, help="receivers per GPU per step")
    ap.add_argument("--mode", default="train", choices=["train", "infer"],
                    help="train: fwd+bwd (+ gradient exchange); infer: no_grad forward only (BASELINE configs[3])")
    ap.add_argument("--receivers", type=int, default=0,
                    help="infer mode: render this many receivers in total (e.g. 3969 = MeshRIR S1-M3969), sharded over the "
                         "ranks in DistributedSampler order, --bs receivers per pass; the result is all-gathered")
    ap.add_argument("--grid-grad", default="deterministic", choices=["deterministic", "atomic"])
    ap.add_argument("--flat-allreduce", action="store_true",
                    help="N > 1: all-reduce the whole gradient arena (round-1 behaviour) instead of exchanging the per-ray / "
                         "per-receiver table gradients as rows (GradArena.attach)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the pass with the other hash-table-gradient mode (profiling runs)")
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


def meshrir_grid(n):
    """receivers on a regular 63 x 63 grid over [-0.5, 0.5]^2 at z = 0, one transmitter at (2, 0, 0)."""
    side = int(round(n ** 0.5))
    if side * side == n:
        ax = torch.linspace(-0.5, 0.5, side)
        gx, gy = torch.meshgrid(ax, ax, indexing="ij")
        rx = torch.stack([gx.reshape(-1), gy.reshape(-1), torch.zeros(n)], 1)
    else:
        rx = torch.cat([torch.rand(n, 2, generator=torch.Generator().manual_seed(3)) - 0.5, torch.zeros(n, 1)], 1)
    tx = torch.tensor([[2.0, 0.0, 0.0]]).expand(n, 3).contiguous()
    return rx.float(), tx.float()


# ------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline of the native line, and the whole --impl reference arm)
# ------------------------------------------------------------------------------------------------
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


def time_cpu_oracle(cfg, steps, warmup, budget_s):
    """ONE receiver per step at the FULL ray count through the oracle, fwd+bwd, all host threads.
    -> (IR/s = 1 / seconds per step, seconds per step, steps actually timed, description).  Nothing is extrapolated;
    the run stops early once ``budget_s`` seconds of timed work are spent (steps are 5-15 s each)."""
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_oracle_step_fn(cfg, 1)
    for _ in range(min(warmup, 1)):                          # one full-size warm-up pays first-touch of the ~8 GB working set
        step()
    done, t0 = 0, time.perf_counter()
    while done < max(1, steps):
        step()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    r = cfg["render"]
    rays = r["n_azi"] * r["n_ele"] + 2
    desc = (f"1 receiver per step at the full ray count ({rays} rays x {r['n_samples']} samples, T={cfg['model']['signal_output_dim']}), "
            f"fwd+bwd through oracle/ (fp32 torch CPU, {torch.get_num_threads()} threads), {done} timed step(s) of {dt:.2f} s, "
            f"measured (no extrapolation)")
    return 1.0 / dt, dt, done, desc


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    val, dt, done, desc = time_cpu_oracle(cfg, args.steps, args.warmup, budget_s=150.0)
    ns = argparse.Namespace(**vars(args))
    ns.bs, ns.mode = 1, "train"
    config = workload_config(cfg, ns, 1)
    config["sample"] = "the CPU arm renders 1 receiver per step (the GPU arm's step is %d receivers per GPU); rays, samples, IR length and networks are the full-size ones" % args.bs
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(cfg, args, world):
    r = cfg["render"]
    R = r["n_azi"] * r["n_ele"] + 2
    which = WHICH.get(args.config, 1)
    what = "fwd+bwd" if args.mode == "train" else "inference (no_grad forward)"
    return {"workload": f"avr_{args.config}.yml render step (BASELINE configs[{which}]): {what}, {args.bs} receivers/GPU",
            "rays": R, "samples_per_ray": r["n_samples"], "ir_len": cfg["model"]["signal_output_dim"],
            "receivers_per_gpu": args.bs, "global_receivers": args.bs * world, "field": cfg["model_class"],
            "parallelism": f"dp{world}", "weights": "random init (hash tables N(0,0.1))",
            "grid_grad": getattr(args, "grid_grad", "deterministic"),
            "l2": "per-step working set (activations, >10 GB) >> 126 MB L2; no explicit flush"}


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
                                          "--format=csv,noheader,nounits", "-lms", "50"],
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
class Workload:
    """One (config, batch, mode) on this rank's GPU: field, renderer, gradient arena, synthetic host + device inputs."""

    def __init__(self, name, bs, mode, grid_grad, dev, rank, world, flat_allreduce=False):
        import avr_b200
        from avr_b200.configs import get_config
        self.name, self.bs, self.mode, self.dev, self.world = name, bs, mode, dev, world
        self.cfg = cfg = get_config(name)
        cls = avr_b200.AVRModel if cfg["model_class"] == "AVRModel" else avr_b200.AVRModel_complex
        field = cls(cfg["model"], seed=1337)                       # same init on every rank (DDP broadcast semantics)
        with torch.no_grad():
            g = torch.Generator().manual_seed(7)
            for m in field.modules():
                if isinstance(m, avr_b200.Encoding):
                    m.params.copy_(torch.randn(m.params.shape, generator=g) * 0.1)
        self.field = field.to(dev)
        self.ren = avr_b200.AVRRender(self.field, **cfg["render"], max_receivers_per_pass=bs, grid_grad=grid_grad)
        self.arena = avr_b200.GradArena(self.ren.parameters())
        if world > 1 and mode == "train" and not flat_allreduce:
            self.arena.attach(self.ren)
        self.complex_field = cfg["model_class"] != "AVRModel"
        rx_h, tx_h, dtx_h = synthetic_inputs(cfg["render"], bs, 100 + rank)
        self.host = (rx_h.pin_memory(), tx_h.pin_memory(), dtx_h.pin_memory())
        self.device = tuple(t.to(dev) for t in self.host)
        self.ch = (torch.arange(bs, device=dev) % 8) if name == "real_exp_ch_emb_1" else None    # one 8-mic array
        self.F = cfg["model"]["signal_output_dim"] // 2 + 1
        self.out_h = torch.empty(bs, self.F, 2).pin_memory()
        # end-to-end steps: the inputs of step k+1 are copied from pinned host memory on a copy stream while step k computes
        # (two device buffers), and the spectra leave on the copy stream as soon as the forward pass has produced them --
        # what a training loop with a prefetching loader does; every step still copies its own inputs and its own result
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.dbuf = [tuple(torch.empty_like(t) for t in self.device) for _ in range(2)]
        self.h2d_done = [None, None]
        self.read_done = [None, None]
        self.k = 0

    def _prefetch(self, i):
        cs = self.copy_stream
        if self.read_done[i] is not None:
            cs.wait_event(self.read_done[i])                   # the step that last read buffer i has been through it
        with torch.cuda.stream(cs):
            for dst, src in zip(self.dbuf[i], self.host):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        self.h2d_done[i] = ev

    def step(self, host_io: bool):
        self.arena.zero_()
        cur = torch.cuda.current_stream()
        if host_io:
            i = self.k & 1
            if self.h2d_done[i] is None:
                self._prefetch(i)                              # first step: nothing was prefetched yet
            cur.wait_event(self.h2d_done[i])
            self.h2d_done[i] = None
            rx, tx, dtx = self.dbuf[i][0], self.dbuf[i][1], (self.dbuf[i][2] if self.complex_field else None)
            self._prefetch(i ^ 1)                              # next step's inputs, concurrently with this step
        else:
            rx, tx, dtx = self.device[0], self.device[1], (self.device[2] if self.complex_field else None)
        if self.mode == "infer":
            with torch.no_grad():
                out = self.ren(rx, tx, dtx, ch_idx=self.ch)
        else:
            out = self.ren(rx, tx, dtx, ch_idx=self.ch)
        if host_io:
            fwd = torch.cuda.Event()
            fwd.record(cur)
            self.copy_stream.wait_event(fwd)
            with torch.cuda.stream(self.copy_stream):
                self.out_h.copy_(out.detach(), non_blocking=True)
            out.record_stream(self.copy_stream)
        if self.mode != "infer":
            out.square().sum().backward()
            self.arena.all_reduce_mean()
        if host_io:
            ev = torch.cuda.Event()
            ev.record(cur)
            self.read_done[self.k & 1] = ev
            self.k += 1

    def timed(self, host_io, steps, profile=False):
        import torch.distributed as dist
        from avr_b200 import _lib, ops
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.PROFILE = [] if profile else None
        _lib.launch_count(reset=True)
        e0.record()
        for _ in range(steps):
            self.step(host_io)
        e1.record()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        launches = _lib.launch_count()
        prof, ops.PROFILE = ops.PROFILE, None
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, launches, prof

    def h2d_bytes(self):
        r = self.cfg["render"]
        return int(self.host[0].numel() * 4 * (3 if self.complex_field else 2) + (r["n_azi"] * r["n_ele"] + 2) * 12)

    def close(self):
        self.ren = self.field = self.arena = None
        gc.collect()
        torch.cuda.empty_cache()


def settle_steps(w, seconds, max_steps):
    """Untimed steps until ``seconds`` of device time have passed (the same count on every rank: decided from the first
    steps' duration on rank 0's clock would need a broadcast, so it is derived from the step time measured here and a
    MAX over the ranks)."""
    import torch.distributed as dist
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        w.step(False)
    e1.record()
    torch.cuda.synchronize()
    per = torch.tensor([e0.elapsed_time(e1) / 3e3], device=w.dev)
    if w.world > 1:
        dist.all_reduce(per, op=dist.ReduceOp.MAX)
    n = int(min(max_steps, max(0.0, seconds / max(float(per), 1e-4) - 3)))
    for _ in range(n):
        w.step(False)
    torch.cuda.synchronize()
    return n + 3


def kernel_table(prof, steps, ms_total, pk):
    per = {}
    for name, work, unit, s, e, executed in prof:
        d = per.setdefault(name, {"ms": 0.0, "work": 0.0, "n": 0, "unit": unit, "executed": 0.0})
        d["ms"] += s.elapsed_time(e); d["work"] += work; d["n"] += 1; d["executed"] += executed
    kernels = {}
    for name, d in per.items():
        rate = d["work"] / (d["ms"] * 1e-3) if d["ms"] > 0 else 0.0
        if d["unit"] == "flop":
            kernels[name] = {"bound": "tensor", "achieved": rate / 1e12, "peak": pk["tensor"], "unit": "TFLOP/s",
                             "frac": rate / 1e12 / pk["tensor"], "launches_per_step": d["n"] / steps,
                             "ms_per_step": d["ms"] / steps, "share_of_step": d["ms"] / ms_total,
                             "tensor_pipe_tflops": d["executed"] / (d["ms"] * 1e-3) / 1e12,
                             "tensor_pipe_frac": d["executed"] / (d["ms"] * 1e-3) / 1e12 / pk["tensor"]}
        else:
            kernels[name] = {"bound": "hbm", "achieved": rate / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                             "frac": rate / 1e9 / pk["hbm"], "launches_per_step": d["n"] / steps,
                             "ms_per_step": d["ms"] / steps, "share_of_step": d["ms"] / ms_total}
    return kernels


def run_sharded_inference(args, cfg, dev, rank, world):
    """BASELINE configs[3]: render ALL receivers of a scene once (MeshRIR S1-M3969: 3969 positions), sharded over the ranks
    in DistributedSampler order (tools/README.md:3-5, eval_rotate_doa_avr.py:104-110), no collective on the data path; the
    ``[n_local, F, 2]`` spectra are all-gathered at the end."""
    import torch.distributed as dist

    import avr_b200
    from avr_b200 import _lib
    from avr_b200.ddp import shard_receivers
    w = Workload(args.config, args.bs, "infer", args.grid_grad, dev, rank, world)
    n = args.receivers
    rx_all, tx_all = meshrir_grid(n)
    mine = shard_receivers(n, rank, world)
    rx_h, tx_h = rx_all[mine].pin_memory(), tx_all[mine].pin_memory()
    free, _ = torch.cuda.mem_get_info(dev)
    w.ren.max_receivers_per_pass = args.bs
    out_h = torch.empty(len(mine), w.F, 2).pin_memory()

    def render_all():
        with torch.no_grad():
            out = w.ren(rx_h.to(dev, non_blocking=True), tx_h.to(dev, non_blocking=True))
            if world > 1:
                gathered = torch.empty(world * out.shape[0], *out.shape[1:], device=dev)
                dist.all_gather_into_tensor(gathered, out.contiguous())
            out_h.copy_(out, non_blocking=True)
        return out

    with torch.no_grad():                                        # warm-up: one pass of --bs receivers
        w.ren(rx_h[:args.bs].to(dev), tx_h[:args.bs].to(dev))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index) if rank == 0 else None
    _lib.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = render_all()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        ns = argparse.Namespace(**vars(args))
        config = workload_config(cfg, ns, world)
        config.update({"workload": f"avr_{args.config}.yml: inference render of all {n} receivers of a scene, sharded over "
                                   f"{world} GPU(s) in DistributedSampler order, {args.bs} receivers per pass",
                       "global_receivers": n, "receivers_per_gpu": len(mine), "free_hbm_gb_before": free / 1e9})
        line = {"metric": "rendered IRs/sec (inference, all receivers of a scene)", "value": n / (ms * 1e-3), "unit": UNIT,
                "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": n / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(rx_h.numel() * 8),
                        "d2h_bytes_per_step": int(out_h.numel() * 4),
                        "note": "host positions in, spectra out to pinned host memory: this mode IS the end-to-end call"},
                "gpu_launches": int(launches), "finite": bool(torch.isfinite(out).all())}
        print(json.dumps(line), flush=True)


def run_native(args, cfg):
    import torch.distributed as dist

    from avr_b200 import _lib

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
    if args.mode == "infer" and args.receivers > 0:
        run_sharded_inference(args, cfg, dev, rank, world)
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    warm = max(3, args.warmup)
    w = Workload(args.config, args.bs, args.mode, args.grid_grad, dev, rank, world, args.flat_allreduce)
    for _ in range(warm):
        w.step(False)
    # ... and on until the GPU has been under this load for ~1.5 s: the boxes are power-capped, and the first ~0.3 s after an
    # idle period run 2-3 % faster than the steady state every later pass sees (profiles/r2/value_vs_e2e_order.txt: 11.60 ms
    # for the first 20 steps, 11.9 ms for every pass after it, device-resident or end-to-end alike)
    settle = settle_steps(w, 1.5, 200)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches, _ = w.timed(False, args.steps)                         # headline: no per-kernel events
    w.step(True)
    ms_e2e, _, _ = w.timed(True, args.steps)
    prof_steps = min(args.steps, 5)
    ms_prof, _, prof = w.timed(False, prof_steps, profile=True)          # third pass: CUDA events around every library call
    checksums = None
    if world > 1 and args.mode == "train":
        cs = torch.tensor([w.arena.checksum()], dtype=torch.float64, device=dev)
        allc = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allc, cs)
        checksums = [float(c) for c in allc]
        if len(set(checksums)) != 1:
            raise SystemExit(f"gradient arenas differ across ranks after the exchange: {checksums}")
    grad_bytes = int(w.arena.reduce_numel * 4) if world > 1 and args.mode == "train" else 0
    arena_bytes = int(w.arena.numel() * 4)
    h2d, d2h = w.h2d_bytes(), int(w.out_h.numel() * 4)
    w.close()

    total_ir = args.bs * world * args.steps
    value = total_ir / (ms * 1e-3)
    e2e = total_ir / (ms_e2e * 1e-3)

    # the same workload with the other accumulation mode of the hash-table gradients
    other_mode = "atomic" if args.grid_grad == "deterministic" else "deterministic"
    alt = None
    if args.mode == "train" and not args.no_alt:
        w2 = Workload(args.config, args.bs, args.mode, other_mode, dev, rank, world, args.flat_allreduce)
        for _ in range(warm):
            w2.step(False)
        ms2, _, _ = w2.timed(False, args.steps)
        alt = {"grid_grad": other_mode, "value": total_ir / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / args.steps}
        w2.close()

    others = {}
    if not args.no_other_configs and args.config == "simu" and args.mode == "train":
        plan = [("raf_furnished", 4, "train"), ("real_exp_ch_emb_1", 8, "train"), ("meshrir", 4, "train"), ("meshrir", 8, "infer")]
        for name, bs, mode in plan:
            wo = Workload(name, bs, mode, args.grid_grad, dev, rank, world, args.flat_allreduce)
            for _ in range(3):
                wo.step(False)
            mso, lo, _ = wo.timed(False, 5)
            mse, _, _ = wo.timed(True, 5)
            msp, _, profo = wo.timed(False, 2, profile=True)
            if rank == 0:
                ko = kernel_table(profo, 2, msp, pk)
                dom = max(ko, key=lambda k: ko[k]["ms_per_step"]) if ko else None
                ns = argparse.Namespace(**vars(args)); ns.config, ns.bs, ns.mode = name, bs, mode
                others[f"{name}:{mode}"] = {
                    "metric": METRIC if mode == "train" else "rendered IRs/sec (inference)",
                    "value": bs * world * 5 / (mso * 1e-3), "unit": UNIT, "ms_per_step": mso / 5, "steps": 5, "warmup": 3,
                    "e2e": {"value": bs * world * 5 / (mse * 1e-3), "unit": UNIT}, "gpu_launches": int(lo),
                    "config": workload_config(wo.cfg, ns, world),
                    "roofline": dict(ko[dom], kernel=dom) if dom else None,
                    "grad_exchange_bytes": int(wo.arena.reduce_numel * 4) if world > 1 and mode == "train" else 0}
            wo.close()
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        kernels = kernel_table(prof, prof_steps, ms_prof, pk)
        if os.environ.get("AVR_BENCH_DETAIL"):
            per_step = len(prof) // prof_steps
            for name, work, unit, s0, e0, _ex in prof[-per_step:]:
                t = s0.elapsed_time(e0)
                rate = work / (t * 1e-3) / (1e12 if unit == "flop" else 1e9)
                sys.stderr.write(f"DETAIL {name:22s} {t:8.3f} ms  work {work:.3e} {unit}  {rate:8.1f} {'TFLOP/s' if unit == 'flop' else 'GB/s'}\n")
        dominant = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
        roof = dict(kernels[dominant]) if dominant else {}
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "umma_traffic.json")
        if dominant == "umma_gemm" and args.config == "simu" and args.bs == 4 and os.path.exists(tpath):
            # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture of this very
            # command (profilers are never run inside a timed bench); null for any other workload
            with open(tpath) as fh:
                tj = json.load(fh)
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        roof.update({"kernel": dominant, "traffic": traffic, "traffic_unit": "DRAM bytes per launch", "traffic_source": traffic_src,
                     "peak_source": pk["src"],
                     "note": "achieved = algorithmic flops (2MNK of the fp32-grade product) or bytes (SURVEY 8d) of the launches "
                             "of the profiling pass / their CUDA-event time; each algorithmic product costs 3 (backward bf16 "
                             "pairs, forward fp16 pairs) or 6 (forward bf16 triples) tcgen05 products, see tensor_pipe_*"})
        cpu = None
        if not args.no_cpu_baseline:
            v, dt, done, desc = time_cpu_oracle(cfg, steps=1, warmup=0, budget_s=30.0)
            cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": desc}
        line = {
            "metric": METRIC if args.mode == "train" else "rendered IRs/sec (inference)", "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(cfg, args, world), "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "every step copies its inputs from pinned host memory (copy stream, one step ahead of the compute stream, "
                            "two device buffers) and its spectra back to pinned host memory (copy stream, right after the forward "
                            "pass); steps are not synchronised one by one (throughput), the timed region ends with a device "
                            "synchronise over all streams"},
            "warmup_note": f"{warm} warm-up steps as requested + {settle} more untimed steps (~1.5 s under load) so that the timed "
                           "passes see the power-capped steady state, not the first tenths of a second after idle",
            "gpu_launches": int(launches), "roofline": roof, "kernels": kernels,
            "profile_pass": {"steps": prof_steps, "ms_per_step": ms_prof / prof_steps,
                             "note": "kernels / roofline come from this separate pass with CUDA events around every library call"},
            "grid_grad_alt": alt, "other_configs": others, "cpu_baseline": cpu,
            "grad_exchange": {"allreduce_bytes": grad_bytes, "arena_bytes": arena_bytes,
                              "row_exchange": bool(world > 1 and args.mode == "train" and not args.flat_allreduce)},
            "grad_checksums": checksums,
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
