"""Tensor-level wrappers over the C-ABI (no autograd here; see ``functional.py`` / ``fused.py``).

PyTorch is used for device memory and streams only.  Every function launches hand-written kernels
from ``libavr_b200.so`` on the current CUDA stream of the tensors' device and fails loudly otherwise.

Buffers come in two formats: plain fp32 tensors, and ``PlanePair`` -- the error-compensated bf16
hi/lo pair that feeds the tcgen05 tensor-core GEMMs (``include/avr_b200.h``, "dense layers on the
tensor cores").  Kernels that produce or consume activations accept either.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import (GEMM_ACCUM, GEMM_MASK, GEMM_RELU, GEMM_RELU_A, GEMM_RELU_B, I_CONTIG, K_CONTIG,  # noqa: F401
                   UMMA_ACCUM, UMMA_BIAS, UMMA_BITS, UMMA_DUAL_COPY, UMMA_DUAL_RELU, UMMA_MASK, UMMA_OUT_F32, UMMA_RELU, GridMeta,
                   RenderGeom)

# ---- optional per-launch timing (bench.py): CUDA events on the launching stream ----------------------
PROFILE = None          # set to a list to collect (name, work, unit, start_event, end_event)


class _timed:
    __slots__ = ("name", "work", "unit", "start", "executed")

    def __init__(self, name, work, unit, executed=None):
        self.name, self.work, self.unit, self.start = name, work, unit, None
        self.executed = work if executed is None else executed      # e.g. tensor-pipe flops incl. the 3 / 6 products

    def __enter__(self):
        if PROFILE is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()
        return self

    def __exit__(self, *exc):
        if self.start is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            PROFILE.append((self.name, self.work, self.unit, self.start, end, self.executed))
        return False


PLANES_BF16x2, PLANES_BF16x3, PLANES_F16x2 = 2, 3, 18       # include/avr_b200.h, plane-set kinds


class PlanePair:
    """A plane set: a 16-bit buffer ``[n, rows, ld]``, optionally a row / column window of it.

    bf16 buffer: ``x = hi + mid (+ lo)`` (n = 2: 16 mantissa bits, n = 3: 24, fp32 range);
    fp16 buffer (n = 2): ``x = hi + lo' * 2^-11`` with ``lo' = fp16((x - hi) * 2^11)`` -- 24 bits in two planes within
    fp16's range, so a product needs three tensor-core MMAs where two bf16 triples need six."""

    __slots__ = ("buf", "ld", "row0", "rows", "col0", "cols")

    def __init__(self, buf, col0=0, cols=None, row0=0, rows=None):
        assert buf.dim() == 3 and buf.is_contiguous()
        assert (buf.dtype == torch.bfloat16 and buf.shape[0] in (2, 3)) or (buf.dtype == torch.float16 and buf.shape[0] == 2)
        self.buf, self.ld = buf, buf.shape[2]
        self.row0 = row0
        self.rows = buf.shape[1] - row0 if rows is None else rows
        self.col0 = col0
        self.cols = self.ld - col0 if cols is None else cols
        assert col0 % 8 == 0 and self.ld % 8 == 0, "plane windows must stay 16-byte aligned"

    @staticmethod
    def empty(rows, cols, device, ld=None, n=2, kind=None):
        """``kind`` (PLANES_*) overrides ``n``: PLANES_F16x2 allocates an fp16 pair."""
        ld = cols if ld is None else ld
        ld = (ld + 7) // 8 * 8
        if kind == PLANES_F16x2:
            return PlanePair(torch.empty(2, rows, ld, dtype=torch.float16, device=device), 0, cols)
        n = n if kind is None else kind
        return PlanePair(torch.empty(n, rows, ld, dtype=torch.bfloat16, device=device), 0, cols)

    @staticmethod
    def zeros(rows, cols, device, ld=None, n=2, kind=None):
        pp = PlanePair.empty(rows, cols, device, ld, n, kind)
        pp.buf.zero_()
        return pp

    @property
    def n(self):
        return self.buf.shape[0]

    @property
    def f16(self):
        return self.buf.dtype == torch.float16

    @property
    def kind(self):
        return PLANES_F16x2 if self.f16 else self.buf.shape[0]

    def window(self, col0, cols):
        return PlanePair(self.buf, self.col0 + col0, cols, self.row0, self.rows)

    def row_window(self, row0, rows):
        return PlanePair(self.buf, self.col0, self.cols, self.row0 + row0, rows)

    @property
    def ptr(self):
        return C.c_void_p(self.buf.data_ptr() + 2 * (self.row0 * self.ld + self.col0))

    @property
    def plane(self):
        return self.buf.shape[1] * self.ld

    @property
    def device(self):
        return self.buf.device


def _ctx(t):
    t = t.buf if isinstance(t, PlanePair) else t
    if not t.is_cuda:
        raise _lib.AVRLibraryError("avr_b200 ops need CUDA tensors (there is no CPU fallback)")
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    return dev, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _p(t, dtype=torch.float32):
    if t is None:
        return None
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_cuda:
        raise _lib.AVRLibraryError("expected a CUDA tensor")
    return C.c_void_p(t.data_ptr())


def _dense(t):
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t


def _np(x):
    return x.kind if isinstance(x, PlanePair) else 0


def _mat(x):
    """-> (pointer, leading dimension, plane stride [0 = fp32]) of a 2-D fp32 tensor / view or a PlanePair."""
    if isinstance(x, PlanePair):
        return x.ptr, x.ld, x.plane
    if x.dtype != torch.float32 or not x.is_cuda or x.stride(-1) != 1:
        raise TypeError("expected a CUDA fp32 tensor with unit inner stride")
    return C.c_void_p(x.data_ptr()), x.stride(0), 0


def _rows(x):
    return x.rows if isinstance(x, PlanePair) else x.shape[0]


def make_geom(render_cfg: dict, bs: int, T: int) -> RenderGeom:
    """``render_cfg`` as in the reference YAML ``render:`` section (renderer.py:20-29)."""
    R = int(render_cfg["n_azi"]) * int(render_cfg["n_ele"]) + 2
    lo, hi = render_cfg["xyz_min"], render_cfg["xyz_max"]
    return RenderGeom(int(bs), R, int(render_cfg["n_samples"]), int(T), float(lo), float(float(hi) - float(lo)),
                      float(render_cfg["fs"]), float(render_cfg["speed"]))


def make_grid_meta(geom: dict) -> GridMeta:
    """``geom`` from ``hashgrid_geometry`` (model.py)."""
    m = GridMeta()
    m.n_levels, m.n_feat, m.total = geom["n_levels"], geom["n_feat"], geom["total"]
    m.stride32 = 1 if geom.get("index_stride", "uint32") == "uint32" else 0
    for l in range(geom["n_levels"]):
        m.scale[l], m.res[l], m.size[l], m.offset[l] = geom["scale"][l], geom["res"][l], geom["size"][l], geom["offset"][l]
    return m


def headroom_bits(n_points: int) -> int:
    """log2 of the largest number of contributions one table entry can receive (8 corners/point)."""
    return max(4, int(math.ceil(math.log2(max(1, n_points) * 8))) + 1)


# ---- geometry -------------------------------------------------------------------------------------
def sample_points(g: RenderGeom, rays_o, pos_tx, dirs, d_vals, want_pts=True, want_delay=True):
    dev, st = _ctx(rays_o)
    pts = torch.empty(g.bs, g.R * g.S, 3, device=rays_o.device) if want_pts else None
    view = torch.empty_like(pts) if want_pts else None
    txn = torch.empty_like(pts) if want_pts else None
    delay = torch.empty(g.bs, g.R, g.S, dtype=torch.int32, device=rays_o.device) if want_delay else None
    _lib.check(_lib.load().avr_sample_points(C.byref(g), _p(_dense(rays_o)), _p(_dense(pos_tx)), _p(_dense(dirs)),
                                             _p(_dense(d_vals)), _p(pts), _p(view), _p(txn), _p(delay, torch.int32),
                                             dev, st), "avr_sample_points")
    return pts, view, txn, delay


def aux_inputs(g: RenderGeom, pos_tx, dirs, dir_tx=None):
    dev, st = _ctx(pos_tx)
    u_view = torch.empty(g.R, 3, device=pos_tx.device)
    u_tx = torch.empty(g.bs, 3, device=pos_tx.device)
    u_dtx = torch.empty(g.bs, 3, device=pos_tx.device) if dir_tx is not None else None
    _lib.check(_lib.load().avr_aux_inputs(C.byref(g), _p(_dense(pos_tx)), _p(_dense(dirs)),
                                          _p(_dense(dir_tx)) if dir_tx is not None else None,
                                          _p(u_view), _p(u_tx), _p(u_dtx), dev, st), "avr_aux_inputs")
    return u_view, u_tx, u_dtx


# ---- hash grid ------------------------------------------------------------------------------------
def raygen_encode_fwd(g, meta, rays_o, pos_tx, dirs, d_vals, table, out, col0=0, n_ones=0, delay=None):
    """``out``: fp32 ``[N, ld]`` tensor or PlanePair; encoded columns land at ``col0``."""
    dev, st = _ctx(out)
    ptr, ld, plane = _mat(out)
    n_pts = g.bs * g.R * g.S
    with _timed("raygen_encode_fwd", float(n_pts) * meta.n_levels * 72, "byte"):     # 8 corners*8 B + 8 B out
        _lib.check(_lib.load().avr_raygen_encode_fwd(C.byref(g), C.byref(meta), _p(_dense(rays_o)),
                                                     _p(_dense(pos_tx)) if pos_tx is not None else None,
                                                     _p(_dense(dirs)), _p(_dense(d_vals)), _p(_dense(table)), ptr, ld,
                                                     plane, _np(out), col0, n_ones, _p(delay, torch.int32), dev, st),
                   "avr_raygen_encode_fwd")


def grid_encode_fwd(meta, u, table, out, col0=0, n_ones=0):
    dev, st = _ctx(out)
    ptr, ld, plane = _mat(out)
    _lib.check(_lib.load().avr_grid_encode_fwd(C.byref(meta), _p(_dense(u)), u.shape[0], _p(_dense(table)), ptr, ld, plane,
                                               _np(out), col0, n_ones, dev, st), "avr_grid_encode_fwd")


GRID_GRAD_F32 = -1       # include/avr_b200.h AVR_GRID_GRAD_F32
GRID_GRAD_MODES = ("atomic", "deterministic")


class GridGradAccumulator:
    """Accumulation of hash-table gradients (see include/avr_b200.h).

    ``mode="deterministic"``: 2^e-scaled int64 fixed point (bit-reproducible, two reductions per cell corner);
    ``mode="atomic"``: fp32 vector reductions straight into the gradient (one per corner; the summation order
    varies from run to run, as with tcnn's own atomics).
    """

    def __init__(self, meta: GridMeta, device, n_points: int, scratch=None, mode: str = "deterministic"):
        if mode not in GRID_GRAD_MODES:
            raise ValueError(f"grid_grad must be one of {GRID_GRAD_MODES}")
        self.meta, self.n, self.mode = meta, int(meta.total) * 2, mode
        if mode == "atomic":
            self.headroom = GRID_GRAD_F32
            self.acc = torch.zeros(self.n, dtype=torch.float32, device=device)
            self.gmax = None
            return
        self.headroom = headroom_bits(n_points)
        if scratch is not None and scratch.numel() >= self.n:
            self.acc = scratch[: self.n]
        else:
            self.acc = torch.empty(self.n, dtype=torch.int64, device=device)
        self.acc.zero_()
        self.gmax = torch.zeros(1, dtype=torch.int32, device=device)

    def observe(self, d_out, col0, ncols):
        if self.mode == "atomic":
            return
        dev, st = _ctx(d_out)
        ptr, ld, plane = _mat(d_out)
        _lib.check(_lib.load().avr_absmax_bits(ptr, _rows(d_out), ld, plane, col0, ncols, _p(self.gmax, torch.int32),
                                               dev, st), "avr_absmax_bits")

    def add_rays(self, g, rays_o, dirs, d_vals, d_out, col0=0, sample_step=0.0):
        """``sample_step``: spacing of a ray's samples in unit-cube coordinates (0: unknown), see avr_raygen_encode_bwd."""
        dev, st = _ctx(d_out)
        ptr, ld, plane = _mat(d_out)
        n_pts = g.bs * g.R * g.S
        rmw = 8 if self.mode == "atomic" else 16                   # bytes read-modify-written per cell corner
        with _timed("raygen_encode_bwd", float(n_pts) * self.meta.n_levels * (8 + 8 * rmw), "byte"):
            _lib.check(_lib.load().avr_raygen_encode_bwd(C.byref(g), C.byref(self.meta), _p(_dense(rays_o)),
                                                         _p(_dense(dirs)), _p(_dense(d_vals)), ptr, ld, plane, col0,
                                                         _p(self.gmax, torch.int32), self.headroom,
                                                         C.c_void_p(self.acc.data_ptr()), float(sample_step), dev, st),
                       "avr_raygen_encode_bwd")

    def add_points(self, u, d_out, col0=0):
        dev, st = _ctx(d_out)
        ptr, ld, plane = _mat(d_out)
        _lib.check(_lib.load().avr_grid_encode_bwd(C.byref(self.meta), _p(_dense(u)), u.shape[0], ptr, ld, plane, col0,
                                                   _p(self.gmax, torch.int32), self.headroom,
                                                   C.c_void_p(self.acc.data_ptr()), dev, st), "avr_grid_encode_bwd")

    def finalize(self, grad=None, accumulate=False):
        if self.mode == "atomic":
            if grad is None:
                return self.acc
            return grad.add_(self.acc) if accumulate else grad.copy_(self.acc)
        if grad is None:
            grad = torch.empty(self.n, device=self.acc.device)
        dev, st = _ctx(grad)
        _lib.check(_lib.load().avr_grid_grad_finalize(_p(self.acc, torch.int64), self.n, _p(self.gmax, torch.int32),
                                                      self.headroom, _p(grad), 1 if accumulate else 0, dev, st),
                   "avr_grid_grad_finalize")
        return grad


# ---- dense layers, fp32 SIMT ------------------------------------------------------------------------
def gemm_workspace_bytes(M, N, K) -> int:
    return int(_lib.load().avr_gemm_workspace_bytes(M, N, K))


def gemm(la, lb, M, N, K, A, lda, B, ldb, Cmat, ldc, flags=0, aux=None, ldaux=0, workspace=None):
    """C[i,j] (+)= sum_k A(i,k) B(j,k); A/B/C/aux are tensors (possibly column views), ld in elements."""
    dev, st = _ctx(Cmat)
    ws_ptr, ws_bytes = (None, 0)
    if workspace is not None:
        ws_ptr, ws_bytes = C.c_void_p(workspace.data_ptr()), workspace.numel() * workspace.element_size()
    with _timed("sgemm", 2.0 * M * N * K, "flop"):
        _lib.check(_lib.load().avr_gemm(la, lb, M, N, K, _p(A), lda, _p(B), ldb, _p(Cmat), ldc, flags, _p(aux), ldaux,
                                        ws_ptr, ws_bytes, dev, st), "avr_gemm")


def linear_fwd(x, w, y, relu=False, relu_in=False, accum=False):
    """y[n,m] (+)= sum_k x[n,k] w[m,k]   (x, y may be column views of wider row-major buffers)."""
    flags = (GEMM_RELU if relu else 0) | (GEMM_RELU_A if relu_in else 0) | (GEMM_ACCUM if accum else 0)
    gemm(K_CONTIG, K_CONTIG, x.shape[0], w.shape[0], w.shape[1], x, x.stride(0), w, w.stride(0), y, y.stride(0), flags)


def linear_bwd_data(dy, w, dx, mask_src=None, accum=False):
    """dx[n,k] (+)= (sum_m dy[n,m] w[m,k]) * (mask_src[n,k] > 0)."""
    flags = (GEMM_MASK if mask_src is not None else 0) | (GEMM_ACCUM if accum else 0)
    gemm(K_CONTIG, I_CONTIG, dy.shape[0], w.shape[1], w.shape[0], dy, dy.stride(0), w, w.stride(0), dx, dx.stride(0),
         flags, mask_src, mask_src.stride(0) if mask_src is not None else 0)


def linear_bwd_weight(dy, x, dw, workspace, accum=False, relu_in=False):
    """dw[m,k] (+)= sum_n dy[n,m] x[n,k]  (deterministic split-K over the sample points)."""
    flags = (GEMM_ACCUM if accum else 0) | (GEMM_RELU_B if relu_in else 0)
    gemm(I_CONTIG, I_CONTIG, dy.shape[1], x.shape[1], x.shape[0], dy, dy.stride(0), x, x.stride(0), dw, dw.stride(0),
         flags, None, 0, workspace)


# ---- dense layers, tensor cores (tcgen05) on bf16 plane pairs -----------------------------------------
def planes_split(x, out: PlanePair, transpose=False, relu=False):
    """fp32 ``x[rows, cols]`` (row stride ``x.stride(0)``) -> plane pair (``out[c, r]`` when ``transpose``)."""
    dev, st = _ctx(x)
    rows, cols = x.shape
    assert x.stride(1) == 1
    _lib.check(_lib.load().avr_planes_split(_p(x), rows, cols, x.stride(0), out.ptr, out.ld, out.plane, out.kind,
                                            1 if transpose else 0, 1 if relu else 0, dev, st), "avr_planes_split")
    return out


def planes_split_many(items):
    """``items``: list of ``(x fp32 [rows, cols], out PlanePair, transpose)`` -- all of them in one launch."""
    items = list(items)
    for k0 in range(0, len(items), _lib.SPLIT_BATCH_MAX):
        chunk = items[k0:k0 + _lib.SPLIT_BATCH_MAX]
        dev, st = _ctx(chunk[0][0])
        arr = (_lib.SplitDesc * len(chunk))()
        for a, (x, out, transpose) in zip(arr, chunk):
            assert x.stride(1) == 1 and x.dtype == torch.float32 and x.is_cuda
            a.x, a.rows, a.cols, a.ld = x.data_ptr(), x.shape[0], x.shape[1], x.stride(0)
            a.planes, a.ldp, a.plane_stride, a.kind, a.transpose = out.ptr.value, out.ld, out.plane, out.kind, 1 if transpose else 0
        _lib.check(_lib.load().avr_planes_split_batch(arr, len(chunk), dev, st), "avr_planes_split_batch")
    return [out for _, out, _ in items]


def planes_merge(pp: PlanePair):
    dev, st = _ctx(pp)
    out = torch.empty(pp.rows, pp.cols, device=pp.device)
    _lib.check(_lib.load().avr_planes_merge(pp.ptr, pp.rows, pp.cols, pp.ld, pp.plane, pp.kind, _p(out), pp.cols, dev, st),
               "avr_planes_merge")
    return out


def relu_bits_empty(rows, cols, device):
    """Storage of a ReLU bitmask: int32 ``[rows, ceil(cols/32) rounded up to 4 words]``."""
    words = ((cols + 31) // 32 + 3) // 4 * 4
    return torch.empty(rows, words, dtype=torch.int32, device=device)


class NearZeroGuard:
    """Storage of the near-zero guard of the forward layers (``include/avr_b200.h``, ``avr_umma_gemm_nt``): one list
    shared by the layers of a pass (each layer's fix-up kernel has consumed it before the next GEMM starts on the
    stream) and one zeroed counter per layer, kept for inspection (``counts()``)."""

    TAU_PER_K = 1e-7        # threshold relative to the row-chunk scale: TAU_PER_K * K (tensor-core error ~6e-9 * K), at least
    TAU_MIN = 4e-6

    def __init__(self, device, layers=32, capacity=1 << 20):
        self.list = torch.empty(capacity, 2, dtype=torch.int32, device=device)
        self.counters = torch.zeros(layers, dtype=torch.int32, device=device)
        self.used = 0

    def next_slot(self):
        if self.used >= self.counters.numel():
            raise RuntimeError("NearZeroGuard: out of counters")
        self.used += 1
        return self.counters[self.used - 1:]

    def counts(self):
        return self.counters[:self.used].tolist()


def umma_nt(a: PlanePair, b: PlanePair, flags=0, c: PlanePair = None, c2: PlanePair = None, mask=None, c_f32=None,
            bits_out=None, bias_ray=None, bias_rcv=None, geom=None, guard: NearZeroGuard = None, split_k=False):
    """C[M,N] = A[M,K] B[N,K]^T on the tensor cores; C is a plane set or (UMMA_OUT_F32) an fp32 tensor.

    ``mask`` (with UMMA_MASK): int32 bitmask ``[M, words]`` gating the product (ReLU backward);
    ``bits_out``: if given, the bitmask ``(C > 0)`` is written (forward ReLU layers).
    ``bias_ray[R,N]`` / ``bias_rcv[bs,N]`` (fp32, with ``geom``): rows added to every sample point of a ray / receiver.
    ``guard``: near-zero guard storage; elements too close to zero for the tensor core's accumulation error are
    re-evaluated in fp32 by a second kernel of the same call (forward layers).
    ``split_k`` (plain fp32 outputs): short interleaved K slices into separate accumulators, summed in fp32 (long reductions).
    """
    dev, st = _ctx(a)
    M, K, N = a.rows, a.cols, b.rows
    assert b.cols == K, (b.cols, K)
    none = C.c_void_p(None)
    if bits_out is not None:
        flags |= UMMA_BITS
    if bias_ray is not None or bias_rcv is not None:
        flags |= UMMA_BIAS
    products = 6 if (a.n == 3 and b.n == 3) else 3
    assert a.f16 == b.f16, "A and B must both be bf16 plane sets or both fp16 pairs"
    ws = None
    if split_k:
        slices = int(_lib.load().avr_umma_gemm_nt_splitk_slices(K))
        ws = torch.empty(slices * M * ((N + 7) // 8 * 8), device=a.device)
    with _timed("umma_gemm", 2.0 * M * N * K, "flop", 2.0 * M * N * K * products):
        _lib.check(_lib.load().avr_umma_gemm_nt(
            M, N, K, a.ptr, a.ld, a.plane, a.kind, b.ptr, b.ld, b.plane, b.kind, flags,
            c.ptr if c is not None else none, c.ld if c is not None else 0, c.plane if c is not None else 0,
            c.kind if c is not None else 2,
            c2.ptr if c2 is not None else none, c2.ld if c2 is not None else 0, c2.plane if c2 is not None else 0,
            c2.kind if c2 is not None else 2,
            _p(mask, torch.int32), mask.stride(0) if mask is not None else 0,
            _p(bits_out, torch.int32), bits_out.stride(0) if bits_out is not None else 0,
            _p(bias_ray), bias_ray.stride(0) if bias_ray is not None else 0,
            _p(bias_rcv), bias_rcv.stride(0) if bias_rcv is not None else 0,
            geom.R if geom is not None else 0, geom.S if geom is not None else 0,
            _p(c_f32), c_f32.stride(0) if c_f32 is not None else 0,
            _p(guard.list, torch.int32) if guard is not None else none, guard.list.shape[0] if guard is not None else 0,
            _p(guard.next_slot(), torch.int32) if guard is not None else none,
            max(NearZeroGuard.TAU_MIN, NearZeroGuard.TAU_PER_K * K) if guard is not None else 0.0,
            _p(ws), ws.numel() * 4 if ws is not None else 0, dev, st), "avr_umma_gemm_nt")


def mlp_chain(x0: PlanePair, layers):
    """Fused chain of 128-wide dense layers on one launch (``avr_mlp_chain``): ``y_{l+1} = act_l(y_l W_l^T)``.

    ``x0``: bf16 plane set ``[M, k0]`` (pair or triple).  ``layers``: dicts with ``w`` (bf16 planes of ``W[n_out, k_in]``)
    and the optional ``relu`` (bool), ``mask`` (int32 bitmask multiplied into the output: ReLU backward), ``accumulate``
    (add what ``save`` already holds), ``bias`` + ``bias_group_rows`` (fp32 rows added before the activation, one per group of
    consecutive points: the per-receiver channel-embedding rows), outputs ``save`` (PlanePair), ``save_raw`` (PlanePair, un-rectified output), ``bits``
    (int32 ReLU bitmask out), ``out_f32`` (fp32 ``[M, >= n_out]`` tensor; last layer only).  Bit-identical to running the
    layers through ``umma_nt`` one by one."""
    dev, st = _ctx(x0)
    assert not x0.f16, "the chain input must be a bf16 plane set"
    n = len(layers)
    arr = (_lib.ChainLayer * n)()
    flops, executed, planes_in = 0.0, 0.0, x0.n
    for a, L in zip(arr, layers):
        w = L["w"]
        assert not w.f16
        a.w, a.ldw, a.w_plane, a.w_kind, a.n_out, a.k_in = w.ptr.value, w.ld, w.plane, w.kind, w.rows, w.cols
        a.relu = 1 if L.get("relu") else 0
        sv, raw, bits, o32, mask = L.get("save"), L.get("save_raw"), L.get("bits"), L.get("out_f32"), L.get("mask")
        if sv is not None:
            assert sv.rows == x0.rows and sv.cols == w.rows
            a.save, a.ld_save, a.save_plane, a.save_kind = sv.ptr.value, sv.ld, sv.plane, sv.kind
        if raw is not None:
            assert raw.rows == x0.rows and raw.cols == w.rows
            a.save_raw, a.ld_raw, a.raw_plane, a.raw_kind = raw.ptr.value, raw.ld, raw.plane, raw.kind
        if bits is not None:
            a.bits, a.ldbits = _p(bits, torch.int32).value, bits.stride(0)
        if mask is not None:
            a.mask, a.ldmask = _p(mask, torch.int32).value, mask.stride(0)
        a.accumulate = 1 if L.get("accumulate") else 0
        if o32 is not None:
            a.out_f32, a.ld_f32 = _p(o32).value, o32.stride(0)
        if L.get("bias") is not None:                      # fp32 [groups, n_out]: row (point // bias_group_rows) is added before the activation
            bias = L["bias"]
            assert bias.dtype == torch.float32 and bias.stride(1) == 1 and bias.shape[1] == w.rows
            a.bias, a.ld_bias, a.bias_group_rows = _p(bias).value, bias.stride(0), int(L["bias_group_rows"])
        f = 2.0 * x0.rows * w.rows * w.cols
        flops += f
        executed += f * (6 if (planes_in == 3 and w.n == 3) else 3)
        planes_in = 3                                      # (upper bound for the executed-product count of later layers)
    with _timed("mlp_chain", flops, "flop", executed):
        _lib.check(_lib.load().avr_mlp_chain(x0.rows, x0.ptr, x0.ld, x0.plane, x0.kind, x0.cols, arr, n, dev, st), "avr_mlp_chain")


mlp_chain_fwd = mlp_chain


def umma_tn_workspace_bytes(M, N, K) -> int:
    return int(_lib.load().avr_umma_gemm_tn_workspace_bytes(M, N, K))


def umma_tn(a: PlanePair, b: PlanePair, c_f32, workspace, accumulate=False):
    """C[M,N] (+)= sum_k A[k,M] B[k,N]  (A, B plane pairs over the same rows k); fp32 C (may be a column view).
    ``a`` (the gradient) is a bf16 plane set; ``b`` (the activation) a bf16 set or an fp16 pair, which the kernel turns into
    the bf16 (hi, mid) pair ``planes_split`` would have written, in shared memory."""
    dev, st = _ctx(a)
    K, M, N = a.rows, a.cols, b.cols
    assert b.rows == K
    assert not a.f16, "the gradient operand is a bf16 plane set"
    products = 6 if (a.n == 3 and b.n == 3) else 3       # six products only when both operands carry 24 bits
    with _timed("umma_gemm", 2.0 * M * N * K, "flop", 2.0 * M * N * K * products):
        _lib.check(_lib.load().avr_umma_gemm_tn(M, N, K, a.ptr, a.ld, a.plane, a.kind, b.ptr, b.ld, b.plane, b.kind, _p(c_f32),
                                                c_f32.stride(0), 1 if accumulate else 0,
                                                C.c_void_p(workspace.data_ptr()),
                                                workspace.numel() * workspace.element_size(), dev, st), "avr_umma_gemm_tn")


# ---- broadcast inputs -----------------------------------------------------------------------------
def rows_broadcast(g, src, per_receiver, dst, col0):
    dev, st = _ctx(dst)
    ptr, ld, plane = _mat(dst)
    _lib.check(_lib.load().avr_rows_broadcast(C.byref(g), _p(_dense(src)), src.shape[1], 1 if per_receiver else 0,
                                              ptr, ld, plane, _np(dst), col0, dev, st), "avr_rows_broadcast")


def rows_broadcast2(g, src, per_receiver, src2, per_receiver2, dst: PlanePair, col0):
    """Two adjacent column blocks of a plane set in one launch (whole sectors instead of two half-written ones)."""
    dev, st = _ctx(dst)
    ptr, ld, plane = _mat(dst)
    _lib.check(_lib.load().avr_rows_broadcast2(C.byref(g), _p(_dense(src)), src.shape[1], 1 if per_receiver else 0,
                                               _p(_dense(src2)), src2.shape[1], 1 if per_receiver2 else 0,
                                               ptr, ld, plane, _np(dst), col0, dev, st), "avr_rows_broadcast2")


def rows_block_sum(g, x):
    """-> fp32 ``[bs*R, w]``: sum over the S sample rows of every (receiver, ray) of ``x`` (fp32 2-D or PlanePair)."""
    dev, st = _ctx(x)
    ptr, ld, plane = _mat(x)
    w = x.cols if isinstance(x, PlanePair) else x.shape[1]
    out = torch.empty(g.bs * g.R, w, device=x.device)
    _lib.check(_lib.load().avr_rows_block_sum(C.byref(g), ptr, ld, plane, w, _p(out), dev, st), "avr_rows_block_sum")
    return out


def rows_reduce(g, d_dst, col0, w, per_receiver):
    dev, st = _ctx(d_dst)
    ptr, ld, plane = _mat(d_dst)
    rows = g.bs if per_receiver else g.R
    device = d_dst.device
    nbytes = int(_lib.load().avr_rows_reduce_workspace_bytes(C.byref(g), w, 1 if per_receiver else 0))
    ws = torch.empty(max(1, nbytes // 4), device=device)
    out = torch.empty(rows, w, device=device)
    _lib.check(_lib.load().avr_rows_reduce(C.byref(g), ptr, ld, plane, col0, w, 1 if per_receiver else 0,
                                           _p(out), _p(ws), nbytes, dev, st), "avr_rows_reduce")
    return out


# ---- ray weights / compositing / spectrum ------------------------------------------------------------
def ray_weights_fwd(g, raw, ld_raw, delta, slope, want_attn=False):
    dev, st = _ctx(raw)
    w = torch.empty(g.bs, g.R, g.S, device=raw.device)
    attn = torch.empty_like(w) if want_attn else None
    _lib.check(_lib.load().avr_ray_weights_fwd(C.byref(g), _p(raw), ld_raw, _p(_dense(delta)), float(slope), _p(attn),
                                               _p(w), dev, st), "avr_ray_weights_fwd")
    return w, attn


def ray_weights_bwd(g, raw, ld_raw, delta, slope, d_w, d_raw, ld_draw):
    dev, st = _ctx(raw)
    _lib.check(_lib.load().avr_ray_weights_bwd(C.byref(g), _p(raw), ld_raw, _p(_dense(delta)), float(slope),
                                               _p(_dense(d_w)), _p(d_raw), ld_draw, dev, st), "avr_ray_weights_bwd")


def composite_fwd(g, sig, w, delay):
    dev, st = _ctx(sig)
    nbytes = int(_lib.load().avr_composite_workspace_bytes(C.byref(g)))
    ws = torch.empty(max(1, nbytes // 4), device=sig.device)
    y = torch.empty(g.bs, g.S, g.T, device=sig.device)
    n_pts = g.bs * g.R * g.S
    with _timed("composite_fwd", float(n_pts) * (g.T * 4 + 8) + g.bs * g.S * g.T * 4.0, "byte"):
        _lib.check(_lib.load().avr_composite_fwd(C.byref(g), _p(_dense(sig)), _p(_dense(w)),
                                                 _p(_dense(delay), torch.int32), _p(y), _p(ws), nbytes, dev, st),
                   "avr_composite_fwd")
    return y


def composite_bwd(g, sig, w, delay, d_y, want_dsig=True, want_dw=True, d_sig_out=None):
    dev, st = _ctx(d_y)
    d_sig = None
    if want_dsig:
        d_sig = d_sig_out if d_sig_out is not None else torch.empty(g.bs, g.R, g.S, g.T, device=d_y.device)
    d_w = torch.empty(g.bs, g.R, g.S, device=d_y.device) if want_dw else None
    n_pts = g.bs * g.R * g.S
    nbytes = float(n_pts) * g.T * 4 * ((1 if want_dsig else 0) + (1 if (want_dw and sig is not None) else 0))
    with _timed("composite_bwd", nbytes, "byte"):
        _lib.check(_lib.load().avr_composite_bwd(C.byref(g), _p(_dense(sig)) if sig is not None else None,
                                                 _p(_dense(w)), _p(_dense(delay), torch.int32), _p(_dense(d_y)),
                                                 _p(d_sig), _p(d_w), dev, st), "avr_composite_bwd")
    return d_sig, d_w


def spectrum_fwd(g, y, tables):
    dev, st = _ctx(y)
    ldd = tables["dft"].shape[1]
    zbuf = torch.empty(g.bs * g.S, g.T, device=y.device)
    xbuf = torch.empty(g.bs * g.S, ldd, device=y.device)
    out = torch.empty(g.bs, g.T // 2 + 1, 2, device=y.device)
    _lib.check(_lib.load().avr_spectrum_fwd(C.byref(g), _p(_dense(y)), _p(tables["gain"]), _p(tables["phase"]),
                                            _p(tables["dft"]), ldd, _p(zbuf), _p(xbuf), _p(out), dev, st),
               "avr_spectrum_fwd")
    return out


def spectrum_bwd(g, d_out, tables):
    dev, st = _ctx(d_out)
    ldd = tables["dft"].shape[1]
    xbuf = torch.empty(g.bs * g.S, ldd, device=d_out.device)
    d_y = torch.empty(g.bs, g.S, g.T, device=d_out.device)
    _lib.check(_lib.load().avr_spectrum_bwd(C.byref(g), _p(_dense(d_out)), _p(tables["gain"]), _p(tables["phase"]),
                                            _p(tables["dft"]), ldd, _p(xbuf), _p(d_y), dev, st), "avr_spectrum_bwd")
    return d_y


# ---- output layer fused with the ray reduction ---------------------------------------------------------
def delay_sort(g, delay, w):
    """-> (order, sdelay int32 [bs,S,R], sw fp32 [bs,S,R]): rays of each (b,s) sorted by delay (stable)."""
    dev, st = _ctx(w)
    order = torch.empty(g.bs, g.S, g.R, dtype=torch.int32, device=w.device)
    sdelay = torch.empty_like(order)
    sw = torch.empty(g.bs, g.S, g.R, device=w.device)
    _lib.check(_lib.load().avr_delay_sort(C.byref(g), _p(_dense(delay), torch.int32), _p(_dense(w)), _p(order, torch.int32),
                                          _p(sdelay, torch.int32), _p(sw), dev, st), "avr_delay_sort")
    return order, sdelay, sw


def collapse_tspan(render_cfg) -> int:
    """Static bound on the delay spread within one (b,s): |dist(tx,p) - dist(tx,rx)| <= |p - rx| <= far."""
    return int(math.ceil(2.0 * float(render_cfg["far"]) * float(render_cfg["fs"]) / float(render_cfg["speed"]))) + 4


def collapse_fwd(g, act: PlanePair, sort, w_out, tspan):
    """-> (y[bs,S,T], prefix workspace to hand to ``collapse_bwd``)."""
    dev, st = _ctx(act)
    order, sdelay, sw = sort
    y = torch.empty(g.bs, g.S, g.T, device=act.device)
    nbytes = int(_lib.load().avr_collapse_prefix_bytes(C.byref(g), act.cols, tspan))
    prefix = torch.empty((nbytes + 3) // 4, device=act.device)
    n_pts = g.bs * g.R * g.S
    with _timed("collapse_fwd", float(n_pts) * act.cols * 4 + float(g.bs * g.S) * g.T * act.cols * 4, "byte"):
        _lib.check(_lib.load().avr_collapse_fwd(C.byref(g), act.ptr, act.ld, act.plane, act.kind, act.cols, _p(order, torch.int32),
                                                _p(sdelay, torch.int32), _p(sw), _p(w_out), w_out.stride(0), tspan,
                                                _p(prefix), prefix.numel() * 4, _p(y), dev, st), "avr_collapse_fwd")
    return y, prefix


def collapse_bwd(g, act: PlanePair, sort, w_out, d_y, tspan, prefix, d_act: PlanePair, d_wout, accumulate=False):
    """Fills ``d_act`` (plane pair) and ``d_wout`` (fp32 ``[T, width]`` view); returns ``d_w[bs,R,S]``."""
    dev, st = _ctx(act)
    order, sdelay, sw = sort
    d_w = torch.empty(g.bs, g.R, g.S, device=act.device)
    nbytes = int(_lib.load().avr_collapse_suffix_bytes(C.byref(g), act.cols, tspan))
    suffix = torch.empty((nbytes + 3) // 4, device=act.device)
    n_pts = g.bs * g.R * g.S
    with _timed("collapse_bwd", float(n_pts) * act.cols * 8 + 2.0 * g.bs * g.S * g.T * act.cols * 4, "byte"):
        _lib.check(_lib.load().avr_collapse_bwd(C.byref(g), act.ptr, act.ld, act.plane, act.kind, act.cols, _p(order, torch.int32),
                                                _p(sdelay, torch.int32), _p(sw), _p(w_out), w_out.stride(0),
                                                _p(_dense(d_y)), tspan, _p(prefix), _p(suffix), suffix.numel() * 4,
                                                d_act.ptr, d_act.ld, d_act.plane, _p(d_w), _p(d_wout), d_wout.stride(0),
                                                1 if accumulate else 0, dev, st), "avr_collapse_bwd")
    return d_w


# ---- spectrum stage with the DFT on the tensor cores ---------------------------------------------------------
SPECTRUM_SPLIT_K = True     # False only for A/B measurements of the accumulator-truncation bias

def dft_planes(tables):
    """Plane sets of the DFT matrix, built once per table set: dft^T [ldd, T] and dft [T, ldd], 3 planes each.

    The adjoint DFT also runs with six products: d_y feeds d_w = H . g, and the density gradient that follows
    (d_alpha = d_w*Tr - (sum_{k>s} d_w_k w_k)/q) is a difference of nearly equal terms that amplifies its error."""
    if "dftT_p3" not in tables:
        dft = tables["dft"]
        T, ldd = dft.shape
        tables["dftT_p3"] = planes_split(dft, PlanePair.empty(ldd, T, dft.device, n=3), transpose=True)
        tables["dft_p3"] = planes_split(dft, PlanePair.empty(T, ldd, dft.device, n=3))
    return tables["dftT_p3"], tables["dft_p3"]


def spectrum_fwd_tc(g, y, tables):
    """out[bs,F,2] = sum_s phase * DFT(y * gain), DFT as a six-product tcgen05 GEMM (fp32-grade)."""
    dev, st = _ctx(y)
    dft_t, _ = dft_planes(tables)
    rows, ldd = g.bs * g.S, dft_t.rows
    z = PlanePair.empty(rows, g.T, y.device, n=3)
    _lib.check(_lib.load().avr_spectrum_gain(C.byref(g), _p(_dense(y)), _p(tables["gain"]), z.ptr, z.ld, z.plane, z.n, dev, st),
               "avr_spectrum_gain")
    xbuf = torch.empty(rows, ldd, device=y.device)
    umma_nt(z, dft_t, UMMA_OUT_F32, c_f32=xbuf, split_k=SPECTRUM_SPLIT_K)
    out = torch.empty(g.bs, g.T // 2 + 1, 2, device=y.device)
    _lib.check(_lib.load().avr_spectrum_phase_sum(C.byref(g), _p(xbuf), ldd, _p(tables["phase"]), _p(out), dev, st),
               "avr_spectrum_phase_sum")
    return out


def spectrum_bwd_tc(g, d_out, tables):
    dev, st = _ctx(d_out)
    _, dft = dft_planes(tables)
    rows, ldd = g.bs * g.S, dft.cols
    q = PlanePair.empty(rows, ldd, d_out.device, n=3)
    _lib.check(_lib.load().avr_spectrum_phase_bwd(C.byref(g), _p(_dense(d_out)), _p(tables["phase"]), q.ptr, q.ld, q.plane,
                                                  q.n, dev, st), "avr_spectrum_phase_bwd")
    d_y = torch.empty(g.bs, g.S, g.T, device=d_out.device)
    umma_nt(q, dft, UMMA_OUT_F32, c_f32=d_y.view(rows, g.T), split_k=SPECTRUM_SPLIT_K)
    _lib.check(_lib.load().avr_spectrum_gain(C.byref(g), _p(d_y), _p(tables["gain"]), C.c_void_p(d_y.data_ptr()), g.T, 0, 0,
                                             dev, st), "avr_spectrum_gain")
    return d_y
