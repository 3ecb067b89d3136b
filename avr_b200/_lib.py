"""ctypes binding of ``libavr_b200.so`` -- the C-ABI declared in ``include/avr_b200.h``.

There is NO fallback: if the shared library is missing or a call fails, the caller gets an
exception (``AVRLibraryError``).  The library is built in-tree by ``avr_b200/build.py``
(``__graft_entry__.build()``); importing this module never compiles anything.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_LEVELS = 32
ABI_VERSION = 2
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libavr_b200.so")

GEMM_RELU, GEMM_ACCUM, GEMM_MASK, GEMM_RELU_A, GEMM_RELU_B = 1, 2, 4, 8, 16
K_CONTIG, I_CONTIG = 0, 1
UMMA_RELU, UMMA_ACCUM, UMMA_MASK, UMMA_OUT_F32, UMMA_DUAL_RELU, UMMA_BITS, UMMA_BIAS = 1, 2, 4, 8, 16, 32, 64
UMMA_DUAL_COPY = 2048


class AVRLibraryError(RuntimeError):
    pass


class RenderGeom(C.Structure):
    _fields_ = [("bs", C.c_int32), ("R", C.c_int32), ("S", C.c_int32), ("T", C.c_int32),
                ("xyz_min", C.c_float), ("xyz_span", C.c_float), ("fs", C.c_float), ("speed", C.c_float)]


class GridMeta(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("n_feat", C.c_int32),
                ("scale", C.c_float * MAX_LEVELS), ("res", C.c_uint32 * MAX_LEVELS),
                ("size", C.c_uint32 * MAX_LEVELS), ("offset", C.c_uint32 * MAX_LEVELS),
                ("total", C.c_uint32), ("stride32", C.c_int32)]


class ChainLayer(C.Structure):
    """``avr_chain_layer`` (include/avr_b200.h, fused chain of 128-wide dense layers)."""
    _fields_ = [("w", C.c_void_p), ("ldw", C.c_int64), ("w_plane", C.c_int64), ("w_kind", C.c_int32), ("n_out", C.c_int32),
                ("k_in", C.c_int32), ("relu", C.c_int32),
                ("save", C.c_void_p), ("ld_save", C.c_int64), ("save_plane", C.c_int64), ("save_kind", C.c_int32),
                ("save_raw", C.c_void_p), ("ld_raw", C.c_int64), ("raw_plane", C.c_int64), ("raw_kind", C.c_int32),
                ("bits", C.c_void_p), ("ldbits", C.c_int64),
                ("mask", C.c_void_p), ("ldmask", C.c_int64), ("accumulate", C.c_int32),
                ("out_f32", C.c_void_p), ("ld_f32", C.c_int64),
                ("bias", C.c_void_p), ("ld_bias", C.c_int64), ("bias_group_rows", C.c_int64)]


class SplitDesc(C.Structure):
    """``avr_split_desc`` (include/avr_b200.h)."""
    _fields_ = [("x", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int64), ("ld", C.c_int64),
                ("planes", C.c_void_p), ("ldp", C.c_int64), ("plane_stride", C.c_int64),
                ("kind", C.c_int32), ("transpose", C.c_int32)]


SPLIT_BATCH_MAX = 32
CHAIN_MAX_LAYERS = 8
_P = C.c_void_p
_I32, _I64, _F = C.c_int32, C.c_int64, C.c_float
_G, _M = C.POINTER(RenderGeom), C.POINTER(GridMeta)

# name -> (restype, argtypes); must list every symbol of include/avr_b200.h (tests check this)
SIGNATURES = {
    "avr_abi_version": (C.c_int, []),
    "avr_last_error": (C.c_char_p, []),
    "avr_launch_count": (_I64, [C.c_int]),
    "avr_sample_points": (C.c_int, [_G, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "avr_aux_inputs": (C.c_int, [_G, _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "avr_raygen_encode_fwd": (C.c_int, [_G, _M, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _I32, _I32, _P, C.c_int, _P]),
    "avr_raygen_encode_bwd": (C.c_int, [_G, _M, _P, _P, _P, _P, _I64, _I64, _I32, _P, _I32, _P, C.c_float, C.c_int, _P]),
    "avr_grid_encode_fwd": (C.c_int, [_M, _P, _I64, _P, _P, _I64, _I64, _I32, _I32, _I32, C.c_int, _P]),
    "avr_grid_encode_bwd": (C.c_int, [_M, _P, _I64, _P, _I64, _I64, _I32, _P, _I32, _P, C.c_int, _P]),
    "avr_absmax_bits": (C.c_int, [_P, _I64, _I64, _I64, _I32, _I32, _P, C.c_int, _P]),
    "avr_grid_grad_finalize": (C.c_int, [_P, _I64, _P, _I32, _P, C.c_int, C.c_int, _P]),
    "avr_gemm_workspace_bytes": (_I64, [_I64, _I64, _I64]),
    "avr_gemm": (C.c_int, [C.c_int, C.c_int, _I64, _I64, _I64, _P, _I64, _P, _I64, _P, _I64, C.c_int, _P, _I64,
                           _P, _I64, C.c_int, _P]),
    "avr_planes_split": (C.c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "avr_planes_split_batch": (C.c_int, [C.POINTER(SplitDesc), _I32, C.c_int, _P]),
    "avr_planes_merge": (C.c_int, [_P, _I64, _I64, _I64, _I64, C.c_int, _P, _I64, C.c_int, _P]),
    "avr_umma_gemm_nt": (C.c_int, [_I64, _I64, _I64, _P, _I64, _I64, C.c_int, _P, _I64, _I64, C.c_int, C.c_int, _P, _I64,
                                   _I64, C.c_int, _P, _I64, _I64, C.c_int, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _P, _I64,
                                   _P, _I64, _P, C.c_float, _P, _I64, C.c_int, _P]),
    "avr_umma_gemm_nt_splitk_slices": (_I64, [_I64]),
    "avr_mlp_chain": (C.c_int, [_I64, _P, _I64, _I64, _I32, _I32, C.POINTER(ChainLayer), _I32, C.c_int, _P]),
    "avr_umma_gemm_tn_workspace_bytes": (_I64, [_I64, _I64, _I64]),
    "avr_umma_gemm_tn": (C.c_int, [_I64, _I64, _I64, _P, _I64, _I64, C.c_int, _P, _I64, _I64, C.c_int, _P, _I64, C.c_int,
                                   _P, _I64, C.c_int, _P]),
    "avr_delay_sort": (C.c_int, [_G, _P, _P, _P, _P, _P, C.c_int, _P]),
    "avr_collapse_prefix_bytes": (_I64, [_G, _I32, _I32]),
    "avr_collapse_suffix_bytes": (_I64, [_G, _I32, _I32]),
    "avr_collapse_fwd": (C.c_int, [_G, _P, _I64, _I64, _I32, _I32, _P, _P, _P, _P, _I64, _I32, _P, _I64, _P, C.c_int, _P]),
    "avr_collapse_bwd": (C.c_int, [_G, _P, _I64, _I64, _I32, _I32, _P, _P, _P, _P, _I64, _P, _I32, _P, _P, _I64, _P, _I64, _I64,
                                   _P, _P, _I64, C.c_int, C.c_int, _P]),
    "avr_rows_broadcast": (C.c_int, [_G, _P, _I32, C.c_int, _P, _I64, _I64, _I32, _I32, C.c_int, _P]),
    "avr_rows_broadcast2": (C.c_int, [_G, _P, _I32, C.c_int, _P, _I32, C.c_int, _P, _I64, _I64, _I32, _I32, C.c_int, _P]),
    "avr_rows_block_sum": (C.c_int, [_G, _P, _I64, _I64, _I32, _P, C.c_int, _P]),
    "avr_rows_reduce_workspace_bytes": (_I64, [_G, _I32, C.c_int]),
    "avr_rows_reduce": (C.c_int, [_G, _P, _I64, _I64, _I32, _I32, C.c_int, _P, _P, _I64, C.c_int, _P]),
    "avr_ray_weights_fwd": (C.c_int, [_G, _P, _I64, _P, _F, _P, _P, C.c_int, _P]),
    "avr_ray_weights_bwd": (C.c_int, [_G, _P, _I64, _P, _F, _P, _P, _I64, C.c_int, _P]),
    "avr_composite_workspace_bytes": (_I64, [_G]),
    "avr_composite_fwd": (C.c_int, [_G, _P, _P, _P, _P, _P, _I64, C.c_int, _P]),
    "avr_composite_bwd": (C.c_int, [_G, _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "avr_spectrum_fwd": (C.c_int, [_G, _P, _P, _P, _P, _I64, _P, _P, _P, C.c_int, _P]),
    "avr_spectrum_bwd": (C.c_int, [_G, _P, _P, _P, _P, _I64, _P, _P, C.c_int, _P]),
    "avr_crit_irfft": (C.c_int, [_P, _I32, _I32, _P, _I64, _P, C.c_int, _P]),
    "avr_crit_irfft_adjoint": (C.c_int, [_P, _I32, _I32, _P, _I64, _P, C.c_int, _P]),
    "avr_crit_freq_terms": (C.c_int, [_P, _P, _I32, _I32, _F, _F, _F, _P, _P, C.c_int, _P]),
    "avr_crit_time_l1": (C.c_int, [_P, _P, _I32, _I32, _F, _P, _P, C.c_int, _P]),
    "avr_crit_energy_workspace_bytes": (_I64, [_I32, _I32, _I32]),
    "avr_crit_energy": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _F, _P, _P, _P, _I64, C.c_int, _P]),
    "avr_crit_stft_fwd": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P, _F, _P, _P, _P, C.c_int, _P]),
    "avr_crit_stft_bwd": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P, _F, _F, _F, _P, _P, C.c_int, _P]),
    "avr_adam_workspace_bytes": (_I64, []),
    "avr_fused_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _F, _I64, C.c_int, _P, _P, _I64, C.c_int, _P]),
    "avr_spectrum_gain": (C.c_int, [_G, _P, _P, _P, _I64, _I64, _I32, C.c_int, _P]),
    "avr_spectrum_phase_sum": (C.c_int, [_G, _P, _I64, _P, _P, C.c_int, _P]),
    "avr_spectrum_phase_bwd": (C.c_int, [_G, _P, _P, _P, _I64, _I64, _I32, C.c_int, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load (once) and return the library; raises ``AVRLibraryError`` when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AVRLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m avr_b200.build` (or __graft_entry__.build()). "
            "avr_b200 has no CPU / PyTorch fallback.")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:
        raise AVRLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise AVRLibraryError(f"{LIB_PATH} does not export {name}") from exc
        fn.restype, fn.argtypes = res, args
    if lib.avr_abi_version() != ABI_VERSION:
        raise AVRLibraryError("libavr_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().avr_last_error().decode("utf-8", "replace")
        raise AVRLibraryError(f"{what or 'avr_b200 call'} failed (code {rc}): {msg}")


def launch_count(reset: bool = False) -> int:
    return int(load().avr_launch_count(1 if reset else 0))
