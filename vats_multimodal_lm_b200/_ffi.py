"""Thin ctypes layer over the C-ABI library ``csrc/libvats_attn.so`` (declared in ``include/vats_attn.h``).

There is no fallback of any kind: if the shared library is missing or a call returns a non-zero code this module
raises.  PyTorch is only used by the callers for device memory and streams; the signatures here are plain pointers
and sizes.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
# VATS_ATTN_LIB lets the tools/ scripts load a differently-built copy (e.g. the -DVATS_ENABLE_TRACE build)
LIB_PATH = os.environ.get("VATS_ATTN_LIB") or os.path.join(_HERE, "csrc", "libvats_attn.so")

KERNEL_AUTO = 0
KERNEL_TCGEN05 = 1
KERNEL_SIMT = 2
KERNEL_MID = 3

# every symbol include/vats_attn.h declares (tests check that the library exports exactly these)
EXPORTED_SYMBOLS = (
    "vats_attn_prefill",
    "vats_attn_prefill_ex",
    "vats_attn_prefill_ws",
    "vats_attn_prefill_workspace_bytes",
    "vats_attn_prefill_gather",
    "vats_attn_prefill_backward",
    "vats_attn_prefill_backward_workspace_bytes",
    "vats_attn_prefill_plan",
    "vats_attn_decode",
    "vats_attn_decode_workspace_bytes",
    "vats_attn_decode_prepare",
    "vats_attn_prefill_prepare",
    "vats_attn_prefill_prepare_table",
    "vats_attn_last_launch_count",
    "vats_attn_last_kernel",
    "vats_attn_debug_mask",
    "vats_attn_debug_tile_range",
    "vats_attn_debug_tile_is_full",
    "vats_attn_debug_set_trace",
    "vats_attn_last_error",
    "vats_attn_version",
)


class VatsAttnError(RuntimeError):
    """A C-ABI call failed; ``code`` is the VATS_ERR_* value."""

    def __init__(self, code: int, message: str):
        super().__init__(f"vats_attn error {code}: {message}")
        self.code = code


_lib = None
_lock = threading.Lock()

_i64x3 = ctypes.c_int64 * 3
_i64x2 = ctypes.c_int64 * 2
_i64x4 = ctypes.c_int64 * 4

# ctypes stride arrays are immutable inputs: build each distinct one once (the same few geometries repeat every step)
_s3_cache: dict = {}
_s2_cache: dict = {}


def _s3(strides) -> "ctypes.Array":
    key = tuple(strides)
    arr = _s3_cache.get(key)
    if arr is None:
        if len(_s3_cache) > 4096:
            _s3_cache.clear()
        arr = _s3_cache[key] = _i64x3(*key)
    return arr


def _s2(strides) -> "ctypes.Array":
    key = tuple(strides)
    arr = _s2_cache.get(key)
    if arr is None:
        if len(_s2_cache) > 4096:
            _s2_cache.clear()
        arr = _s2_cache[key] = _i64x2(*key)
    return arr


def load() -> ctypes.CDLL:
    """Load the library once; raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first (__graft_entry__.build()). "
                "vats_multimodal_lm_b200 has no CPU or PyTorch fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        vp, i, f, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
        p3 = ctypes.POINTER(ctypes.c_int64)
        lib.vats_attn_prefill.restype = i
        lib.vats_attn_prefill.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, p3, p3, p3, p3, f, i, i, i, vp]
        lib.vats_attn_prefill_ex.restype = i
        lib.vats_attn_prefill_ex.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, p3, p3, p3, p3, f, i, i, i, i, vp]
        lib.vats_attn_prefill_ws.restype = i
        lib.vats_attn_prefill_ws.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, p3, p3, p3, p3, f, i, i, i, i, f, vp, sz, vp]
        lib.vats_attn_prefill_workspace_bytes.restype = sz
        lib.vats_attn_prefill_workspace_bytes.argtypes = [i, i, i, i, i, i, p3, p3, p3, vp, vp, vp]
        lib.vats_attn_prefill_gather.restype = i
        lib.vats_attn_prefill_gather.argtypes = [vp, vp, vp, ctypes.POINTER(vp), i, i, i, i, i, i, vp, vp, i, i, i, i, i, i,
                                                 p3, p3, p3, p3, f, i, i, i, f, vp, sz, vp]
        lib.vats_attn_prefill_backward.restype = i
        lib.vats_attn_prefill_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i,
                                                   p3, p3, p3, p3, p3, f, i, i, i, vp, sz, vp]
        lib.vats_attn_prefill_backward_workspace_bytes.restype = sz
        lib.vats_attn_prefill_backward_workspace_bytes.argtypes = [i, i, i]
        lib.vats_attn_prefill_plan.restype = i
        lib.vats_attn_prefill_plan.argtypes = [i, i, i, i, i, i, p3, p3, p3, p3, vp, vp, vp]
        lib.vats_attn_decode.restype = i
        lib.vats_attn_decode.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, p3, p3, p3, p3, f, i, vp, sz, vp]
        lib.vats_attn_decode_workspace_bytes.restype = sz
        lib.vats_attn_decode_workspace_bytes.argtypes = [i, i, i, i, i, i]
        lib.vats_attn_decode_prepare.restype = i
        lib.vats_attn_decode_prepare.argtypes = [vp, vp, vp, i, vp, vp, vp, vp, vp, vp, i, i, i, i, i, p3, p3, p3, p3, p3, p3,
                                                 i, f, vp]
        lib.vats_attn_prefill_prepare.restype = i
        lib.vats_attn_prefill_prepare.argtypes = [vp, vp, vp, i, vp, vp, vp, vp, vp, i, i, i, i, i, i, p3, p3, p3, p3, p3, p3,
                                                  i, f, vp]
        lib.vats_attn_prefill_prepare_table.restype = i
        lib.vats_attn_prefill_prepare_table.argtypes = [vp, vp, vp, i, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i,
                                                        p3, p3, p3, p3, p3, p3, i, f, vp]
        lib.vats_attn_last_launch_count.restype = i
        lib.vats_attn_last_launch_count.argtypes = []
        lib.vats_attn_last_kernel.restype = i
        lib.vats_attn_last_kernel.argtypes = []
        lib.vats_attn_debug_mask.restype = i
        lib.vats_attn_debug_mask.argtypes = [vp, vp, vp, i, i, i, i, i, i, vp]
        lib.vats_attn_debug_tile_range.restype = i
        lib.vats_attn_debug_tile_range.argtypes = [i, i, i, i, i, i, i, i, ctypes.POINTER(i), ctypes.POINTER(i)]
        lib.vats_attn_debug_tile_is_full.restype = i
        lib.vats_attn_debug_tile_is_full.argtypes = [i, i, i, i, i, i, i, i, i]
        lib.vats_attn_debug_set_trace.restype = None
        lib.vats_attn_debug_set_trace.argtypes = [vp, i]
        lib.vats_attn_last_error.restype = ctypes.c_char_p
        lib.vats_attn_last_error.argtypes = []
        lib.vats_attn_version.restype = i
        lib.vats_attn_version.argtypes = []
        _lib = lib
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise VatsAttnError(rc, load().vats_attn_last_error().decode("utf-8", "replace"))


def version() -> int:
    return load().vats_attn_version()


def last_launch_count() -> int:
    return load().vats_attn_last_launch_count()


LAUNCHED = {0: "none", 1: "prefill_tc", 2: "prefill_tc_ldg", 3: "prefill_short", 4: "prefill_simt", 5: "decode_mma",
            6: "decode_split", 7: "prefill_prepare", 8: "decode_prepare", 9: "prefill_mid", 10: "backward"}


def last_kernel() -> str:
    """Name of the kernel the last successful compute call of this thread ended in (include/vats_attn.h)."""
    return LAUNCHED.get(load().vats_attn_last_kernel(), "unknown")


def prefill(q_ptr: int, k_ptr: int, v_ptr: int, o_ptr: int, q_valid_ptr: Optional[int], k_valid_ptr: Optional[int],
            N: int, Tq: int, Tk: int, H: int, G: int, hd: int,
            q_strides: Sequence[int], k_strides: Sequence[int], v_strides: Sequence[int], o_strides: Sequence[int],
            scale: float, causal: bool, left: int, right: int, stream: int, kernel: int = KERNEL_AUTO,
            workspace_ptr: Optional[int] = None, workspace_bytes: int = 0, logit_bound: float = 0.0) -> None:
    lib = load()
    _check(lib.vats_attn_prefill_ws(
        q_ptr, k_ptr, v_ptr, o_ptr, q_valid_ptr, k_valid_ptr, N, Tq, Tk, H, G, hd,
        _s3(q_strides), _s3(k_strides), _s3(v_strides), _s3(o_strides),
        float(scale), int(bool(causal)), int(left), int(right), int(kernel), float(logit_bound), workspace_ptr,
        workspace_bytes, stream))


def prefill_gather(q_ptr: int, k_ptr: int, v_ptr: int, o_rank_ptrs: Sequence[int], rank: int, seq_offset: int,
                   head_offset: int, N_total: int, H_total: int, q_valid_ptr: Optional[int], k_valid_ptr: Optional[int],
                   N: int, Tq: int, Tk: int, H: int, G: int, hd: int, q_strides, k_strides, v_strides, o_strides,
                   scale: float, causal: bool, left: int, right: int, stream: int,
                   workspace_ptr: Optional[int] = None, workspace_bytes: int = 0, logit_bound: float = 0.0) -> None:
    """vats_attn_prefill_gather: local attention whose O tiles are stored into every rank's gathered output."""
    world = len(o_rank_ptrs)
    arr = (ctypes.c_void_p * world)(*[int(p) for p in o_rank_ptrs])
    _check(load().vats_attn_prefill_gather(
        q_ptr, k_ptr, v_ptr, arr, world, int(rank), int(seq_offset), int(head_offset), int(N_total), int(H_total),
        q_valid_ptr, k_valid_ptr, N, Tq, Tk, H, G, hd, _s3(q_strides), _s3(k_strides), _s3(v_strides), _s3(o_strides),
        float(scale), int(bool(causal)), int(left), int(right), float(logit_bound), workspace_ptr, workspace_bytes,
        stream))


def prefill_backward(q_ptr, k_ptr, v_ptr, o_ptr, do_ptr, dq_ptr, dk_ptr, dv_ptr, q_valid_ptr, k_valid_ptr,
                     N: int, Tq: int, Tk: int, H: int, G: int, hd: int, q_strides, k_strides, v_strides, o_strides,
                     do_strides, scale: float, causal: bool, left: int, right: int, workspace_ptr, workspace_bytes: int,
                     stream: int) -> None:
    _check(load().vats_attn_prefill_backward(
        q_ptr, k_ptr, v_ptr, o_ptr, do_ptr, dq_ptr, dk_ptr, dv_ptr, q_valid_ptr, k_valid_ptr, N, Tq, Tk, H, G, hd,
        _s3(q_strides), _s3(k_strides), _s3(v_strides), _s3(o_strides), _s3(do_strides), float(scale),
        int(bool(causal)), int(left), int(right), workspace_ptr, workspace_bytes, stream))


def prefill_backward_workspace_bytes(N: int, Tq: int, H: int) -> int:
    return int(load().vats_attn_prefill_backward_workspace_bytes(N, Tq, H))


def prefill_workspace_bytes(N: int, Tq: int, Tk: int, H: int, G: int, hd: int, q_strides, k_strides, v_strides,
                            q_ptr: int, k_ptr: int, v_ptr: int) -> int:
    """Scratch the tensor-core kernel wants for tensors TMA cannot address (0 = none); see include/vats_attn.h."""
    return int(load().vats_attn_prefill_workspace_bytes(N, Tq, Tk, H, G, hd, _s3(q_strides), _s3(k_strides),
                                                        _s3(v_strides), q_ptr, k_ptr, v_ptr))


def prefill_plan(N: int, Tq: int, Tk: int, H: int, G: int, hd: int, q_strides, k_strides, v_strides, o_strides,
                 q_ptr: int, k_ptr: int, v_ptr: int) -> int:
    return load().vats_attn_prefill_plan(N, Tq, Tk, H, G, hd, _s3(q_strides), _s3(k_strides),
                                         _s3(v_strides), _s3(o_strides), q_ptr, k_ptr, v_ptr)


def decode_workspace_bytes(B: int, H: int, G: int, hd: int, S_max: int, left: int) -> int:
    return int(load().vats_attn_decode_workspace_bytes(B, H, G, hd, S_max, left))


def decode(q_ptr: int, k_ptr: int, v_ptr: int, o_ptr: int, seq_lens_ptr: int, B: int, H: int, G: int, hd: int,
           S_max: int, q_strides: Sequence[int], k_strides: Sequence[int], v_strides: Sequence[int],
           o_strides: Sequence[int], scale: float, left: int, workspace_ptr: Optional[int], workspace_bytes: int,
           stream: int) -> None:
    lib = load()
    _check(lib.vats_attn_decode(
        q_ptr, k_ptr, v_ptr, o_ptr, seq_lens_ptr, B, H, G, hd, S_max,
        _s2(q_strides), _s3(k_strides), _s3(v_strides), _s2(o_strides),
        float(scale), int(left), workspace_ptr, workspace_bytes, stream))


def decode_prepare(q_in_ptr: int, k_in_ptr: int, v_in_ptr: int, in_fp32: bool, q_out_ptr: int, k_cache_ptr: int,
                   v_cache_ptr: int, seq_lens_ptr: int, cos_ptr: Optional[int], sin_ptr: Optional[int], B: int, H: int,
                   G: int, hd: int, S_max: int, qin_strides, kin_strides, vin_strides, qout_strides, k_strides,
                   v_strides, qk_norm: bool, eps: float, stream: int) -> None:
    _check(load().vats_attn_decode_prepare(
        q_in_ptr, k_in_ptr, v_in_ptr, int(bool(in_fp32)), q_out_ptr, k_cache_ptr, v_cache_ptr, seq_lens_ptr, cos_ptr,
        sin_ptr, B, H, G, hd, S_max, _s2(qin_strides), _s2(kin_strides), _s2(vin_strides),
        _s2(qout_strides), _s3(k_strides), _s3(v_strides), int(bool(qk_norm)), float(eps), stream))


def prefill_prepare(q_in_ptr: int, k_in_ptr: int, v_in_ptr: int, in_fp32: bool, q_out_ptr: int, k_out_ptr: int,
                    v_out_ptr: int, cos_ptr: Optional[int], sin_ptr: Optional[int], N: int, T: int, H: int, G: int,
                    hd: int, pos0: int, qin_strides, kin_strides, vin_strides, qout_strides, kout_strides, vout_strides,
                    qk_norm: bool, eps: float, stream: int) -> None:
    _check(load().vats_attn_prefill_prepare(
        q_in_ptr, k_in_ptr, v_in_ptr, int(bool(in_fp32)), q_out_ptr, k_out_ptr, v_out_ptr, cos_ptr, sin_ptr, N, T, H, G,
        hd, pos0, _s3(qin_strides), _s3(kin_strides), _s3(vin_strides), _s3(qout_strides),
        _s3(kout_strides), _s3(vout_strides), int(bool(qk_norm)), float(eps), stream))


def prefill_prepare_table(q_in_ptr, k_in_ptr, v_in_ptr, in_fp32: bool, q_out_ptr, k_out_ptr, v_out_ptr, cos_ptr, sin_ptr,
                          partner_ptr, No: int, Ni: int, T: int, H: int, G: int, hd: int, qin_strides, kin_strides,
                          vin_strides, qout_strides, kout_strides, vout_strides, qk_norm: bool, eps: float,
                          stream: int) -> None:
    _check(load().vats_attn_prefill_prepare_table(
        q_in_ptr, k_in_ptr, v_in_ptr, int(bool(in_fp32)), q_out_ptr, k_out_ptr, v_out_ptr, cos_ptr, sin_ptr, partner_ptr,
        No, Ni, T, H, G, hd, _i64x4(*qin_strides), _i64x4(*kin_strides), _i64x4(*vin_strides), _s3(qout_strides),
        _s3(kout_strides), _s3(vout_strides), int(bool(qk_norm)), float(eps), stream))


def debug_mask(out_ptr: int, q_valid_ptr: Optional[int], k_valid_ptr: Optional[int], N: int, Tq: int, Tk: int,
               causal: bool, left: int, right: int, stream: int) -> None:
    _check(load().vats_attn_debug_mask(out_ptr, q_valid_ptr, k_valid_ptr, N, Tq, Tk, int(bool(causal)), int(left),
                                       int(right), stream))


def debug_tile_range(q0: int, block_m: int, block_n: int, Tq: int, Tk: int, causal: bool, left: int, right: int):
    first, last = ctypes.c_int(0), ctypes.c_int(0)
    _check(load().vats_attn_debug_tile_range(q0, block_m, block_n, Tq, Tk, int(bool(causal)), int(left), int(right),
                                             ctypes.byref(first), ctypes.byref(last)))
    return first.value, last.value


def debug_tile_is_full(tile: int, q0: int, block_m: int, block_n: int, Tq: int, Tk: int, causal: bool, left: int,
                       right: int) -> bool:
    return bool(load().vats_attn_debug_tile_is_full(tile, q0, block_m, block_n, Tq, Tk, int(bool(causal)), int(left),
                                                    int(right)))
