"""``torch.library`` custom ops over the C-ABI library — the only compute entry points of the package.

    torch.ops.vats.gqa_swa_prefill(q, k, v, q_valid, k_valid, scale, causal, left, right, kernel) -> o
    torch.ops.vats.gqa_swa_decode(q, k_cache, v_cache, seq_lens, scale, left) -> o
    torch.ops.vats.decode_prepare(q, k, v, k_cache, v_cache, seq_lens, cos, sin, qk_norm, eps) -> q_rotated  (appends k, v)
    ops.attn_mask(q_valid, k_valid, N, Tq, Tk, causal, left, right)   (plain function) -> uint8 [N,Tq,Tk]

They replace the reference's single library call ``F.scaled_dot_product_attention``
(src/optimized_attention.py:709-714, vit_2d/optimized_attention.py:396-402, vit_3d/optimized_attention.py:302-307)
plus the K/V head expansion in front of it (utils/attention_utils.py:7-27).  Tensors must live on a CUDA (sm_100)
device in bf16; anything else raises — there is deliberately no eager / CPU path here.
"""
from __future__ import annotations

import collections
from typing import List, Optional

import torch

from . import _ffi

__all__ = ["gqa_swa_prefill", "gqa_swa_prefill_bwd", "gqa_swa_prefill_gather", "gqa_swa_decode", "decode_prepare", "reset_decode_workspaces", "prefill_prepare", "prefill_prepare_views", "prefill_prepare_table", "prefill_prepare_table_views", "attn_mask", "KERNEL_AUTO", "KERNEL_TCGEN05", "KERNEL_SIMT", "KERNEL_MID"]

KERNEL_AUTO = _ffi.KERNEL_AUTO
KERNEL_TCGEN05 = _ffi.KERNEL_TCGEN05
KERNEL_SIMT = _ffi.KERNEL_SIMT
KERNEL_MID = _ffi.KERNEL_MID


def _require_cuda_bf16(name: str, t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"vats attention ops run only on a CUDA sm_100 device; `{name}` is on {t.device}. "
            "There is no CPU fallback (use oracle/ for CPU checking in tests)."
        )
    if t.dtype != torch.bfloat16:
        raise RuntimeError(f"`{name}` must be bfloat16, got {t.dtype}")


def _rowmajor_last(t: torch.Tensor) -> torch.Tensor:
    """Kernels need the head_dim axis contiguous; every other stride is passed through."""
    return t if t.stride(-1) == 1 or t.size(-1) == 1 else t.contiguous()


def _valid_u8(name: str, m: Optional[torch.Tensor], N: int, T: int, device) -> Optional[torch.Tensor]:
    if m is None:
        return None
    if m.shape != (N, T):
        raise ValueError(f"`{name}` must have shape {(N, T)}, got {tuple(m.shape)}")
    if m.device != device:
        raise RuntimeError(f"`{name}` must be on {device}")
    if m.dtype == torch.bool:
        return m.contiguous().view(torch.uint8)
    return (m != 0).to(torch.uint8).contiguous()


@torch.library.custom_op("vats::gqa_swa_prefill", mutates_args=(), device_types="cuda")
def gqa_swa_prefill(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, q_valid: Optional[torch.Tensor],
                    k_valid: Optional[torch.Tensor], scale: float, causal: bool, left: int, right: int,
                    kernel: int = 0, logit_bound: float = 0.0) -> torch.Tensor:
    """q [N,Tq,H,hd], k/v [N,Tk,G,hd] (bf16, any strides with hd contiguous) -> o [N,Tq,H,hd] bf16.

    logit_bound > 0 promises |<q, k>| <= logit_bound for every pair (1.0 behind the reference's qk-norm): the kernels
    then skip the row-maximum pass (softmax is shift-invariant; same result).  0 = unknown."""
    _require_cuda_bf16("q", q)
    _require_cuda_bf16("k", k)
    _require_cuda_bf16("v", v)
    if q.dim() != 4 or k.dim() != 4 or v.dim() != 4:
        raise ValueError("q, k, v must be 4-D [N, T, heads, head_dim]")
    N, Tq, H, hd = q.shape
    Nk, Tk, G, hdk = k.shape
    if v.shape != k.shape or Nk != N or hdk != hd:
        raise ValueError(f"shape mismatch: q {tuple(q.shape)} k {tuple(k.shape)} v {tuple(v.shape)}")
    if G <= 0 or H % G != 0:
        raise ValueError(f"num_heads ({H}) must be divisible by query_groups ({G})")
    q, k, v = _rowmajor_last(q), _rowmajor_last(k), _rowmajor_last(v)
    qv = _valid_u8("q_valid", q_valid, N, Tq, q.device)
    kv = _valid_u8("k_valid", k_valid, N, Tk, q.device)
    o = torch.empty((N, Tq, H, hd), dtype=torch.bfloat16, device=q.device)
    if o.numel() == 0:
        return o
    with torch.cuda.device(q.device):
        stream = torch.cuda.current_stream().cuda_stream
        qs, ks, vs = q.stride()[:3], k.stride()[:3], v.stride()[:3]
        kp, vp = (k.data_ptr(), v.data_ptr()) if k.numel() else (None, None)
        # rows TMA cannot address (dense hd 60 / 66): the kernel repacks them into scratch that WE own (the library
        # allocates nothing); tensors in the modules' padded layout need none
        ws, ws_bytes = None, 0
        if kp is not None and kernel != KERNEL_SIMT and (hd % 8 != 0 or not _tma_strides(qs, ks, vs) or
                                                         (q.data_ptr() | kp | vp) & 15):
            ws_bytes = _ffi.prefill_workspace_bytes(N, Tq, Tk, H, G, hd, qs, ks, vs, q.data_ptr(), kp, vp)
            if ws_bytes:
                ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=q.device)
        _ffi.prefill(q.data_ptr(), kp, vp, o.data_ptr(), qv.data_ptr() if qv is not None else None,
                     kv.data_ptr() if kv is not None else None,
                     N, Tq, Tk, H, G, hd, qs, ks, vs, o.stride()[:3],
                     scale, causal, left, right, stream, kernel, ws.data_ptr() if ws is not None else None, ws_bytes,
                     logit_bound)
    return o


@torch.library.custom_op("vats::gqa_swa_prefill_bwd", mutates_args=(), device_types="cuda")
def gqa_swa_prefill_bwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, dout: torch.Tensor,
                        q_valid: Optional[torch.Tensor], k_valid: Optional[torch.Tensor], scale: float, causal: bool,
                        left: int, right: int) -> List[torch.Tensor]:
    """Backward of `vats::gqa_swa_prefill` (vats_attn_prefill_backward): forward inputs, forward output `o` and dL/do
    -> [dq, dk, dv] (bf16, dense; dk / dv carry the G un-expanded KV heads, summed over the heads of each group)."""
    for name, t in (("q", q), ("k", k), ("v", v), ("o", o), ("dout", dout)):
        _require_cuda_bf16(name, t)
    N, Tq, H, hd = q.shape
    Tk, G = k.size(1), k.size(2)
    q, k, v, o, dout = (_rowmajor_last(t) for t in (q, k, v, o, dout))
    qv = _valid_u8("q_valid", q_valid, N, Tq, q.device)
    kv = _valid_u8("k_valid", k_valid, N, Tk, q.device)
    dq = torch.empty((N, Tq, H, hd), dtype=torch.bfloat16, device=q.device)
    dk = torch.empty((N, Tk, G, hd), dtype=torch.bfloat16, device=q.device)
    dv = torch.empty((N, Tk, G, hd), dtype=torch.bfloat16, device=q.device)
    if q.numel() == 0 or k.numel() == 0:
        return [dq.zero_(), dk.zero_(), dv.zero_()]
    with torch.cuda.device(q.device):
        stream = torch.cuda.current_stream().cuda_stream
        nbytes = max(_ffi.prefill_backward_workspace_bytes(N, Tq, H), 16)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=q.device)
        _ffi.prefill_backward(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), dout.data_ptr(), dq.data_ptr(),
                              dk.data_ptr(), dv.data_ptr(), qv.data_ptr() if qv is not None else None,
                              kv.data_ptr() if kv is not None else None, N, Tq, Tk, H, G, hd, q.stride()[:3],
                              k.stride()[:3], v.stride()[:3], o.stride()[:3], dout.stride()[:3], scale, causal, left,
                              right, ws.data_ptr(), ws.numel(), stream)
    return [dq, dk, dv]


@gqa_swa_prefill_bwd.register_fake
def _(q, k, v, o, dout, q_valid, k_valid, scale, causal, left, right):
    return [q.new_empty(q.shape, dtype=torch.bfloat16), k.new_empty(k.shape, dtype=torch.bfloat16),
            v.new_empty(v.shape, dtype=torch.bfloat16)]


def _prefill_setup_context(ctx, inputs, output):
    q, k, v, q_valid, k_valid, scale, causal, left, right, _kernel, _bound = inputs
    ctx.save_for_backward(q, k, v, output, q_valid, k_valid)
    ctx.attn_args = (scale, causal, left, right)


def _prefill_backward(ctx, dout):
    q, k, v, o, q_valid, k_valid = ctx.saved_tensors
    scale, causal, left, right = ctx.attn_args
    dq, dk, dv = torch.ops.vats.gqa_swa_prefill_bwd(q, k, v, o, dout.contiguous(), q_valid, k_valid, scale, causal, left,
                                                    right)
    return dq, dk, dv, None, None, None, None, None, None, None, None


gqa_swa_prefill.register_autograd(_prefill_backward, setup_context=_prefill_setup_context)


def gqa_swa_prefill_gather(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor,
                           peer_ptrs: List[int], rank: int, seq_offset: int, head_offset: int,
                           q_valid: Optional[torch.Tensor], k_valid: Optional[torch.Tensor], scale: float, causal: bool,
                           left: int, right: int, logit_bound: float = 0.0) -> None:
    """Local attention with the multi-GPU output gather fused into the kernel epilogue (vats_attn_prefill_gather).

    q [N,Tq,H,hd], k/v [N,Tk,G,hd] are this rank's units; `out` is this rank's copy of the gathered
    [N_total, Tq, H_total, hd] tensor and `peer_ptrs[r]` the device pointer of rank r's copy as mapped into this process
    (`sharding.FusedGather` obtains them from symmetric memory).  Every O tile is written into all copies at
    (seq_offset + n, :, head_offset + h, :).  Plain function (not a custom op): it writes peer memory, which the
    dispatcher cannot describe; the caller brackets it with cross-rank barriers (see `sharding.FusedGather`)."""
    _require_cuda_bf16("q", q)
    _require_cuda_bf16("k", k)
    _require_cuda_bf16("v", v)
    _require_cuda_bf16("out", out)
    if q.dim() != 4 or k.dim() != 4 or v.shape != k.shape or out.dim() != 4:
        raise ValueError("q, k, v, out must be 4-D [N, T, heads, head_dim]")
    N, Tq, H, hd = q.shape
    Tk, G = k.size(1), k.size(2)
    N_total, Tq_o, H_total, hd_o = out.shape
    if Tq_o != Tq or hd_o != hd or k.size(0) != N or k.size(3) != hd or H % G != 0:
        raise ValueError(f"shape mismatch: q {tuple(q.shape)} k {tuple(k.shape)} out {tuple(out.shape)}")
    if peer_ptrs[rank] != out.data_ptr():
        raise ValueError("peer_ptrs[rank] must be the local gathered tensor")
    q, k, v = _rowmajor_last(q), _rowmajor_last(k), _rowmajor_last(v)
    qv = _valid_u8("q_valid", q_valid, N, Tq, q.device)
    kv = _valid_u8("k_valid", k_valid, N, Tk, q.device)
    if q.numel() == 0:
        return
    with torch.cuda.device(q.device):
        stream = torch.cuda.current_stream().cuda_stream
        qs, ks, vs = q.stride()[:3], k.stride()[:3], v.stride()[:3]
        ws, ws_bytes = None, 0
        if hd % 8 != 0 or not _tma_strides(qs, ks, vs) or (q.data_ptr() | k.data_ptr() | v.data_ptr()) & 15:
            ws_bytes = _ffi.prefill_workspace_bytes(N, Tq, Tk, H, G, hd, qs, ks, vs, q.data_ptr(), k.data_ptr(), v.data_ptr())
            if ws_bytes:
                ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=q.device)
        _ffi.prefill_gather(q.data_ptr(), k.data_ptr(), v.data_ptr(), list(peer_ptrs), rank, seq_offset, head_offset,
                            N_total, H_total, qv.data_ptr() if qv is not None else None,
                            kv.data_ptr() if kv is not None else None, N, Tq, Tk, H, G, hd, qs, ks, vs,
                            out.stride()[:3], scale, causal, left, right, stream,
                            ws.data_ptr() if ws is not None else None, ws_bytes, logit_bound)


def _tma_strides(*stride_sets) -> bool:
    return all(s % 8 == 0 for ss in stride_sets for s in ss)


@gqa_swa_prefill.register_fake
def _(q, k, v, q_valid, k_valid, scale, causal, left, right, kernel=0, logit_bound=0.0):
    return q.new_empty(q.shape, dtype=torch.bfloat16)


# Split-K scratch of the decode kernels.  The C-ABI wants its head (the split counters) ZERO on entry and leaves it
# zero on exit, so in eager mode one zero-filled buffer per (device, stream, geometry) is created once and reused by
# the stream-ordered calls that share it — no memset per step.  Rules that keep that invariant sound:
#   * while a CUDA graph is being captured the cache is neither used nor touched: the call gets a fresh torch.zeros
#     from the graph's own pool (the memset becomes a graph node, the buffer lives as long as the graph);
#   * eviction is LRU, one entry at a time (never a wholesale clear);
#   * a failed decode call drops its entry (the counters may be dirty); `reset_decode_workspaces()` drops all of them,
#     e.g. after a device-side fault.
_DECODE_WS: "collections.OrderedDict" = collections.OrderedDict()
_DECODE_WS_MAX = 64


def reset_decode_workspaces() -> None:
    """Forget every cached decode workspace (they are re-created zero-filled on next use)."""
    _DECODE_WS.clear()


_DECODE_WS_BYTES: dict = {}   # (device index, geometry) -> bytes: the C-ABI query is a ctypes call per decode step otherwise


def _decode_workspace(device, stream: int, B: int, H: int, G: int, hd: int, S_max: int, left: int):
    gkey = (device.index, B, H, G, hd, S_max, left)
    nbytes = _DECODE_WS_BYTES.get(gkey)
    if nbytes is None:
        nbytes = _DECODE_WS_BYTES[gkey] = max(_ffi.decode_workspace_bytes(B, H, G, hd, S_max, left), 16)
    if torch.cuda.is_current_stream_capturing():
        return torch.zeros((nbytes,), dtype=torch.uint8, device=device), None
    key = (device.index, stream, B, H, G, hd, S_max, left)
    ws = _DECODE_WS.get(key)
    if ws is None:
        while len(_DECODE_WS) >= _DECODE_WS_MAX:
            _DECODE_WS.popitem(last=False)
        ws = _DECODE_WS[key] = torch.zeros((nbytes,), dtype=torch.uint8, device=device)
    else:
        _DECODE_WS.move_to_end(key)
    return ws, key


@torch.library.custom_op("vats::gqa_swa_decode", mutates_args=(), device_types="cuda")
def gqa_swa_decode(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, seq_lens: torch.Tensor,
                   scale: float, left: int) -> torch.Tensor:
    """q [B,H,hd], caches [B,S_max,G,hd] bf16, seq_lens [B] int32 (valid tokens incl. the new one) -> o [B,H,hd]."""
    _require_cuda_bf16("q", q)
    _require_cuda_bf16("k_cache", k_cache)
    _require_cuda_bf16("v_cache", v_cache)
    if q.dim() != 3 or k_cache.dim() != 4 or v_cache.shape != k_cache.shape:
        raise ValueError("q must be [B,H,hd]; k_cache and v_cache must be [B,S_max,G,hd] with equal shapes")
    B, H, hd = q.shape
    Bc, S_max, G, hdc = k_cache.shape
    if Bc != B or hdc != hd:
        raise ValueError(f"shape mismatch: q {tuple(q.shape)} cache {tuple(k_cache.shape)}")
    if G <= 0 or H % G != 0:
        raise ValueError(f"num_heads ({H}) must be divisible by query_groups ({G})")
    if seq_lens.shape != (B,) or seq_lens.dtype != torch.int32 or seq_lens.device != q.device:
        raise ValueError("seq_lens must be an int32 tensor of shape [B] on the same device")
    q = _rowmajor_last(q)
    if k_cache.stride(-1) != 1 or v_cache.stride(-1) != 1:
        raise ValueError("KV cache must have a contiguous head_dim axis (refusing to copy a cache)")
    seq_lens = seq_lens.contiguous()
    o = torch.empty((B, H, hd), dtype=torch.bfloat16, device=q.device)
    if o.numel() == 0:
        return o
    with torch.cuda.device(q.device):
        stream = torch.cuda.current_stream().cuda_stream
        ws, key = _decode_workspace(q.device, stream, B, H, G, hd, S_max, left)
        try:
            _ffi.decode(q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), o.data_ptr(), seq_lens.data_ptr(),
                        B, H, G, hd, S_max, q.stride()[:2], k_cache.stride()[:3], v_cache.stride()[:3], o.stride()[:2],
                        scale, left, ws.data_ptr(), ws.numel(), stream)
        except _ffi.VatsAttnError:
            _DECODE_WS.pop(key, None)   # its counters may be dirty: never reuse it
            raise
    return o


@gqa_swa_decode.register_fake
def _(q, k_cache, v_cache, seq_lens, scale, left):
    return q.new_empty(q.shape, dtype=torch.bfloat16)


@torch.library.custom_op("vats::decode_prepare", mutates_args=("k_cache", "v_cache"), device_types="cuda")
def decode_prepare(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                   seq_lens: torch.Tensor, cos: Optional[torch.Tensor], sin: Optional[torch.Tensor], qk_norm: bool,
                   eps: float) -> torch.Tensor:
    """Fused pre-core step of one cached decode token: qk L2-norm + RoPE at position seq_lens-1 + bf16 rounding + append
    of k, v to the caches (in place).  q [B,H,hd], k/v [B,G,hd] (bf16 or fp32), caches [B,S_max,G,hd] bf16, seq_lens [B]
    int32 (length including the new token), cos/sin [positions, hd/2] fp32 or None.  Returns the rotated q (bf16)."""
    if q.dim() != 3 or k.dim() != 3 or v.shape != k.shape:
        raise ValueError("q must be [B,H,hd]; k and v must be [B,G,hd] with equal shapes")
    if q.dtype not in (torch.bfloat16, torch.float32) or k.dtype != q.dtype or v.dtype != q.dtype:
        raise ValueError("q, k, v must share one dtype, bf16 or fp32")
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        raise RuntimeError("decode_prepare needs CUDA tensors (no CPU fallback)")
    _require_cuda_bf16("k_cache", k_cache)
    _require_cuda_bf16("v_cache", v_cache)
    B, H, hd = q.shape
    G = k.size(1)
    if k_cache.dim() != 4 or v_cache.shape != k_cache.shape or k_cache.size(0) != B or k_cache.size(2) != G or \
            k_cache.size(3) != hd or k.size(0) != B or k.size(2) != hd:
        raise ValueError(f"shape mismatch: q {tuple(q.shape)} k {tuple(k.shape)} cache {tuple(k_cache.shape)}")
    if k_cache.stride(-1) != 1 or v_cache.stride(-1) != 1:
        raise ValueError("KV cache must have a contiguous head_dim axis")
    if seq_lens.shape != (B,) or seq_lens.dtype != torch.int32 or seq_lens.device != q.device:
        raise ValueError("seq_lens must be an int32 tensor of shape [B] on the same device")
    if (cos is None) != (sin is None):
        raise ValueError("cos and sin must both be given or both be None")
    if cos is not None:
        if cos.dtype != torch.float32 or sin.dtype != torch.float32 or cos.shape != sin.shape or cos.dim() != 2 or \
                cos.size(1) != hd // 2 or hd % 2 != 0 or cos.size(0) < k_cache.size(1):
            raise ValueError("cos / sin must be fp32 [>= S_max, hd/2] tables")
        cos, sin = cos.contiguous(), sin.contiguous()
    q, k, v = _rowmajor_last(q), _rowmajor_last(k), _rowmajor_last(v)
    seq_lens = seq_lens.contiguous()
    q_out = torch.empty((B, H, hd), dtype=torch.bfloat16, device=q.device)
    if q_out.numel() == 0:
        return q_out
    with torch.cuda.device(q.device):
        stream = torch.cuda.current_stream().cuda_stream
        _ffi.decode_prepare(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.dtype == torch.float32, q_out.data_ptr(),
                            k_cache.data_ptr(), v_cache.data_ptr(), seq_lens.data_ptr(),
                            cos.data_ptr() if cos is not None else None, sin.data_ptr() if sin is not None else None,
                            B, H, G, hd, k_cache.size(1), q.stride()[:2], k.stride()[:2], v.stride()[:2],
                            q_out.stride()[:2], k_cache.stride()[:3], v_cache.stride()[:3], qk_norm, eps, stream)
    return q_out


@decode_prepare.register_fake
def _(q, k, v, k_cache, v_cache, seq_lens, cos, sin, qk_norm, eps):
    return q.new_empty(q.shape, dtype=torch.bfloat16)


@torch.library.custom_op("vats::prefill_prepare", mutates_args=(), device_types="cuda")
def prefill_prepare(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cos: Optional[torch.Tensor],
                    sin: Optional[torch.Tensor], pos0: int, qk_norm: bool, eps: float) -> List[torch.Tensor]:
    """Fused pre-core producers of a prefill chunk: qk L2-norm + 1-D RoPE at positions pos0 .. pos0+T-1 + bf16 rounding,
    written in the TMA-addressable layout (head stride rounded up to 8 elements; sequences of <= 32 tokens stay
    dense for the short-sequence kernel).  q [N,T,H,hd], k/v [N,T,G,hd] (bf16 or fp32, any strides with a contiguous
    head_dim), cos/sin [>= pos0+T, hd/2] fp32 or None.  Returns the three padded buffers [N,T,heads,hd_pad]; slice
    `[..., :hd]` to get q', k', v' (`prefill_prepare_views` does)."""
    if q.dim() != 4 or k.dim() != 4 or v.shape != k.shape or q.shape[:2] != k.shape[:2] or q.size(3) != k.size(3):
        raise ValueError("q must be [N,T,H,hd]; k and v must be [N,T,G,hd]")
    if q.dtype not in (torch.bfloat16, torch.float32) or k.dtype != q.dtype or v.dtype != q.dtype:
        raise ValueError("q, k, v must share one dtype, bf16 or fp32")
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        raise RuntimeError("prefill_prepare needs CUDA tensors (no CPU fallback)")
    N, T, H, hd = q.shape
    G = k.size(2)
    if (cos is None) != (sin is None):
        raise ValueError("cos and sin must both be given or both be None")
    if cos is not None:
        if cos.dtype != torch.float32 or sin.dtype != torch.float32 or cos.shape != sin.shape or cos.dim() != 2 or \
                cos.size(1) != hd // 2 or hd % 2 != 0 or cos.size(0) < pos0 + T:
            raise ValueError("cos / sin must be fp32 [>= pos0 + T, hd/2] tables")
        cos, sin = cos.contiguous(), sin.contiguous()
    q, k, v = _rowmajor_last(q), _rowmajor_last(k), _rowmajor_last(v)
    hp = hd if (hd % 8 == 0 or T <= 32) else (hd + 7) // 8 * 8
    bufs = [torch.empty((N, T, heads, hp), dtype=torch.bfloat16, device=q.device) for heads in (H, G, G)]
    if q.numel() == 0:
        return bufs
    with torch.cuda.device(q.device):
        stream = torch.cuda.current_stream().cuda_stream
        _ffi.prefill_prepare(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.dtype == torch.float32, bufs[0].data_ptr(),
                             bufs[1].data_ptr(), bufs[2].data_ptr(), cos.data_ptr() if cos is not None else None,
                             sin.data_ptr() if sin is not None else None, N, T, H, G, hd, int(pos0), q.stride()[:3],
                             k.stride()[:3], v.stride()[:3], bufs[0].stride()[:3], bufs[1].stride()[:3],
                             bufs[2].stride()[:3], qk_norm, eps, stream)
    return bufs


@prefill_prepare.register_fake
def _(q, k, v, cos, sin, pos0, qk_norm, eps):
    hd, T = q.size(3), q.size(1)
    hp = hd if (hd % 8 == 0 or T <= 32) else (hd + 7) // 8 * 8
    return [t.new_empty((*t.shape[:3], hp), dtype=torch.bfloat16) for t in (q, k, v)]


@torch.library.custom_op("vats::prefill_prepare_table", mutates_args=(), device_types="cuda")
def prefill_prepare_table(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cos: Optional[torch.Tensor],
                          sin: Optional[torch.Tensor], partner: Optional[torch.Tensor], qk_norm: bool, eps: float
                          ) -> List[torch.Tensor]:
    """Fused pre-core producers with a table-driven rotation (2-D axial / 3-D RoPE of the ViTs, any variant):
    out[c] = xn[c] * cos[tok][c] + xn[partner[c]] * sin[tok][c] after the optional qk L2-norm; v passes through; one
    rounding to bf16, written in the kernels' layout.  q [No, Ni, T, H, hd], k / v [No, Ni, T, G, hd] (bf16 or fp32, ANY
    strides with a contiguous head_dim — e.g. permuted views) -> three buffers [No*Ni, T, heads, hd_pad]
    (slice `[..., :hd]`).  cos / sin [T, hd] fp32, partner [hd] int32, or all None."""
    if q.dim() != 5 or k.dim() != 5 or v.shape != k.shape or q.shape[:3] != k.shape[:3] or q.size(4) != k.size(4):
        raise ValueError("q must be [No,Ni,T,H,hd]; k and v must be [No,Ni,T,G,hd]")
    if q.dtype not in (torch.bfloat16, torch.float32) or k.dtype != q.dtype or v.dtype != q.dtype:
        raise ValueError("q, k, v must share one dtype, bf16 or fp32")
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        raise RuntimeError("prefill_prepare_table needs CUDA tensors (no CPU fallback)")
    No, Ni, T, H, hd = q.shape
    G = k.size(3)
    given = [t is not None for t in (cos, sin, partner)]
    if any(given) and not all(given):
        raise ValueError("cos, sin and partner must all be given or all be None")
    if cos is not None:
        if cos.dtype != torch.float32 or sin.dtype != torch.float32 or partner.dtype != torch.int32 or \
                cos.shape != (T, hd) or sin.shape != (T, hd) or partner.shape != (hd,):
            raise ValueError("cos / sin must be fp32 [T, hd] tables and partner an int32 [hd] permutation")
        cos, sin, partner = cos.contiguous(), sin.contiguous(), partner.contiguous()
    fix = lambda t: t if t.stride(-1) == 1 or t.size(-1) == 1 else t.contiguous()
    q, k, v = fix(q), fix(k), fix(v)
    hp = hd if (hd % 8 == 0 or T <= 32) else (hd + 7) // 8 * 8
    bufs = [torch.empty((No * Ni, T, heads, hp), dtype=torch.bfloat16, device=q.device) for heads in (H, G, G)]
    if q.numel() == 0:
        return bufs
    with torch.cuda.device(q.device):
        stream = torch.cuda.current_stream().cuda_stream
        _ffi.prefill_prepare_table(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.dtype == torch.float32, bufs[0].data_ptr(),
                                   bufs[1].data_ptr(), bufs[2].data_ptr(), cos.data_ptr() if cos is not None else None,
                                   sin.data_ptr() if sin is not None else None,
                                   partner.data_ptr() if partner is not None else None, No, Ni, T, H, G, hd,
                                   q.stride()[:4], k.stride()[:4], v.stride()[:4], bufs[0].stride()[:3],
                                   bufs[1].stride()[:3], bufs[2].stride()[:3], qk_norm, eps, stream)
    return bufs


@prefill_prepare_table.register_fake
def _(q, k, v, cos, sin, partner, qk_norm, eps):
    hd, T = q.size(4), q.size(2)
    hp = hd if (hd % 8 == 0 or T <= 32) else (hd + 7) // 8 * 8
    n = q.size(0) * q.size(1)
    return [t.new_empty((n, T, t.size(3), hp), dtype=torch.bfloat16) for t in (q, k, v)]


def prefill_prepare_table_views(q, k, v, cos, sin, partner, qk_norm: bool, eps: float = 1e-6):
    """`vats::prefill_prepare_table`, returning q', k', v' as [..., :hd] views of the padded buffers."""
    hd = q.size(-1)
    qb, kb, vb = prefill_prepare_table(q, k, v, cos, sin, partner, qk_norm, eps)
    return qb[..., :hd], kb[..., :hd], vb[..., :hd]


def prefill_prepare_views(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cos: Optional[torch.Tensor],
                          sin: Optional[torch.Tensor], pos0: int, qk_norm: bool, eps: float = 1e-6):
    """`vats::prefill_prepare`, returning q', k', v' as [..., :hd] views of the padded buffers (what `gqa_swa_prefill`
    takes: the head stride keeps rows 16-byte aligned for TMA)."""
    hd = q.size(3)
    qb, kb, vb = prefill_prepare(q, k, v, cos, sin, pos0, qk_norm, eps)
    return qb[..., :hd], kb[..., :hd], vb[..., :hd]


def attn_mask(q_valid: Optional[torch.Tensor], k_valid: Optional[torch.Tensor], N: int, Tq: int, Tk: int,
              causal: bool, left: int, right: int, device: Optional[torch.device] = None) -> torch.Tensor:
    """The kernels' own mask predicate, materialised (uint8 [N,Tq,Tk]); used for the bit-exact mask tests."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("attn_mask runs the device predicate and needs a CUDA device")
    qv = _valid_u8("q_valid", q_valid, N, Tq, device)
    kv = _valid_u8("k_valid", k_valid, N, Tk, device)
    out = torch.empty((N, Tq, Tk), dtype=torch.uint8, device=device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        _ffi.debug_mask(out.data_ptr(), qv.data_ptr() if qv is not None else None,
                        kv.data_ptr() if kv is not None else None, N, Tq, Tk, causal, left, right, stream)
    return out


def _no_cpu(name):
    def impl(*args, **kwargs):
        raise RuntimeError(
            f"torch.ops.vats.{name} has no CPU implementation: the attention core runs only on a CUDA sm_100 device "
            "(hand-written kernels, no fallback). Move the tensors to the GPU."
        )
    return impl


gqa_swa_prefill.register_kernel("cpu")(_no_cpu("gqa_swa_prefill"))
gqa_swa_prefill_bwd.register_kernel("cpu")(_no_cpu("gqa_swa_prefill_bwd"))
gqa_swa_decode.register_kernel("cpu")(_no_cpu("gqa_swa_decode"))
decode_prepare.register_kernel("cpu")(_no_cpu("decode_prepare"))
prefill_prepare.register_kernel("cpu")(_no_cpu("prefill_prepare"))
prefill_prepare_table.register_kernel("cpu")(_no_cpu("prefill_prepare_table"))


# ---------------------------------------------------------------------------------------------------------------------
# Eager fast path.  A call through the torch.library dispatcher costs ~40 us of host time per op (measured: 38-48 us for
# vats::gqa_swa_decode, against ~2 us for the implementation function itself) — more than the decode kernel's launch and
# a third of a decode step that runs two ops (eager end-to-end 3.5 vs 4.8 TB/s under CUDA-graph replay).  The dispatcher
# is needed for tracing (torch.compile / FakeTensor), functorch transforms, tensor subclasses, autograd and the CPU error
# path; for plain CUDA tensors with nothing to differentiate the public names below call the implementation directly.
# `torch.ops.vats.*` keeps going through the dispatcher.
def _direct_ok(*tensors) -> bool:
    if torch._C._len_torch_dispatch_stack() or torch._C._len_torch_function_stack() or torch.compiler.is_compiling():
        return False
    if torch._C._functorch.peek_interpreter_stack() is not None:
        return False
    for t in tensors:
        if t is not None and (type(t) is not torch.Tensor or not t.is_cuda):
            return False
    return True


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


_op_gqa_swa_prefill, _op_gqa_swa_decode, _op_decode_prepare = gqa_swa_prefill, gqa_swa_decode, decode_prepare
_op_prefill_prepare, _op_prefill_prepare_table = prefill_prepare, prefill_prepare_table


def gqa_swa_prefill(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, q_valid: Optional[torch.Tensor],   # noqa: F811
                    k_valid: Optional[torch.Tensor], scale: float, causal: bool, left: int, right: int,
                    kernel: int = 0, logit_bound: float = 0.0) -> torch.Tensor:
    if _direct_ok(q, k, v, q_valid, k_valid) and not _needs_grad(q, k, v):
        return _op_gqa_swa_prefill._init_fn(q, k, v, q_valid, k_valid, scale, causal, left, right, kernel, logit_bound)
    return _op_gqa_swa_prefill(q, k, v, q_valid, k_valid, scale, causal, left, right, kernel, logit_bound)


def gqa_swa_decode(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, seq_lens: torch.Tensor,   # noqa: F811
                   scale: float, left: int) -> torch.Tensor:
    if _direct_ok(q, k_cache, v_cache, seq_lens):
        return _op_gqa_swa_decode._init_fn(q, k_cache, v_cache, seq_lens, scale, left)
    return _op_gqa_swa_decode(q, k_cache, v_cache, seq_lens, scale, left)


def decode_prepare(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,   # noqa: F811
                   seq_lens: torch.Tensor, cos: Optional[torch.Tensor], sin: Optional[torch.Tensor], qk_norm: bool,
                   eps: float) -> torch.Tensor:
    if _direct_ok(q, k, v, k_cache, v_cache, seq_lens, cos, sin) and not _needs_grad(q, k, v):
        return _op_decode_prepare._init_fn(q, k, v, k_cache, v_cache, seq_lens, cos, sin, qk_norm, eps)
    return _op_decode_prepare(q, k, v, k_cache, v_cache, seq_lens, cos, sin, qk_norm, eps)


def prefill_prepare(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cos: Optional[torch.Tensor],   # noqa: F811
                    sin: Optional[torch.Tensor], pos0: int, qk_norm: bool, eps: float) -> List[torch.Tensor]:
    if _direct_ok(q, k, v, cos, sin) and not _needs_grad(q, k, v):
        return _op_prefill_prepare._init_fn(q, k, v, cos, sin, pos0, qk_norm, eps)
    return _op_prefill_prepare(q, k, v, cos, sin, pos0, qk_norm, eps)


def prefill_prepare_table(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cos: Optional[torch.Tensor],   # noqa: F811
                          sin: Optional[torch.Tensor], partner: Optional[torch.Tensor], qk_norm: bool,
                          eps: float = 1e-6) -> List[torch.Tensor]:
    if _direct_ok(q, k, v, cos, sin, partner) and not _needs_grad(q, k, v):
        return _op_prefill_prepare_table._init_fn(q, k, v, cos, sin, partner, qk_norm, eps)
    return _op_prefill_prepare_table(q, k, v, cos, sin, partner, qk_norm, eps)
