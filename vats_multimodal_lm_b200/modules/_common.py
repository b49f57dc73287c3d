"""Pieces shared by the three drop-in attention modules: projections, QK L2-norm, the bridge into the custom op.

Mirrors reference utils/attention_utils.py (setup_projections :29-78, apply_qk_norm :80-102).  `extend_kv_heads`
(:7-27) has no counterpart on purpose: the op indexes K/V head h // (H/G) inside the kernel, so the expanded copy
is never written.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

# How the window arguments reach the kernel.
#   "swa"            : (left, right) are honoured — the behaviour the reference intends (its FA2 call,
#                      src/optimized_attention.py:628-635) and the north-star default.
#   "reference_sdpa" : windows are dropped, exactly like the only path the reference can execute
#                      (src/optimized_attention.py:709-714 passes no window) — used for module-level parity tests.
WINDOW_MODES = ("swa", "reference_sdpa")
_default_window_mode = "swa"


def set_default_window_mode(mode: str) -> None:
    global _default_window_mode
    if mode not in WINDOW_MODES:
        raise ValueError(f"window mode must be one of {WINDOW_MODES}, got {mode!r}")
    _default_window_mode = mode


def get_default_window_mode() -> str:
    return _default_window_mode


def setup_projections(d_model: int, num_heads: int, head_dim: int, use_fused_proj: bool, use_gqa: bool,
                      use_proj_bias: bool, query_groups: Optional[int] = None
                      ) -> Union[Tuple[nn.Linear, nn.Linear], Tuple[nn.Linear, nn.Linear, nn.Linear, nn.Linear]]:
    """Same parameter shapes (and creation order, so seeded inits agree) as reference utils/attention_utils.py:29-78."""
    if use_gqa and query_groups is None:
        raise AssertionError("Must have query groups for GQA.")
    kv_width = query_groups * head_dim if use_gqa else num_heads * head_dim
    o_proj = nn.Linear(d_model, d_model, bias=use_proj_bias)
    if use_fused_proj:
        qkv_proj = nn.Linear(d_model, num_heads * head_dim + 2 * kv_width if use_gqa else 3 * d_model,
                             bias=use_proj_bias)
        return qkv_proj, o_proj
    q_proj = nn.Linear(d_model, num_heads * head_dim, bias=use_proj_bias)
    k_proj = nn.Linear(d_model, kv_width, bias=use_proj_bias)
    v_proj = nn.Linear(d_model, kv_width, bias=use_proj_bias)
    return q_proj, k_proj, v_proj, o_proj


# |<q, k>| behind `apply_qk_norm` (unit vectors; RoPE is a rotation): what the modules promise the kernels as
# `logit_bound` when use_qk_norm is set, so the row-maximum pass of the softmax is skipped
QK_NORM_LOGIT_BOUND = 1.0


def apply_qk_norm(query: torch.Tensor, key: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """L2-normalise q and k over head_dim (reference utils/attention_utils.py:99-102, eps = 1e-6)."""
    return F.normalize(query, p=2, dim=-1, eps=1e-6), F.normalize(key, p=2, dim=-1, eps=1e-6)


def _to_kernel_layout(x: torch.Tensor, pad: bool = True) -> torch.Tensor:
    """bf16 copy of a [N, T, heads, hd] tensor in a TMA-addressable layout.

    The cast to bf16 writes a new tensor anyway; for head dims that are not a multiple of 8 (60, 66) it is written
    into a buffer whose head stride is rounded up to 8 elements and returned as a [..., :hd] view.  Rows then start
    16-byte aligned, so the tensor-core kernel can fetch them with TMA instead of its 32-bit staging loads; the
    padding is never read (the tensor map's inner extent is hd) and costs no extra pass over the data.
    """
    hd = x.size(-1)
    if hd % 8 == 0 or not pad:
        return x.to(torch.bfloat16).contiguous()
    hd8 = (hd + 7) // 8 * 8
    buf = torch.empty(*x.shape[:-1], hd8, dtype=torch.bfloat16, device=x.device)
    view = buf[..., :hd]
    view.copy_(x)
    return view


def attention_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, scale: float, causal: bool, left: int,
                   right: int, q_valid: Optional[torch.Tensor] = None, k_valid: Optional[torch.Tensor] = None,
                   out_dtype: Optional[torch.dtype] = None, logit_bound: float = 0.0) -> torch.Tensor:
    """q [N,Tq,H,hd], k/v [N,Tk,G,hd] (G heads, un-expanded) -> [N,Tq,H,hd].

    The one call that replaces `F.scaled_dot_product_attention` at the reference's three call sites.  Inputs are cast
    to bf16 (the kernels' arithmetic type: bf16 operands, fp32 accumulation) and the result is cast back.
    """
    out_dtype = out_dtype or q.dtype
    # sequences of at most 32 keys go to the one-CTA-per-sequence kernel, which streams dense [T, heads, hd] blocks
    # with bulk copies: no head-stride padding there
    pad = k.size(1) > 32
    o = ops.gqa_swa_prefill(_to_kernel_layout(q, pad), _to_kernel_layout(k, pad), _to_kernel_layout(v, pad), q_valid,
                            k_valid, float(scale), bool(causal), int(left), int(right), 0, float(logit_bound))
    return o.to(out_dtype)
