"""Mask predicate of the attention core, restated with plain torch integer ops (TEST INFRASTRUCTURE).

Follows, clause by clause:
  causal, and `if causal: right_window = 0`   reference src/optimized_attention.py:519-520, 632-634
  window_size=(left, right) passed verbatim   reference src/optimized_attention.py:634,
                                              vit_2d/optimized_attention.py:337, vit_3d/optimized_attention.py:162
      meaning "query i attends keys [i + off - left, i + off + right], -1 = unlimited, off = Tk - Tq"
      (flash-attn 2.8.3 docstring, flash_attn_interface.py:1232-1233 — the third-party call the reference makes)
  query-row padding                           reference src/optimized_attention.py:673-675
  key padding                                 reference vit_3d/optimized_attention.py:276-277
"""
from __future__ import annotations

from typing import Optional

import torch


def mask_predicate(N: int, Tq: int, Tk: int, causal: bool, left: int, right: int,
                   q_valid: Optional[torch.Tensor] = None, k_valid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bool [N, Tq, Tk]: allowed(n, i, j)."""
    i = torch.arange(Tq, dtype=torch.int64)[:, None]
    j = torch.arange(Tk, dtype=torch.int64)[None, :]
    off = Tk - Tq
    ok = torch.ones(Tq, Tk, dtype=torch.bool)
    if causal:
        ok &= j <= i + off
    if left >= 0:
        ok &= j >= i + off - left
    if right >= 0:
        ok &= j <= i + off + right
    ok = ok[None].expand(N, Tq, Tk).clone()
    if q_valid is not None:
        ok &= q_valid.bool().cpu()[:, :, None]
    if k_valid is not None:
        ok &= k_valid.bool().cpu()[:, None, :]
    return ok
