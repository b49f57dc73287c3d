"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's pre-core producers for one decode token.

Follows, for the new token of every sequence (position p_b = seq_lens[b] - 1):
    apply_qk_norm      utils/attention_utils.py:80-102        F.normalize(x, p=2, dim=-1, eps=1e-6)
    RoPE._apply_rope   src/optimized_attention.py:112-143     interleaved pairs (2i, 2i+1), cos/sin = cos/sin(p * inv_freq)
    RoPE.__init__      src/optimized_attention.py:22-44       inv_freq = theta ** (-2i / hd), fp32
    order              src/optimized_attention.py:467-474     norm first, then RoPE; v untouched
    KVCache.update     src/optimized_attention.py:224-257     (intended contract) k, v written at position p_b
Pinned against the reference itself by tests/golden/prepare_*.pt (oracle/gen_golden.py: prepare_case).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def rope_tables(hd: int, theta: float, positions: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos / sin caches exactly as RoPE._update_cache builds them (fp32, [positions, hd/2])."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    pos = torch.arange(positions, dtype=torch.float32)
    freqs = torch.outer(pos, inv_freq)
    return torch.cos(freqs), torch.sin(freqs)


def _rope_at(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [B, heads, hd]; cos/sin [B, hd/2] (row of each sequence's position)."""
    x1, x2 = x[..., ::2], x[..., 1::2]
    c, s = cos[:, None, :], sin[:, None, :]
    return torch.stack([x1 * c - x2 * s, x1 * s + x2 * c], dim=-1).flatten(-2)


def decode_prepare_explicit(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, k_cache: torch.Tensor,
                            v_cache: torch.Tensor, seq_lens: torch.Tensor, cos: Optional[torch.Tensor],
                            sin: Optional[torch.Tensor], qk_norm: bool, eps: float = 1e-6):
    """fp32 reference of torch.ops.vats.decode_prepare.  Returns (q_out fp32, k_cache', v_cache') with the caches updated
    out of place (fp32 values of what the kernel rounds to bf16)."""
    q, k, v = q.float(), k.float(), v.float()
    if qk_norm:
        q = torch.nn.functional.normalize(q, p=2, dim=-1, eps=eps)
        k = torch.nn.functional.normalize(k, p=2, dim=-1, eps=eps)
    pos = (seq_lens.long() - 1).clamp(min=0)
    if cos is not None:
        q = _rope_at(q, cos[pos], sin[pos])
        k = _rope_at(k, cos[pos], sin[pos])
    kc, vc = k_cache.float().clone(), v_cache.float().clone()
    for b in range(q.size(0)):
        L = int(seq_lens[b])
        if 0 < L <= kc.size(1):
            kc[b, L - 1] = k[b]
            vc[b, L - 1] = v[b]
    return q, kc, vc


def prefill_prepare_explicit(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cos: Optional[torch.Tensor],
                             sin: Optional[torch.Tensor], pos0: int, qk_norm: bool, eps: float = 1e-6):
    """fp32 reference of torch.ops.vats.prefill_prepare: q, k [N,T,heads,hd] normalised and rotated at positions
    pos0 .. pos0+T-1 (the reference's RoPE.forward with the drop-in's position offset), v unchanged."""
    q, k, v = q.float(), k.float(), v.float()
    if qk_norm:
        q = torch.nn.functional.normalize(q, p=2, dim=-1, eps=eps)
        k = torch.nn.functional.normalize(k, p=2, dim=-1, eps=eps)
    if cos is not None:
        T = q.size(1)
        c = cos[pos0:pos0 + T][None, :, None, :]
        s = sin[pos0:pos0 + T][None, :, None, :]

        def rot(x):
            x1, x2 = x[..., ::2], x[..., 1::2]
            return torch.stack([x1 * c - x2 * s, x1 * s + x2 * c], dim=-1).flatten(-2)

        q, k = rot(q), rot(k)
    return q, k, v
