"""softmax(scale * Q K^T | mask) V restated on the CPU (TEST INFRASTRUCTURE).

Restates the third-party call the reference makes — torch.nn.functional.scaled_dot_product_attention with a boolean
mask (reference src/optimized_attention.py:709-714; vit_2d/optimized_attention.py:396-402;
vit_3d/optimized_attention.py:302-307) — including the K/V head expansion that precedes it
(`extend_kv_heads`, reference utils/attention_utils.py:7-27: repeat_interleave, i.e. query head h reads KV head
h // (H/G)) and the convention that a query row with no allowed key yields zeros (torch >= 2.5 SDPA and flash-attn).
"""
from __future__ import annotations

from typing import Optional

import torch


def expand_kv(x: torch.Tensor, H: int) -> torch.Tensor:
    """[N, T, G, hd] -> [N, T, H, hd] exactly as reference utils/attention_utils.py:27 does."""
    G = x.size(2)
    if G == H:
        return x
    assert H % G == 0
    return x.repeat_interleave(H // G, dim=2)


def sdpa_explicit(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: Optional[torch.Tensor], scale: float,
                  dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """q [N,Tq,H,hd], k/v [N,Tk,G,hd], mask bool [N,Tq,Tk] (True = attend) or None -> o [N,Tq,H,hd] in `dtype`."""
    N, Tq, H, hd = q.shape
    qf = q.to(dtype).permute(0, 2, 1, 3)                      # [N,H,Tq,hd]
    kf = expand_kv(k, H).to(dtype).permute(0, 2, 1, 3)        # [N,H,Tk,hd]
    vf = expand_kv(v, H).to(dtype).permute(0, 2, 1, 3)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale        # [N,H,Tq,Tk]
    if mask is not None:
        m = mask[:, None, :, :]
        s = s.masked_fill(~m, float("-inf"))
        dead = ~m.any(dim=-1, keepdim=True)                   # rows with nothing to attend
        s = s.masked_fill(dead, 0.0)
        p = torch.softmax(s, dim=-1)
        p = p.masked_fill(dead, 0.0)
    else:
        p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, vf)                                   # [N,H,Tq,hd]
    return o.permute(0, 2, 1, 3).contiguous()


def decode_explicit(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, seq_lens: torch.Tensor,
                    scale: float, left: int, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Intended KV-cache step (reference src/optimized_attention.py:508-516 + 709-714, never executed by the reference;
    contract in SURVEY.md §8a-3): q [B,H,hd]; caches [B,S_max,G,hd]; the query of sequence b sits at position
    L-1 (L = seq_lens[b]) and attends keys max(0, L-1-left) .. L-1.  Returns [B,H,hd]."""
    B, H, hd = q.shape
    out = torch.zeros(B, H, hd, dtype=dtype)
    for b in range(B):
        L = int(seq_lens[b])
        if L <= 0:
            continue
        lo = 0 if left < 0 else max(0, L - 1 - left)
        kb = k_cache[b:b + 1, lo:L]
        vb = v_cache[b:b + 1, lo:L]
        out[b] = sdpa_explicit(q[b:b + 1, None], kb, vb, None, scale, dtype)[0, 0]
    return out


def sdpa_rows(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, rows: torch.Tensor, scale: float, causal: bool,
              left: int, right: int, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Reference for a SAMPLE of query rows of one problem (used by bench.py to report max-abs / relative error at
    full size, where the whole [Tq, Tk] score matrix would not fit): q [N,Tq,H,hd], k/v [N,Tk,G,hd] (CPU tensors),
    `rows` int64 [R] query indices -> [N, R, H, hd].  Mask predicate of oracle/mask.py, bottom-right aligned; only the
    key range the sampled rows can see is touched."""
    N, Tq, H, hd = q.shape
    Tk = k.size(1)
    off = Tk - Tq
    rows = rows.to(torch.int64)
    i = rows[:, None] + off
    lo = int(max(0, (rows.min().item() + off - left) if left >= 0 else 0))
    hi = Tk - 1
    if causal:
        hi = min(hi, int(rows.max().item() + off))
    if right >= 0:
        hi = min(hi, int(rows.max().item() + off + right))
    if hi < lo:
        return torch.zeros(N, rows.numel(), H, hd, dtype=dtype)
    j = torch.arange(lo, hi + 1, dtype=torch.int64)[None, :]
    ok = torch.ones(rows.numel(), hi - lo + 1, dtype=torch.bool)
    if causal:
        ok &= j <= i
    if left >= 0:
        ok &= j >= i - left
    if right >= 0:
        ok &= j <= i + right
    mask = ok[None].expand(N, -1, -1)
    return sdpa_explicit(q[:, rows], k[:, lo:hi + 1], v[:, lo:hi + 1], mask, scale, dtype)
