"""The reference's CPU attention path, restated for timing (TEST / BENCH INFRASTRUCTURE — never the product path).

What the reference executes on a CPU host for each workload is: expand K/V to H heads with `repeat_interleave`
(utils/attention_utils.py:27), transpose to [B,H,T,hd], build a boolean mask when one is needed, and call
`torch.nn.functional.scaled_dot_product_attention` (src/optimized_attention.py:709-714; vit_2d :396-402;
vit_3d :302-307).  /root/reference does not exist on the GPU box, so `bench.py` times this port ("kind": "port")
with all host threads.  The window mask is passed explicitly so the CPU does the same algorithmic work as the
kernel (the reference itself would silently drop the window and attend every key).
"""
from __future__ import annotations

import time
from typing import Optional

import torch
import torch.nn.functional as F

from .mask import mask_predicate


def reference_core_cpu(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float, causal: bool, left: int,
                       right: int, q_valid: Optional[torch.Tensor] = None, k_valid: Optional[torch.Tensor] = None
                       ) -> torch.Tensor:
    """q [N,Tq,H,hd], k/v [N,Tk,G,hd] on the CPU -> [N,Tq,H,hd]; the reference's own sequence of calls."""
    N, Tq, H, hd = q.shape
    G = k.size(2)
    if G != H:
        k = k.repeat_interleave(H // G, dim=2)
        v = v.repeat_interleave(H // G, dim=2)
    qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    Tk = kt.size(2)
    plain_causal = causal and left < 0 and q_valid is None and k_valid is None and Tq == Tk
    if plain_causal:
        out = F.scaled_dot_product_attention(qt, kt, vt, is_causal=True, scale=scale)
    elif not causal and left < 0 and right < 0 and q_valid is None and k_valid is None:
        out = F.scaled_dot_product_attention(qt, kt, vt, is_causal=False, scale=scale)
    else:
        mask = mask_predicate(N, Tq, Tk, causal, left, right, q_valid, k_valid)[:, None]
        out = F.scaled_dot_product_attention(qt, kt, vt, attn_mask=mask, is_causal=False, scale=scale)
    return out.transpose(1, 2).contiguous()


def reference_decode_cpu(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, seq_len: int, scale: float,
                         left: int) -> torch.Tensor:
    """Single-query step over a cache with uniform length `seq_len`: q [B,H,hd] -> [B,H,hd].
    The reference-style path: slice the cache to the window, expand heads, SDPA with Tq = 1 (no mask needed once
    the window has been sliced)."""
    lo = 0 if left < 0 else max(0, seq_len - 1 - left)
    kw = k_cache[:, lo:seq_len]
    vw = v_cache[:, lo:seq_len]
    return reference_core_cpu(q[:, None], kw, vw, scale, False, -1, -1)[:, 0]


def time_callable(fn, min_seconds: float, max_calls: int = 1000):
    """Run `fn` repeatedly for about `min_seconds`; returns (calls, seconds)."""
    fn()  # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    calls = 0
    while True:
        fn()
        calls += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or calls >= max_calls:
            return calls, dt
