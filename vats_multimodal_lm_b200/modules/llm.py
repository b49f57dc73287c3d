"""Drop-in replacements for reference src/optimized_attention.py: RoPE (:18-166), KVCache (:169-287),
Attention (:290-729), AttentionBlock (:732-817).

Constructor and forward signatures, parameter / buffer names (`w_qkv`, `w_o`, `w_q/w_k/w_v`, `rope.inv_freq`,
`rope.cos_cache`, `rope.sin_cache`) and error conventions follow the reference so call sites such as
src/transformers/nlp/model.py:49-59 and reference state-dicts keep working.  The attention arithmetic itself goes
through `torch.ops.vats.gqa_swa_prefill` / `gqa_swa_decode` (hand-written sm_100a kernels).

Deliberate differences from the reference's *executable* path (see DESIGN.md "quirk ledger"):
  * sliding windows are honoured (`window_mode="swa"`); `window_mode="reference_sdpa"` reproduces the reference,
    whose SDPA call drops them (src/optimized_attention.py:709-714);
  * the KV cache actually works: new k/v are appended at the layer's current length and the query attends the cache
    (the reference's branch at :508-516 is never entered); the cache stores the G un-expanded KV heads;
  * RoPE positions continue from the cache length during cached decoding (the reference restarts at 0, :59);
  * `cache_out` holds the new [B, T, G, hd] k/v (the reference slices an already transposed tensor, :726).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ._common import apply_qk_norm, attention_core, get_default_window_mode, setup_projections, WINDOW_MODES


class RMSNorm(nn.Module):
    """x / sqrt(mean(x^2) + eps) * weight (reference src/rms_norm.py:8-36)."""

    def __init__(self, d_model: int, eps: float):
        super().__init__()
        self.d_model = d_model
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d_model))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.size(-1) == self.d_model, f"expected {self.d_model}, got {x.size(-1)}"
        return self.weight * (x / torch.sqrt(torch.mean(x ** 2, keepdim=True, dim=-1) + self.eps))


class RoPE(nn.Module):
    """Interleaved-pair rotary embedding over [B, T, heads, head_dim] (reference src/optimized_attention.py:18-166).

    `offset` (not in the reference) shifts the positions so cached decoding rotates the new token by its true index.
    """

    def __init__(self, head_dim: int, theta: float):
        super().__init__()
        if head_dim % 2 != 0:
            raise ValueError(f"head_dim ({head_dim}) must be divisible by 2 for even splitting.")
        self.head_dim = head_dim
        inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
        self.register_buffer("inv_freq", inv_freq)
        self.register_buffer("cos_cache", torch.empty(0))
        self.register_buffer("sin_cache", torch.empty(0))
        self.cached_seq_len = 0

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # reference checkpoints carry cos/sin caches of whatever length was last used
        # (src/transformers/nlp/inference/interactive_generation.py strips them before loading); accept any shape.
        for name in ("cos_cache", "sin_cache"):
            t = state_dict.pop(prefix + name, None)
            if t is not None and t.numel() > 0:
                setattr(self, name, t.to(self.inv_freq.device))
                self.cached_seq_len = t.size(0)
        state_dict.setdefault(prefix + "cos_cache", self.cos_cache)
        state_dict.setdefault(prefix + "sin_cache", self.sin_cache)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def _update_cache(self, seq_len: int) -> None:
        pos = torch.arange(seq_len, device=self.inv_freq.device, dtype=self.inv_freq.dtype)
        freqs = torch.outer(pos, self.inv_freq)
        self.cos_cache = torch.cos(freqs)
        self.sin_cache = torch.sin(freqs)
        self.cached_seq_len = seq_len

    def get_cos_sin_cache(self, seq_len: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if seq_len > self.cached_seq_len or self.cos_cache.device != self.inv_freq.device:
            self._update_cache(seq_len)
        return self.cos_cache[:seq_len], self.sin_cache[:seq_len]

    def forward(self, x: torch.Tensor, offset: int = 0) -> torch.Tensor:
        if x.dim() != 4:
            raise ValueError(f"x must have 4 dimensions, got {x.dim()}")
        T = x.size(1)
        cos, sin = self.get_cos_sin_cache(offset + T)
        cos = cos[offset:offset + T][None, :, None, :].to(x.dtype)
        sin = sin[offset:offset + T][None, :, None, :].to(x.dtype)
        x1, x2 = x[..., ::2], x[..., 1::2]
        return torch.stack([x1 * cos - x2 * sin, x1 * sin + x2 * cos], dim=-1).flatten(-2)


class KVCache:
    """Per-layer K/V store with the reference's API (src/optimized_attention.py:169-287): `initialize`, `update`,
    `get`, `reset`, `.cache[layer]['k'|'v']` of shape [B, max_seq_len, num_heads, head_dim], `.current_seq_len`.

    Unlike the reference (one counter advanced by every layer's `update`, :257) the length is tracked per layer;
    `current_seq_len` reports the longest layer, which is what a caller polling it after a full forward expects.
    `num_heads` is the number of heads stored — pass `query_groups` to keep the cache un-expanded (what the drop-in
    Attention writes); a cache built with the full head count also works, the extra heads are simply not used.
    """

    def __init__(self, max_batch_size: int, max_seq_len: int, num_heads: int, head_dim: int, num_layers: int,
                 dtype: torch.dtype = torch.bfloat16, device: Optional[torch.device] = None):
        self.max_batch_size = max_batch_size
        self.max_seq_len = max_seq_len
        self.num_heads = num_heads
        self.head_dim = head_dim
        self.num_layers = num_layers
        self.dtype = dtype
        self.device = device
        self.cache = None
        self.batch_size = None
        self._layer_len = None

    @property
    def current_seq_len(self) -> Optional[int]:
        return None if self._layer_len is None else max(self._layer_len)

    def layer_seq_len(self, layer_idx: int) -> int:
        return 0 if self._layer_len is None else self._layer_len[layer_idx]

    def initialize(self, batch_size: int, device: Optional[torch.device] = None) -> None:
        if batch_size > self.max_batch_size:
            raise ValueError(f"batch_size ({batch_size}) exceeds max_batch_size ({self.max_batch_size})")
        device = device or self.device
        self.batch_size = batch_size
        self._layer_len = [0] * self.num_layers
        shape = (batch_size, self.max_seq_len, self.num_heads, self.head_dim)
        self.cache = [{"k": torch.zeros(shape, dtype=self.dtype, device=device),
                       "v": torch.zeros(shape, dtype=self.dtype, device=device)} for _ in range(self.num_layers)]

    def update(self, layer_idx: int, k: torch.Tensor, v: torch.Tensor) -> None:
        """Append k, v [B, T, num_heads, head_dim] at the layer's current length (truncating at max_seq_len)."""
        if self.cache is None or self.batch_size != k.size(0):
            self.initialize(k.size(0), device=k.device)
        cur = self._layer_len[layer_idx]
        new = k.size(1)
        space = self.max_seq_len - cur
        if space <= 0:
            return
        if new > space:
            k, v, new = k[:, :space], v[:, :space], space
        self.cache[layer_idx]["k"][:, cur:cur + new] = k.to(self.cache[layer_idx]["k"].dtype)
        self.cache[layer_idx]["v"][:, cur:cur + new] = v.to(self.cache[layer_idx]["v"].dtype)
        self._layer_len[layer_idx] = cur + new

    def advance(self, layer_idx: int, new: int) -> None:
        """Record `new` tokens written into the layer's cache in place (the fused decode step appends on the device)."""
        self._layer_len[layer_idx] = min(self._layer_len[layer_idx] + new, self.max_seq_len)

    def get(self, layer_idx: int, seq_len: int) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        if self.cache is None or seq_len > self._layer_len[layer_idx]:
            return None, None
        return self.cache[layer_idx]["k"][:, :seq_len], self.cache[layer_idx]["v"][:, :seq_len]

    def reset(self) -> None:
        self.cache = None
        self.batch_size = None
        self._layer_len = None


class Attention(nn.Module):
    """GQA + causal + sliding-window attention layer (reference src/optimized_attention.py:290-729)."""

    def __init__(self, d_model: int, num_heads: int, query_groups: int, theta: float, softmax_scale: float,
                 use_proj_bias: bool = False, use_qkv_proj: bool = True, *, window_mode: Optional[str] = None):
        super().__init__()
        self.d_model = d_model
        self.num_heads = num_heads
        self.query_groups = query_groups
        self.head_dim = d_model // num_heads
        self.heads_per_group = num_heads // query_groups if query_groups else 0
        self.softmax_scale = softmax_scale
        self.use_qkv_proj = use_qkv_proj
        if window_mode is not None and window_mode not in WINDOW_MODES:
            raise ValueError(f"window_mode must be one of {WINDOW_MODES}")
        self.window_mode = window_mode
        if d_model % num_heads != 0:
            raise ValueError(f"d_model ({d_model}) must be divisible by num_heads ({num_heads})")
        if num_heads % query_groups != 0:
            raise ValueError(f"num_heads ({num_heads}) must be divisible by query_groups ({query_groups})")
        if use_qkv_proj:
            self.w_qkv, self.w_o = setup_projections(d_model, num_heads, self.head_dim, True, True, use_proj_bias,
                                                     query_groups)
        else:
            self.w_q, self.w_k, self.w_v, self.w_o = setup_projections(d_model, num_heads, self.head_dim, False, True,
                                                                       use_proj_bias, query_groups)
        self.rope = RoPE(self.head_dim, theta)

    def _windows(self, left_window: int, right_window: int, causal: bool) -> Tuple[int, int]:
        mode = self.window_mode or get_default_window_mode()
        if mode == "reference_sdpa":
            return -1, -1  # what the reference's SDPA path does with them (src/optimized_attention.py:709-714)
        if causal:
            right_window = 0  # reference src/optimized_attention.py:519-520
        return int(left_window), int(right_window)

    def forward(self, x: torch.Tensor, left_window: int, right_window: int, causal: bool = True,
                padding_mask: Optional[torch.Tensor] = None, kv_cache: Optional[KVCache] = None,
                layer_idx: Optional[int] = None, use_cache: bool = False, use_mqa: bool = False,
                use_qk_norm: bool = True) -> Tuple[torch.Tensor, Optional[Dict[str, torch.Tensor]]]:
        if x.dim() != 3 or x.size(-1) != self.d_model:
            raise ValueError(f"Expected x to have shape [B, T, d_model], got {x.shape}")
        B, T, _ = x.shape
        if T == 0:
            return torch.empty(B, 0, self.d_model, device=x.device, dtype=x.dtype), None

        H, G, hd = self.num_heads, self.query_groups, self.head_dim
        if self.use_qkv_proj:
            qkv = self.w_qkv(x)
            q, k, v = torch.split(qkv, [H * hd, G * hd, G * hd], dim=-1)
        else:
            q, k, v = self.w_q(x), self.w_k(x), self.w_v(x)
        q = q.view(B, T, H, hd)
        k = k.view(B, T, G, hd)
        v = v.view(B, T, G, hd)
        cached = bool(use_cache) and kv_cache is not None and layer_idx is not None

        # ---- single-token cached decode: qk-norm + RoPE + bf16 rounding + cache append in ONE kernel, then the
        #      split-K decode kernel (two launches per layer instead of ~10 elementwise ones + the cache write)
        if cached and T == 1 and padding_mask is None and causal and x.is_cuda and hd % 2 == 0:
            if kv_cache.cache is None or kv_cache.batch_size != B:
                kv_cache.initialize(B, device=x.device)
            past = kv_cache.layer_seq_len(layer_idx)
            k_all, v_all = kv_cache.cache[layer_idx]["k"], kv_cache.cache[layer_idx]["v"]
            if k_all.size(2) != G:
                raise ValueError(f"KVCache stores {k_all.size(2)} heads; the drop-in Attention needs query_groups={G}")
            if past < kv_cache.max_seq_len and k_all.dtype == torch.bfloat16:
                left, _ = self._windows(left_window, right_window, causal)
                cos, sin = self.rope.get_cos_sin_cache(kv_cache.max_seq_len)
                seq_lens = torch.full((B,), past + 1, dtype=torch.int32, device=x.device)
                qd, kd, vd = q[:, 0], k[:, 0], v[:, 0]
                if qd.dtype not in (torch.float32, torch.bfloat16):
                    qd, kd, vd = qd.float(), kd.float(), vd.float()
                q_rot = ops.decode_prepare(qd, kd, vd, k_all, v_all, seq_lens, cos, sin, bool(use_qk_norm), 1e-6)
                kv_cache.advance(layer_idx, 1)
                o = ops.gqa_swa_decode(q_rot, k_all, v_all, seq_lens, float(self.softmax_scale), left)
                cache_out = {"k": k_all[:, past:past + 1].to(x.dtype), "v": v_all[:, past:past + 1].to(x.dtype)}
                return self.w_o(o.to(x.dtype).reshape(B, 1, self.d_model)), cache_out

        past = kv_cache.layer_seq_len(layer_idx) if (cached and kv_cache.cache is not None) else 0
        fused = x.is_cuda and hd % 2 == 0 and q.dtype in (torch.float32, torch.bfloat16)
        if fused:
            # qk-norm + RoPE (positions past .. past+T-1) + bf16 rounding + TMA-addressable layout in ONE launch
            # (reference :467-474 does this with ~12 elementwise passes); the results feed the op directly
            cos, sin = self.rope.get_cos_sin_cache(past + T)
            q, k, v = ops.prefill_prepare_views(q, k, v, cos, sin, past, bool(use_qk_norm), 1e-6)
        else:
            if use_qk_norm:
                q, k = apply_qk_norm(q, k)
            q = self.rope(q, offset=past)
            k = self.rope(k, offset=past)

        if padding_mask is not None:
            if padding_mask.shape != (B, T):
                raise ValueError(f"Expected padding mask of shape ({B, T}), got {padding_mask.shape}")
            padding_mask = padding_mask.bool()

        left, right = self._windows(left_window, right_window, causal)
        cache_out = {"k": k.to(x.dtype), "v": v.to(x.dtype)} if use_cache else None

        def core(q_, k_, v_):
            if fused:   # already bf16 in the kernels' layout: no second cast / pad pass
                return ops.gqa_swa_prefill(q_, k_, v_, padding_mask, None, float(self.softmax_scale), bool(causal),
                                           int(left), int(right)).to(x.dtype)
            return attention_core(q_, k_, v_, scale=self.softmax_scale, causal=causal, left=left, right=right,
                                  q_valid=padding_mask, out_dtype=x.dtype)

        if cached:
            # intended contract of reference :508-516 — append at the layer's length, attend the cache
            kv_cache.update(layer_idx, k.to(torch.bfloat16), v.to(torch.bfloat16))
            total = kv_cache.layer_seq_len(layer_idx)
            k_all = kv_cache.cache[layer_idx]["k"]
            v_all = kv_cache.cache[layer_idx]["v"]
            if k_all.size(2) != G:  # cache built with the expanded head count: use every (H/G)-th head's slot
                raise ValueError(f"KVCache stores {k_all.size(2)} heads; the drop-in Attention needs query_groups={G}")
            if T == 1 and padding_mask is None and causal:
                seq_lens = torch.full((B,), total, dtype=torch.int32, device=x.device)
                o = ops.gqa_swa_decode(q[:, 0].to(torch.bfloat16), k_all, v_all, seq_lens, float(self.softmax_scale),
                                       left).to(x.dtype)[:, None]
            else:
                o = core(q, k_all[:, :total], v_all[:, :total])
        else:
            # reference SDPA-path semantics: padding masks QUERY rows (src/optimized_attention.py:673-675)
            o = core(q, k, v)
        return self.w_o(o.reshape(B, T, self.d_model)), cache_out


class AttentionBlock(nn.Module):
    """x + dropout(attn(rms_norm(x))) (reference src/optimized_attention.py:732-817)."""

    def __init__(self, d_model: int, num_heads: int, query_groups: int, softmax_scale: float, use_proj_bias: bool,
                 use_qkv_proj: bool, dropout: float, theta: float, eps: float):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.rms_norm = RMSNorm(d_model, eps)
        self.attn = Attention(d_model=d_model, num_heads=num_heads, query_groups=query_groups, theta=theta,
                              softmax_scale=softmax_scale, use_proj_bias=use_proj_bias, use_qkv_proj=use_qkv_proj)

    def forward(self, x: torch.Tensor, left_window: int, right_window: int, causal: bool = True,
                padding_mask: Optional[torch.Tensor] = None, kv_cache: Optional[KVCache] = None,
                layer_idx: Optional[int] = None, use_cache: bool = False, use_mqa: bool = False,
                use_qk_norm: bool = True) -> Tuple[torch.Tensor, Optional[Dict[str, torch.Tensor]]]:
        attn_out, cache_out = self.attn(self.rms_norm(x), left_window=left_window, right_window=right_window,
                                        causal=causal, padding_mask=padding_mask, kv_cache=kv_cache,
                                        layer_idx=layer_idx, use_cache=use_cache, use_mqa=use_mqa,
                                        use_qk_norm=use_qk_norm)
        return x + self.dropout(attn_out), cache_out
