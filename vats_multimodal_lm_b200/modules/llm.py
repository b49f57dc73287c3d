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
    (the reference's branch at :508-516 is never entered).  A cache built with `num_heads = H` as the unmodified
    call sites do (src/transformers/nlp/model.py:148-154, inference/generate.py:27-33) is bound to the G un-expanded
    KV heads by its first writer; generate.py's call sequence (prefill with use_cache + mask, uncached re-forward,
    T=1 steps with a [B,1] mask, :96-127) reaches the decode kernel;
  * RoPE positions continue from the cache length during cached decoding (the reference restarts at 0, :59);
  * `cache_out` holds the new [B, T, G, hd] k/v (the reference slices an already transposed tensor, :726).
"""
from __future__ import annotations

import weakref
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ._common import QK_NORM_LOGIT_BOUND, apply_qk_norm, attention_core, get_default_window_mode, setup_projections, WINDOW_MODES


def _on_gpu(x: torch.Tensor) -> bool:
    """The fused producers and the decode kernels exist on the GPU only (CPU tests of the host logic patch this
    together with stand-ins for the ops)."""
    return x.is_cuda


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
    The cos / sin tables are always kept in fp32 (the fused producers read them as fp32), whatever dtype the module
    is cast to or the checkpoint carries.
    """

    def __init__(self, head_dim: int, theta: float):
        super().__init__()
        if head_dim % 2 != 0:
            raise ValueError(f"head_dim ({head_dim}) must be divisible by 2 for even splitting.")
        self.head_dim = head_dim
        self.theta = float(theta)
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
                setattr(self, name, t.to(device=self.inv_freq.device, dtype=torch.float32))
                self.cached_seq_len = t.size(0)
        state_dict.setdefault(prefix + "cos_cache", self.cos_cache)
        state_dict.setdefault(prefix + "sin_cache", self.sin_cache)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def _inv_freq_fp32(self) -> torch.Tensor:
        # `module.to(torch.bfloat16)` rounds the buffer; the tables are rebuilt from the exact fp32 frequencies
        if self.inv_freq.dtype == torch.float32:
            return self.inv_freq
        hd = self.head_dim
        return 1.0 / (self.theta ** (torch.arange(0, hd, 2, dtype=torch.float32, device=self.inv_freq.device) / hd))

    def _update_cache(self, seq_len: int) -> None:
        inv = self._inv_freq_fp32()
        pos = torch.arange(seq_len, device=inv.device, dtype=torch.float32)
        freqs = torch.outer(pos, inv)
        self.cos_cache = torch.cos(freqs)
        self.sin_cache = torch.sin(freqs)
        self.cached_seq_len = seq_len

    def get_cos_sin_cache(self, seq_len: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if seq_len > self.cached_seq_len or self.cos_cache.device != self.inv_freq.device or \
                self.cos_cache.dtype != torch.float32 or self.cos_cache.size(0) < seq_len:
            self._update_cache(max(seq_len, self.cached_seq_len))
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
    `get`, `reset`, `.cache[layer]['k'|'v']` of shape [B, max_seq_len, heads, head_dim], `.current_seq_len`.

    It is built exactly as the unmodified call sites build it — `KVCache(max_batch_size, max_seq_len,
    num_heads=model_args.num_heads, head_dim, num_layers)` at src/transformers/nlp/model.py:148-154 and
    src/transformers/nlp/inference/generate.py:27-33 — i.e. with the FULL head count.  What it stores is decided by
    its first writer: the drop-in `Attention` binds it to its `query_groups` un-expanded KV heads (`bind`), so the
    H/G-fold expanded copy of the reference is never allocated; `kv_heads` reports the stored head count.

    Differences from the reference, all needed for a cache that actually works (SURVEY quirk ledger):
      * the length is tracked per layer (the reference advances ONE counter on every layer's `update`, :257);
        `current_seq_len` reports the longest layer, which is what a caller polling it after a forward expects;
      * storage is allocated lazily, on first use, on the device of the tensors written into it, in bf16, with the
        head stride rounded up to 8 elements so every row is 16-byte aligned (TMA-addressable: head_dim 60 -> 64);
        `.cache[l]['k']` is the `[..., :head_dim]` view of that buffer;
      * `reset()` also resets the live caches of identical geometry ("siblings", `link_sibling_resets`): the
        reference's generator resets and initialises ITS OWN KVCache (generate.py:93-94, 240) while the model only
        ever passes `self.kv_cache` (model.py:302) to the layers — with a cache that fills, the second `_generate`
        call would otherwise continue the first one's sequence.
    """

    link_sibling_resets = True
    _siblings: Dict[tuple, "weakref.WeakSet"] = {}

    def __init__(self, max_batch_size: int, max_seq_len: int, num_heads: int, head_dim: int, num_layers: int,
                 dtype: torch.dtype = torch.bfloat16, device: Optional[torch.device] = None):
        self.max_batch_size = max_batch_size
        self.max_seq_len = max_seq_len
        self.num_heads = num_heads
        self.head_dim = head_dim
        self.num_layers = num_layers
        self.dtype = dtype
        self.device = device
        self.kv_heads = num_heads       # heads actually stored (bound to query_groups by the drop-in Attention)
        self.batch_size = None
        self._layer_len = None
        self._store = None              # [{'k': padded buffer, 'v': ...}] per layer
        self._views = None              # [{'k': buffer[..., :head_dim], 'v': ...}] per layer
        self._alloc_device = device
        self._geometry = (max_batch_size, max_seq_len, num_heads, head_dim, num_layers)
        KVCache._siblings.setdefault(self._geometry, weakref.WeakSet()).add(self)

    # ---- reference attributes
    @property
    def current_seq_len(self) -> Optional[int]:
        return None if self._layer_len is None else max(self._layer_len)

    @current_seq_len.setter
    def current_seq_len(self, value: Optional[int]) -> None:
        if value is None:
            self._layer_len = None
        elif self._layer_len is not None:
            self._layer_len = [int(value)] * self.num_layers

    @property
    def cache(self):
        if self.batch_size is None:
            return None
        self._materialize()
        return self._views

    @cache.setter
    def cache(self, value) -> None:
        if value is not None:
            raise AttributeError("KVCache.cache can only be cleared (set to None); storage is owned by the cache")
        self._store = self._views = None

    @property
    def head_stride(self) -> int:
        """Elements between consecutive heads of one cached token (head_dim rounded up to 8)."""
        return (self.head_dim + 7) // 8 * 8

    def layer_seq_len(self, layer_idx: int) -> int:
        return 0 if self._layer_len is None else self._layer_len[layer_idx]

    def _materialize(self) -> None:
        if self._store is not None:
            return
        device = self._alloc_device
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu"
        shape = (self.batch_size, self.max_seq_len, self.kv_heads, self.head_stride)
        self._store = [{"k": torch.zeros(shape, dtype=self.dtype, device=device),
                        "v": torch.zeros(shape, dtype=self.dtype, device=device)} for _ in range(self.num_layers)]
        self._views = [{n: t[..., :self.head_dim] for n, t in layer.items()} for layer in self._store]

    def initialize(self, batch_size: int, device: Optional[torch.device] = None) -> None:
        if batch_size > self.max_batch_size:
            raise ValueError(f"batch_size ({batch_size}) exceeds max_batch_size ({self.max_batch_size})")
        self.batch_size = batch_size
        self._layer_len = [0] * self.num_layers
        self._alloc_device = device or self.device
        self._store = self._views = None          # zero-filled again on first use

    def bind(self, kv_heads: int, device: Optional[torch.device] = None) -> None:
        """Fix the number of heads stored (and, if still open, the device).  Called by the first writer; changing the
        head count of a cache that already holds tokens is an error."""
        if device is not None and self._alloc_device is None:
            self._alloc_device = device
        if kv_heads == self.kv_heads:
            return
        if self._layer_len is not None and any(self._layer_len):
            raise ValueError(f"KVCache already holds tokens with {self.kv_heads} heads; cannot re-bind to {kv_heads}")
        self.kv_heads = kv_heads
        self._store = self._views = None

    def update(self, layer_idx: int, k: torch.Tensor, v: torch.Tensor) -> None:
        """Append k, v [B, T, heads, head_dim] at the layer's current length (truncating at max_seq_len like the
        reference, :241-250; the drop-in Attention raises before it would come to that)."""
        if self.batch_size is None or self.batch_size != k.size(0):
            self.initialize(k.size(0), device=k.device)
        self.bind(k.size(2), k.device)
        cur = self._layer_len[layer_idx]
        new = k.size(1)
        space = self.max_seq_len - cur
        if space <= 0:
            return
        if new > space:
            k, v, new = k[:, :space], v[:, :space], space
        layer = self.cache[layer_idx]
        layer["k"][:, cur:cur + new] = k.to(layer["k"].dtype)
        layer["v"][:, cur:cur + new] = v.to(layer["v"].dtype)
        self._layer_len[layer_idx] = cur + new

    def advance(self, layer_idx: int, new: int) -> None:
        """Record `new` tokens written into the layer's cache in place (the fused decode step appends on the device)."""
        self._layer_len[layer_idx] = min(self._layer_len[layer_idx] + new, self.max_seq_len)

    def get(self, layer_idx: int, seq_len: int) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        if self.batch_size is None or seq_len > self._layer_len[layer_idx]:
            return None, None
        layer = self.cache[layer_idx]
        return layer["k"][:, :seq_len], layer["v"][:, :seq_len]

    def _reset_self(self) -> None:
        self._store = self._views = None
        self.batch_size = None
        self._layer_len = None
        self.kv_heads = self.num_heads
        self._alloc_device = self.device

    def reset(self) -> None:
        self._reset_self()
        if KVCache.link_sibling_resets:
            for other in list(KVCache._siblings.get(self._geometry, ())):
                if other is not self:
                    other._reset_self()


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
        if padding_mask is not None:
            if padding_mask.shape != (B, T):
                raise ValueError(f"Expected padding mask of shape ({B, T}), got {padding_mask.shape}")
            padding_mask = padding_mask.bool()

        H, G, hd = self.num_heads, self.query_groups, self.head_dim
        if self.use_qkv_proj:
            qkv = self.w_qkv(x)
            q, k, v = torch.split(qkv, [H * hd, G * hd, G * hd], dim=-1)
        else:
            q, k, v = self.w_q(x), self.w_k(x), self.w_v(x)
        q = q.view(B, T, H, hd)
        k = k.view(B, T, G, hd)
        v = v.view(B, T, G, hd)
        left, right = self._windows(left_window, right_window, causal)

        cached = bool(use_cache) and kv_cache is not None and layer_idx is not None
        past = 0
        if cached:
            # the cache arrives as the unmodified call sites build it (num_heads = H, model.py:148-154): bind it to
            # the G heads this layer writes; it allocates on first use, on x's device
            if kv_cache.batch_size is None or kv_cache.batch_size != B:
                kv_cache.initialize(B, device=x.device)
            kv_cache.bind(G, x.device)
            past = kv_cache.layer_seq_len(layer_idx)
            if past + T > kv_cache.max_seq_len:
                raise ValueError(f"KV cache overflow: {past} cached + {T} new tokens > max_seq_len "
                                 f"({kv_cache.max_seq_len}); truncating would shift the causal alignment")

        # ---- single-token cached decode (what generate.py:115-127 issues for every step after the first, always
        #      with a [B, 1] mask of the unfinished sequences): qk-norm + RoPE + bf16 rounding + cache append in ONE
        #      kernel, then the split-K decode kernel — two launches per layer.  A finished sequence (mask False) is
        #      a masked QUERY row (reference :673-675): its k/v are still appended, its output row is zero
        #      (seq_lens 0).  `causal` does not matter for a single bottom-right aligned query.
        if cached and T == 1 and _on_gpu(x) and hd % 2 == 0 and kv_cache.dtype == torch.bfloat16:
            layer = kv_cache.cache[layer_idx]
            k_all, v_all = layer["k"], layer["v"]
            cos, sin = self.rope.get_cos_sin_cache(kv_cache.max_seq_len)
            seq_lens = torch.full((B,), past + 1, dtype=torch.int32, device=x.device)
            qd, kd, vd = q[:, 0], k[:, 0], v[:, 0]
            if qd.dtype not in (torch.float32, torch.bfloat16):
                qd, kd, vd = qd.float(), kd.float(), vd.float()
            q_rot = ops.decode_prepare(qd, kd, vd, k_all, v_all, seq_lens, cos, sin, bool(use_qk_norm), 1e-6)
            kv_cache.advance(layer_idx, 1)
            dec_lens = seq_lens if padding_mask is None else seq_lens * padding_mask[:, 0].to(torch.int32)
            o = ops.gqa_swa_decode(q_rot, k_all, v_all, dec_lens, float(self.softmax_scale), left)
            cache_out = {"k": k_all[:, past:past + 1].to(x.dtype), "v": v_all[:, past:past + 1].to(x.dtype)}
            return self.w_o(o.to(x.dtype).reshape(B, 1, self.d_model)), cache_out

        # (training: the fused producers have no backward — qk-norm / RoPE then run as differentiable PyTorch ops and the
        #  core's own backward kernels take over, `vats::gqa_swa_prefill_bwd`)
        fused = _on_gpu(x) and hd % 2 == 0 and q.dtype in (torch.float32, torch.bfloat16) and not (
            torch.is_grad_enabled() and q.requires_grad)
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

        cache_out = {"k": k.to(x.dtype), "v": v.to(x.dtype)} if use_cache else None

        bound = QK_NORM_LOGIT_BOUND if use_qk_norm else 0.0   # unit-norm q, k (also the cached keys): |q.k| <= 1

        def core(q_, k_, v_):
            if fused:   # already bf16 in the kernels' layout: no second cast / pad pass
                return ops.gqa_swa_prefill(q_, k_, v_, padding_mask, None, float(self.softmax_scale), bool(causal),
                                           int(left), int(right), 0, bound).to(x.dtype)
            return attention_core(q_, k_, v_, scale=self.softmax_scale, causal=causal, left=left, right=right,
                                  q_valid=padding_mask, out_dtype=x.dtype, logit_bound=bound)

        if cached:
            # intended contract of reference :508-516 — append at the layer's length, attend the cache
            # (bottom-right aligned: query t of the chunk sits at position past + t)
            kv_cache.update(layer_idx, k, v)
            total = kv_cache.layer_seq_len(layer_idx)
            layer = kv_cache.cache[layer_idx]
            if total == T and layer["k"].dtype == k.dtype:
                o = core(q, k, v)                    # first chunk: the fresh k, v are the whole context
            else:
                o = core(q, layer["k"][:, :total], layer["v"][:, :total])
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
