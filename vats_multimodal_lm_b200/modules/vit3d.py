"""Drop-in replacements for reference src/transformers/vision/vit_3d/optimized_attention.py
(SpatioTemporalAttention :21-670, SpatioTemporalAttentionBlock :673-741) and rope_3d.py (RoPE3D :9-237).

Factorized attention: a spatial pass over B*T sequences of H*W patches, then a temporal pass over B*H*W sequences
of T frames, sharing one `w_qkv` and followed by one `w_o`.  Both passes run in `torch.ops.vats.gqa_swa_prefill`
(non-causal, default scale 1/sqrt(head_dim), key-padding mask as `k_valid`) — the tensor-core kernel for the
196-token spatial pass, the CUDA-core kernel for the 8-token temporal pass.

Reference layout quirks are reproduced verbatim so module outputs stay comparable: the temporal padding mask is the
[B, T*H*W] mask re-viewed as [-1, T] without a transpose (:271) and the temporal output is re-viewed as
[B, T, H*W, d] without the inverse transpose (:666-668).
"""
from __future__ import annotations

import math
from typing import Literal, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ._common import QK_NORM_LOGIT_BOUND, apply_qk_norm, attention_core, get_default_window_mode, setup_projections, WINDOW_MODES
from .llm import RMSNorm, _on_gpu


class RoPE3D(nn.Module):
    """head_dim is cut into three axis blocks (t, h, w) of head_dim/3; interleaved pairs of a block rotate with the
    patch coordinate on that axis.  Spatial mode rotates the h and w blocks, temporal mode the t block
    (reference rope_3d.py:185-219)."""

    def __init__(self, head_dim: int, theta: float, patch_size: Tuple[int, int, int]):
        super().__init__()
        if head_dim % 6 != 0:
            raise ValueError(
                f"head_dim must be divisible by 6 for 3D RoPE (2 dims per spatial dimension), got {head_dim}")
        self.head_dim = head_dim
        self.theta = theta
        self.patch_size = patch_size
        self.dim_per_axis = head_dim // 3
        if self.dim_per_axis % 2 != 0:
            raise ValueError(f"head_dim // 3 must be even for proper rotation pairs, got head_dim={head_dim}, "
                             f"dim_per_axis={self.dim_per_axis}")
        pairs = self.dim_per_axis // 2
        freqs = 1.0 / (theta ** (torch.arange(0, pairs, dtype=torch.float32) * 2.0 / self.dim_per_axis))
        self.register_buffer("freqs_t", freqs.clone())
        self.register_buffer("freqs_h", freqs.clone())
        self.register_buffer("freqs_w", freqs.clone())

    @staticmethod
    def _rotate_block(x: torch.Tensor, positions: torch.Tensor, freqs: torch.Tensor, start: int) -> torch.Tensor:
        pairs = freqs.numel()
        end = start + 2 * pairs
        ang = positions.to(freqs.dtype)[:, None] * freqs[None]          # [N, pairs]
        cos = torch.cos(ang)[None, :, None, :].to(x.dtype)
        sin = torch.sin(ang)[None, :, None, :].to(x.dtype)
        blk = x[..., start:end].reshape(*x.shape[:-1], pairs, 2)
        a, b = blk[..., 0], blk[..., 1]
        rot = torch.stack([a * cos - b * sin, a * sin + b * cos], dim=-1).reshape(*x.shape[:-1], 2 * pairs)
        return torch.cat([x[..., :start], rot, x[..., end:]], dim=-1)

    def tables(self, grid_shape: Tuple[int, int, int], attn_mode: str
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """The rotation of `forward` as (cos [L, hd], signed sin [L, hd], partner [hd]) for the fused producer: the
        rotated blocks hold interleaved pairs (2i, 2i+1); every other column has cos = 1, sin = 0."""
        key = (tuple(grid_shape), attn_mode, self.freqs_t.device)
        cached = getattr(self, "_tables", {})
        if key in cached:
            return cached[key]
        gt, gh, gw = grid_shape
        dev, hd, dpa = self.freqs_t.device, self.head_dim, self.dim_per_axis
        if attn_mode == "spatial":
            hh, ww = torch.meshgrid(torch.arange(gh, device=dev), torch.arange(gw, device=dev), indexing="ij")
            blocks = [(hh.flatten(), self.freqs_h, dpa), (ww.flatten(), self.freqs_w, 2 * dpa)]
            L = gh * gw
        elif attn_mode == "temporal":
            blocks = [(torch.arange(gt, device=dev), self.freqs_t, 0)]
            L = gt
        else:
            raise ValueError(f"attn_mode must be 'spatial' or 'temporal' got {attn_mode}")
        cos = torch.ones(L, hd, dtype=torch.float32, device=dev)
        sin = torch.zeros(L, hd, dtype=torch.float32, device=dev)
        partner = torch.arange(hd, device=dev)
        for pos, freqs, start in blocks:
            ang = pos.to(torch.float32)[:, None] * freqs.float()[None]          # [L, pairs]
            pairs = freqs.numel()
            ev = start + 2 * torch.arange(pairs, device=dev)
            cos[:, ev] = torch.cos(ang)
            cos[:, ev + 1] = torch.cos(ang)
            sin[:, ev] = -torch.sin(ang)
            sin[:, ev + 1] = torch.sin(ang)
            partner[ev] = ev + 1
            partner[ev + 1] = ev
        out = (cos.contiguous(), sin.contiguous(), partner.to(torch.int32))
        cached[key] = out
        self._tables = cached
        return out

    def forward(self, x: torch.Tensor, grid_shape: Tuple[int, int, int], attn_mode: Literal["spatial", "temporal"]
                ) -> torch.Tensor:
        assert x.dim() == 4, f"x must be a 4 dimensional tensor, got {x.dim()}"
        gt, gh, gw = grid_shape
        dev = x.device
        if attn_mode == "spatial":
            hh, ww = torch.meshgrid(torch.arange(gh, device=dev), torch.arange(gw, device=dev), indexing="ij")
            x = self._rotate_block(x, hh.flatten(), self.freqs_h, self.dim_per_axis)
            return self._rotate_block(x, ww.flatten(), self.freqs_w, 2 * self.dim_per_axis)
        if attn_mode == "temporal":
            return self._rotate_block(x, torch.arange(gt, device=dev), self.freqs_t, 0)
        raise ValueError(f"attn_mode must be 'spatial' or 'temporal' got {attn_mode}")


class SpatioTemporalAttention(nn.Module):
    """Factorized (1 x H x W, then T x 1 x 1) GQA (reference vit_3d/optimized_attention.py:21-670)."""

    def __init__(self, d_model: int, num_heads: int, query_groups: int, rope_theta: float,
                 patch_size: Tuple[int, int, int], *, window_mode: Optional[str] = None):
        super().__init__()
        if d_model % num_heads != 0:
            raise ValueError(f"Expected d_model to be divisble by num_heads, got {d_model} % {num_heads} != 0")
        if num_heads % query_groups != 0:
            raise ValueError(
                f"Expected num_heads to be divisble by query_groups, got {num_heads} % {query_groups} != 0")
        if window_mode is not None and window_mode not in WINDOW_MODES:
            raise ValueError(f"window_mode must be one of {WINDOW_MODES}")
        self.d_model = d_model
        self.num_heads = num_heads
        self.query_groups = query_groups
        self.head_dim = d_model // num_heads
        self.heads_per_group = num_heads // query_groups
        self.window_mode = window_mode
        self.w_qkv, self.w_o = setup_projections(d_model, num_heads, self.head_dim, True, True, False, query_groups)
        self.rope = RoPE3D(self.head_dim, rope_theta, patch_size)

    def _setup_qkv(self, x: torch.Tensor, use_mqa: bool, use_qk_norm: bool, grid_shape: Tuple[int, int, int],
                   attn_mode: str) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Projections, qk-norm and RoPE of one pass.  spatial: q/k/v [B*T, S, heads, hd].  temporal: [B, S, T, heads, hd]
        as PERMUTED VIEWS of tensors laid out [B, T, S, heads, hd] — the reference materialises
        `x.transpose(1, 2).contiguous()` first (vit_3d/optimized_attention.py:474-479, a full extra read + write of
        the activations); the projection is row-wise and qk-norm / RoPE are element-wise along T, so nothing here needs
        the transposed copy, and the one pass that casts q/k/v to the kernels' bf16 layout (`_pass`) does the
        transposition on the way."""
        B, T, S, _ = x.shape
        H, G, hd = self.num_heads, self.query_groups, self.head_dim
        if attn_mode == "spatial":
            x = x.reshape(B * T, S, self.d_model)
            q, k, v = torch.split(self.w_qkv(x), [H * hd, G * hd, G * hd], dim=-1)
            n, L = x.shape[:2]
            q = q.view(n, L, H, hd)
            k = k.view(n, L, G, hd)
            v = v.view(n, L, G, hd)
            if use_qk_norm:
                q, k = apply_qk_norm(q, k)
            return self.rope(q, grid_shape, attn_mode), self.rope(k, grid_shape, attn_mode), v
        if attn_mode != "temporal":
            raise ValueError(f"attn_mode must be 'spatial' or 'temporal', got {attn_mode}")
        q, k, v = torch.split(self.w_qkv(x), [H * hd, G * hd, G * hd], dim=-1)   # [B, T, S, .]: no transposed copy of x
        # sequence axis = T (dim 1); (S, heads) together play the role of the head axis for the element-wise producers
        q = q.reshape(B, T, S * H, hd)
        k = k.reshape(B, T, S * G, hd)
        if use_qk_norm:
            q, k = apply_qk_norm(q, k)
        q = self.rope(q, grid_shape, attn_mode)
        k = self.rope(k, grid_shape, attn_mode)
        return (q.view(B, T, S, H, hd).permute(0, 2, 1, 3, 4), k.view(B, T, S, G, hd).permute(0, 2, 1, 3, 4),
                v.reshape(B, T, S, G, hd).permute(0, 2, 1, 3, 4))

    def _pass(self, x: torch.Tensor, use_mqa: bool, use_qk_norm: bool, grid_shape, window, padding_mask,
              attn_mode: str) -> torch.Tensor:
        B, T, S, _ = x.shape
        k_valid = None
        if padding_mask is not None:
            # reference :264-277 — plain re-views of the [B, T*H*W] mask, key-padding semantics
            k_valid = padding_mask.reshape(B * T, S) if attn_mode == "spatial" else padding_mask.reshape(-1, T)
            k_valid = k_valid.bool()
        left, right = window
        scale = 1.0 / math.sqrt(self.head_dim)
        H, G, hd = self.num_heads, self.query_groups, self.head_dim
        bound = QK_NORM_LOGIT_BOUND if use_qk_norm else 0.0
        if _on_gpu(x) and x.dtype in (torch.float32, torch.bfloat16) and not (
                torch.is_grad_enabled() and (x.requires_grad or self.w_o.weight.requires_grad)):
            # inference: qk-norm + 3-D RoPE + bf16 rounding + kernel layout in ONE launch, reading the projection in place
            # — for the temporal pass through (b, s)-permuted views, so neither the transposed copy of x nor a cast pass
            # exists any more
            q, k, v = torch.split(self.w_qkv(x), [H * hd, G * hd, G * hd], dim=-1)       # [B, T, S, .]
            q5, k5, v5 = q.view(B, T, S, H, hd), k.view(B, T, S, G, hd), v.view(B, T, S, G, hd)
            if attn_mode == "spatial":      # sequences (b, t), tokens s
                pass
            else:                           # sequences (b, s), tokens t
                q5, k5, v5 = (t.permute(0, 2, 1, 3, 4) for t in (q5, k5, v5))
            cos, sin, partner = self.rope.tables(grid_shape, attn_mode)
            qk, kk, vk = ops.prefill_prepare_table_views(q5, k5, v5, cos, sin, partner, bool(use_qk_norm))
            o = ops.gqa_swa_prefill(qk, kk, vk, None, k_valid, scale, False, int(left), int(right), 0, bound).to(x.dtype)
            return o.reshape(qk.size(0), qk.size(1), self.d_model)
        q, k, v = self._setup_qkv(x, use_mqa, use_qk_norm, grid_shape, attn_mode)
        if attn_mode == "temporal":
            # the bf16 cast writes the [B*S, T, heads, hd] layout the kernel reads: transposition fused into the cast
            def cast_transposed(t5):
                buf = torch.empty(t5.shape, dtype=torch.bfloat16, device=t5.device)   # contiguous [B, S, T, heads, hd]
                buf.copy_(t5)
                return buf.view(B * S, T, t5.size(3), self.head_dim)
            o = ops.gqa_swa_prefill(cast_transposed(q), cast_transposed(k), cast_transposed(v), None, k_valid, scale,
                                    False, int(left), int(right), 0, bound).to(x.dtype)
            return o.reshape(B * S, T, self.d_model)
        o = attention_core(q, k, v, scale=scale, causal=False, left=left, right=right, k_valid=k_valid,
                           out_dtype=x.dtype, logit_bound=bound)
        return o.reshape(q.size(0), q.size(1), self.d_model)

    def forward(self, x: torch.Tensor, grid_size: Tuple[int, int, int], use_mqa: bool, use_qk_norm: bool,
                window_size: Optional[Tuple[int, int]] = None, padding_mask: Optional[torch.Tensor] = None
                ) -> torch.Tensor:
        assert x.dim() == 4, f"x must have 4 dimensions, got {x.dim()} dimensions."
        mode = self.window_mode or get_default_window_mode()
        window = (-1, -1)
        if window_size is not None and mode != "reference_sdpa":
            window = (int(window_size[0]), int(window_size[1]))
        B = x.size(0)
        if x.numel() == 0:   # no frames / no patches: nothing to attend (the reference's `.view(..., -1, ...)` is ambiguous here)
            return x.new_zeros(B, x.size(1), x.size(2), self.d_model)
        spatial = self._pass(x, use_mqa, use_qk_norm, grid_size, window, padding_mask, "spatial")
        spatial = spatial.view(B, grid_size[0], -1, self.d_model)
        temporal = self._pass(spatial, use_mqa, use_qk_norm, grid_size, window, padding_mask, "temporal")
        return self.w_o(temporal.reshape(B, grid_size[0], -1, self.d_model))


class SpatioTemporalAttentionBlock(nn.Module):
    """x + dropout(attention(rms_norm(x))) (reference vit_3d/optimized_attention.py:673-741)."""

    def __init__(self, d_model: int, num_heads: int, query_groups: int, rope_theta: float,
                 patch_size: Tuple[int, int, int], eps: float, dropout: float):
        super().__init__()
        self.attention = SpatioTemporalAttention(d_model=d_model, num_heads=num_heads, query_groups=query_groups,
                                                 rope_theta=rope_theta, patch_size=patch_size)
        self.rms_norm = RMSNorm(d_model=d_model, eps=eps)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x: torch.Tensor, grid_size: Tuple[int, int, int], use_mqa: bool, use_qk_norm: bool,
                window_size: Tuple[int, int], padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        return x + self.dropout(self.attention(self.rms_norm(x), grid_size=grid_size, use_mqa=use_mqa,
                                               use_qk_norm=use_qk_norm, window_size=window_size,
                                               padding_mask=padding_mask))
