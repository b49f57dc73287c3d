"""Drop-in replacements for reference src/transformers/vision/vit_2d/optimized_attention.py:
RoPE (2-D axial, :20-197), SpatialAttention (:199-614), SpatialAttentionBlock (:617-697).

Same constructor / forward signatures and parameter names (`qkv_proj`, `o_proj`, `q_proj/k_proj/v_proj`,
`rope.inv_freq`).  The non-causal GQA core over the flattened H x W patches runs in
`torch.ops.vats.gqa_swa_prefill`; like the reference's executable path (`_torch_attention`, :396-402) it uses the
default scale 1/sqrt(head_dim) — `self.softmax_scale` is only consumed by the reference's dead FA2 branch (:336).
Windows: honoured when `use_windowed_attn` (the intent of :331-338) unless window_mode == "reference_sdpa".
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ._common import QK_NORM_LOGIT_BOUND, apply_qk_norm, attention_core, get_default_window_mode, setup_projections, WINDOW_MODES
from .llm import RMSNorm, _on_gpu


class RoPE(nn.Module):
    """2-D rotary embedding: head_dim is cut into 4 blocks (x1, x2, y1, y2) of head_dim/4; (x1,x2) rotate with the
    patch row index, (y1,y2) with the patch column index (reference vit_2d/optimized_attention.py:128-172)."""

    def __init__(self, head_dim: int, target_size: int, patch_size: int, rope_theta: float):
        super().__init__()
        if head_dim % 4 != 0:
            raise ValueError(f"head_dim must be divisible by 4 for 2D RoPE, head_dim: {head_dim}")
        self.head_dim = head_dim
        self.patch_size = patch_size
        self.grid_size = target_size // patch_size
        self.num_patches = self.grid_size ** 2
        freq_dim = head_dim // 4
        inv_freq = 1.0 / (rope_theta ** (torch.arange(0, freq_dim, dtype=torch.float32) / freq_dim))
        self.register_buffer("inv_freq", inv_freq)

    def _angles(self, grid_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
        pos = torch.arange(grid_size, dtype=self.inv_freq.dtype, device=self.inv_freq.device)
        gx, gy = torch.meshgrid(pos, pos, indexing="ij")
        theta_x = gx.flatten()[:, None] * self.inv_freq  # [T, hd/4]
        theta_y = gy.flatten()[:, None] * self.inv_freq
        return theta_x, theta_y

    def tables(self, T: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """The rotation of `forward` as (cos [T, hd], signed sin [T, hd], partner [hd]) for the fused producer
        (`vats::prefill_prepare_table`): out[c] = x[c] * cos[t][c] + x[partner[c]] * sin[t][c]."""
        key = (T, self.inv_freq.device)
        cached = getattr(self, "_tables", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        grid = int(math.sqrt(T))
        assert grid * grid == self.num_patches, (
            f"pos_x_flat must have shape of {(self.num_patches, 1)}, got {(grid * grid, 1)}")
        tx, ty = self._angles(grid)
        tx, ty = tx.float(), ty.float()
        fd = self.head_dim // 4
        cos = torch.cat([torch.cos(tx), torch.cos(tx), torch.cos(ty), torch.cos(ty)], dim=1).contiguous()
        sin = torch.cat([-torch.sin(tx), torch.sin(tx), -torch.sin(ty), torch.sin(ty)], dim=1).contiguous()
        idx = torch.arange(fd, device=cos.device)
        partner = torch.cat([idx + fd, idx, idx + 3 * fd, idx + 2 * fd]).to(torch.int32)
        self._tables = (key, (cos, sin, partner))
        return cos, sin, partner

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.dim() == 4, f"x must have 4 dimensions, got {x.dim()} dimensions."
        T = x.size(1)
        grid = int(math.sqrt(T))
        # the reference asserts the flattened grid against the constructor's num_patches (:95-100)
        assert grid * grid == self.num_patches, (
            f"pos_x_flat must have shape of {(self.num_patches, 1)}, got {(grid * grid, 1)}")
        tx, ty = self._angles(grid)
        cx, sx = torch.cos(tx)[None, :, None, :].to(x.dtype), torch.sin(tx)[None, :, None, :].to(x.dtype)
        cy, sy = torch.cos(ty)[None, :, None, :].to(x.dtype), torch.sin(ty)[None, :, None, :].to(x.dtype)
        fd = self.head_dim // 4
        x1, x2, y1, y2 = x.reshape(*x.shape[:-1], 4, fd).unbind(dim=-2)
        out = torch.stack((x1 * cx - x2 * sx, x1 * sx + x2 * cx, y1 * cy - y2 * sy, y1 * sy + y2 * cy), dim=-2)
        return out.reshape(*x.shape)


class SpatialAttention(nn.Module):
    """Non-causal GQA over flattened patches (reference vit_2d/optimized_attention.py:199-614)."""

    def __init__(self, d_model: int, num_heads: int, query_groups: int, rope_theta: float, target_size: int,
                 patch_size: int, softmax_scale: float, use_windowed_attn: bool, use_proj_bias: bool,
                 use_fused_proj: bool, *, window_mode: Optional[str] = None):
        super().__init__()
        if d_model % num_heads != 0:
            raise ValueError(f"d_model must be divisble by num_heads, got {d_model} % {num_heads} != 0.")
        if num_heads % query_groups != 0:
            raise ValueError(f"num_heads must be divisble by query_groups, got {num_heads} % {query_groups} != 0.")
        if window_mode is not None and window_mode not in WINDOW_MODES:
            raise ValueError(f"window_mode must be one of {WINDOW_MODES}")
        self.d_model = d_model
        self.num_heads = num_heads
        self.query_groups = query_groups
        self.head_dim = d_model // num_heads
        self.heads_per_group = num_heads // query_groups
        self.softmax_scale = softmax_scale
        self.use_windowed_attn = use_windowed_attn
        self.use_fused_proj = use_fused_proj
        self.window_mode = window_mode
        if use_fused_proj:
            self.qkv_proj, self.o_proj = setup_projections(d_model, num_heads, self.head_dim, True, True,
                                                           use_proj_bias, query_groups)
        else:
            self.q_proj, self.k_proj, self.v_proj, self.o_proj = setup_projections(
                d_model, num_heads, self.head_dim, False, True, use_proj_bias, query_groups)
        self.rope = RoPE(head_dim=self.head_dim, target_size=target_size, patch_size=patch_size,
                         rope_theta=rope_theta)

    def _setup_qkv(self, x: torch.Tensor, use_mqa: bool, use_qk_norm: bool
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        assert x.dim() == 3, f"x must have 3 dims, got {x.dim()}"
        B, T, _ = x.shape
        H, G, hd = self.num_heads, self.query_groups, self.head_dim
        if self.use_fused_proj:
            q, k, v = torch.split(self.qkv_proj(x), [H * hd, G * hd, G * hd], dim=-1)
        else:
            q, k, v = self.q_proj(x), self.k_proj(x), self.v_proj(x)
        q = q.view(B, T, H, hd)
        k = k.view(B, T, G, hd)
        v = v.view(B, T, G, hd)
        if use_qk_norm:
            q, k = apply_qk_norm(q, k)
        return self.rope(q), self.rope(k), v  # K/V stay at G heads: the kernel maps head h -> h // (H/G)

    def forward(self, x: torch.Tensor, use_mqa: bool, use_qk_norm: bool, left_window: int, right_window: int
                ) -> torch.Tensor:
        mode = self.window_mode or get_default_window_mode()
        if not self.use_windowed_attn or mode == "reference_sdpa":
            left_window, right_window = -1, -1  # reference :601-603 (and its SDPA path never windows)
        scale = 1.0 / math.sqrt(self.head_dim)
        bound = QK_NORM_LOGIT_BOUND if use_qk_norm else 0.0
        if _on_gpu(x) and x.dim() == 3 and x.size(1) > 0 and x.dtype in (torch.float32, torch.bfloat16) and not (
                torch.is_grad_enabled() and (x.requires_grad or self.o_proj.weight.requires_grad)):
            # inference: qk-norm + 2-D RoPE + bf16 rounding + kernel layout in ONE launch straight from the projection's
            # views (the PyTorch path below needs ~12 element-wise passes and a cast / pad copy)
            B, T, _ = x.shape
            H, G, hd = self.num_heads, self.query_groups, self.head_dim
            if self.use_fused_proj:
                q, k, v = torch.split(self.qkv_proj(x), [H * hd, G * hd, G * hd], dim=-1)
            else:
                q, k, v = self.q_proj(x), self.k_proj(x), self.v_proj(x)
            cos, sin, partner = self.rope.tables(T)
            q, k, v = ops.prefill_prepare_table_views(q.view(B, 1, T, H, hd), k.view(B, 1, T, G, hd), v.view(B, 1, T, G, hd),
                                                      cos, sin, partner, bool(use_qk_norm))
            o = ops.gqa_swa_prefill(q, k, v, None, None, scale, False, int(left_window), int(right_window), 0,
                                    bound).to(x.dtype)
            return self.o_proj(o.reshape(B, T, self.d_model))
        q, k, v = self._setup_qkv(x, use_mqa=use_mqa, use_qk_norm=use_qk_norm)
        o = attention_core(q, k, v, scale=scale, causal=False, left=left_window, right=right_window, out_dtype=x.dtype,
                           logit_bound=bound)
        return self.o_proj(o.reshape(x.size(0), x.size(1), self.d_model))


class SpatialAttentionBlock(nn.Module):
    """x + dropout(attention(rms_norm(x))) (reference vit_2d/optimized_attention.py:617-697)."""

    def __init__(self, d_model: int, num_heads: int, query_groups: int, rope_theta: float, target_size: int,
                 patch_size: int, softmax_scale: float, use_windowed_attn: bool, use_proj_bias: bool,
                 use_fused_proj: bool, eps: float, dropout: float):
        super().__init__()
        self.attention = SpatialAttention(d_model=d_model, num_heads=num_heads, query_groups=query_groups,
                                          rope_theta=rope_theta, target_size=target_size, patch_size=patch_size,
                                          softmax_scale=softmax_scale, use_windowed_attn=use_windowed_attn,
                                          use_proj_bias=use_proj_bias, use_fused_proj=use_fused_proj)
        self.rms_norm = RMSNorm(d_model=d_model, eps=eps)
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, x: torch.Tensor, use_mqa: bool, use_qk_norm: bool, left_window: int, right_window: int
                ) -> torch.Tensor:
        return x + self.dropout(self.attention(self.rms_norm(x), use_mqa=use_mqa, use_qk_norm=use_qk_norm,
                                               left_window=left_window, right_window=right_window))
