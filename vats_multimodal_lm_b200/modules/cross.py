"""Drop-in replacements for the reference's image-generation cross-attention
(src/autoregressive_image_gen/autoregressive_transformer/attention/cross_attention.py: CrossAttention :12-238,
CrossAttentionBlock :240-280) — the first of the "remaining call sites" of SURVEY.md §8f rank 3.

Same constructor / forward signatures and parameter names (`q_proj`, `k_proj`, `v_proj`, `o_proj`, `rms_norm`).
Image tokens (queries, Tq = H*W) attend text tokens (keys / values, Tk = T) without a causal mask; the optional
`padding_mask [B, Tk]` masks KEYS (reference :73-82, `attn_mask = padding_mask[:, None, None, :]`), which is the op's
`k_valid`.  Plain multi-head attention: the op runs it as GQA with G == H.  Unlike the ViT modules the reference passes
`self.softmax_scale` to SDPA here (:93-101), and so does the drop-in.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ._common import attention_core
from .llm import RMSNorm


class CrossAttention(nn.Module):
    def __init__(self, d_model: int, num_heads: int, softmax_scale: float, use_proj_bias: bool):
        super().__init__()
        if d_model % num_heads != 0:
            raise ValueError(f"d_model must be divisble by num_heads, got {d_model} % {num_heads} != 0.")
        self.d_model = d_model
        self.num_heads = num_heads
        self.softmax_scale = softmax_scale
        self.head_dim = d_model // num_heads
        self.q_proj = nn.Linear(d_model, d_model, bias=use_proj_bias)
        self.k_proj = nn.Linear(d_model, d_model, bias=use_proj_bias)
        self.v_proj = nn.Linear(d_model, d_model, bias=use_proj_bias)
        self.o_proj = nn.Linear(d_model, d_model, bias=use_proj_bias)

    def forward(self, x: torch.Tensor, text_embeddings: torch.Tensor,
                padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, Tq, _ = x.shape
        Tk = text_embeddings.size(1)
        assert x.size(0) == text_embeddings.size(0), "First dims of inputs should be equal."
        assert x.size(-1) == text_embeddings.size(-1), "Last dims of inputs should be equal."
        q = self.q_proj(x).view(B, Tq, self.num_heads, self.head_dim)
        k = self.k_proj(text_embeddings).view(B, Tk, self.num_heads, self.head_dim)
        v = self.v_proj(text_embeddings).view(B, Tk, self.num_heads, self.head_dim)
        k_valid = None
        if padding_mask is not None:
            k_valid = padding_mask.bool()
            assert k_valid.shape == (B, Tk), f"expected {B, Tk}, got {k_valid.shape}"
        o = attention_core(q, k, v, scale=self.softmax_scale, causal=False, left=-1, right=-1, k_valid=k_valid,
                           out_dtype=x.dtype)
        return self.o_proj(o.reshape(B, Tq, self.d_model))


class CrossAttentionBlock(nn.Module):
    """x + dropout(cross_attention(rms_norm(x), text, padding_mask)) (reference :240-280)."""

    def __init__(self, d_model: int, num_heads: int, softmax_scale: float, use_proj_bias: bool, eps: float,
                 dropout: float):
        super().__init__()
        self.cross_attention = CrossAttention(d_model=d_model, num_heads=num_heads, softmax_scale=softmax_scale,
                                              use_proj_bias=use_proj_bias)
        self.rms_norm = RMSNorm(d_model, eps)
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, x: torch.Tensor, text_embeddings: torch.Tensor,
                padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        return x + self.dropout(self.cross_attention(self.rms_norm(x), text_embeddings=text_embeddings,
                                                     padding_mask=padding_mask))
