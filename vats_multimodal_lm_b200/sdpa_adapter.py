"""Call-level drop-in for the reference's one library call, `F.scaled_dot_product_attention`.

The LLM, ViT-2D, ViT-3D and image-gen cross-attention modules have class-level drop-ins (`modules/`).  The remaining
attention call sites of the reference —

    src/autoregressive_image_gen/autoregressive_transformer/attention/optimized_attention.py:214-321   CausalSelfAttention
    src/autoregressive_image_gen/text_encoder/encoder_attention.py:224-311                              text-encoder Attention
    src/autoregressive_video_gen/autoregressive_transformer/attention/optimized_attention.py:155-313   CausalFactorizedAttention
    src/autoregressive_video_gen/autoregressive_transformer/attention/cross_attention.py:61-170        FactorizedCrossAttention

— keep their own projections, NTK / 3-D RoPE variants, cache plumbing and mask re-views (all of it the reference's code,
quirks included) and differ only in the masks they hand to SDPA.  They all reach the kernel through the module-level
name `F`, so replacing that name with `FunctionalShim` routes exactly the SDPA call to `torch.ops.vats.gqa_swa_prefill`
and nothing else.  `sdpa_drop_in` has SDPA's signature and semantics for what those call sites pass:

  * q [B, H, Tq, hd], k / v [B, Hk, Tk, hd] with Hk == H (K/V already expanded by `extend_kv_heads`) or Hk == 1
    (the MQA shortcut) — transposed VIEWS of [B, T, heads, hd] tensors, which the op takes by stride (no copy);
  * `is_causal=True` without a mask (Tq == Tk), or a boolean `attn_mask` that is a product of a query-row mask, a key
    mask and optionally the causal triangle — the only forms the reference builds (key padding: image-gen :236-244,
    text-encoder :279-290; query-row padding (+ tril): video-gen :178-236; key padding: cross-attention :112-130).  The
    mask is decomposed into (q_valid, k_valid, causal) — by its strides when it is an `.expand()` view, else by
    comparing it with the product of its row / column supports — and anything else raises: no silent fallback.
  * a row with no allowed key yields zeros (torch >= 2.5 SDPA semantics, as in the op).

The K/V head count passed to the op is Hk: the reference has already written the expanded copy, reading it is what
this call site costs (the class-level drop-ins never materialise it).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as _F

from . import ops
from .modules._common import _to_kernel_layout

__all__ = ["sdpa_drop_in", "decompose_mask", "FunctionalShim", "install", "uninstall", "SDPA_CALL_SITES"]

# modules of the reference whose `F.scaled_dot_product_attention` call is rerouted by `integration.patch_reference`
SDPA_CALL_SITES = (
    "src.autoregressive_image_gen.autoregressive_transformer.attention.optimized_attention",
    "src.autoregressive_image_gen.text_encoder.encoder_attention",
    "src.autoregressive_video_gen.autoregressive_transformer.attention.optimized_attention",
    "src.autoregressive_video_gen.autoregressive_transformer.attention.cross_attention",
)


def decompose_mask(attn_mask: torch.Tensor, B: int, Tq: int, Tk: int
                   ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], bool]:
    """bool attn_mask broadcastable to [B, H, Tq, Tk] -> (q_valid [B,Tq] | None, k_valid [B,Tk] | None, causal)."""
    if attn_mask.dtype != torch.bool:
        raise NotImplementedError("only boolean attention masks are supported (the reference builds no additive mask)")
    m = attn_mask
    while m.dim() < 4:
        m = m[None]
    if m.size(1) > 1 and m.stride(1) != 0 and not bool((m[:, :1] == m).all()):
        raise NotImplementedError("attn_mask differs between heads")
    m3 = m[:, 0].expand(B, Tq, Tk)
    if m3.stride(1) == 0 or Tq == 1:          # the same for every query row: a key mask
        kv = m3[:, 0, :]
        return None, (None if bool(kv.all()) else kv.contiguous()), False
    if m3.stride(2) == 0 or Tk == 1:          # the same for every key: a query-row mask
        qv = m3[:, :, 0]
        return (None if bool(qv.all()) else qv.contiguous()), None, False
    qv, kv = m3.any(dim=2), m3.any(dim=1)
    outer = qv[:, :, None] & kv[:, None, :]
    causal = None
    if torch.equal(m3, outer):
        causal = False
    elif Tq == Tk:
        tril = torch.ones(Tq, Tk, dtype=torch.bool, device=m3.device).tril()
        if torch.equal(m3, outer & tril):
            causal = True
    if causal is None:
        raise NotImplementedError("attn_mask is not (query-row mask) x (key mask) [x causal triangle]: the attention "
                                  "kernels have no arbitrary-mask path and there is no fallback")
    return (None if bool(qv.all()) else qv), (None if bool(kv.all()) else kv), causal


def sdpa_drop_in(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, attn_mask: Optional[torch.Tensor] = None,
                 dropout_p: float = 0.0, is_causal: bool = False, scale: Optional[float] = None,
                 enable_gqa: bool = False) -> torch.Tensor:
    """`F.scaled_dot_product_attention` on the sm_100a kernels: [B, H, Tq, hd] in, [B, H, Tq, hd] out (same dtype)."""
    if dropout_p != 0.0:
        raise NotImplementedError("attention dropout is not supported (the reference passes none)")
    if query.dim() != 4 or key.dim() != 4 or value.shape != key.shape:
        raise ValueError("expected q [B,H,Tq,hd] and k, v [B,Hk,Tk,hd]")
    B, H, Tq, hd = query.shape
    Hk, Tk = key.size(1), key.size(2)
    if H % Hk != 0:
        raise ValueError(f"query heads ({H}) must be a multiple of key/value heads ({Hk})")
    if Hk != H and Hk != 1 and not enable_gqa:
        raise ValueError("key/value heads must equal the query heads (or 1) unless enable_gqa=True")
    causal = bool(is_causal)
    q_valid = k_valid = None
    if attn_mask is not None:
        if causal:
            raise ValueError("attn_mask and is_causal=True are mutually exclusive (as in torch)")
        q_valid, k_valid, causal = decompose_mask(attn_mask, B, Tq, Tk)
    if causal and Tq != Tk:
        raise NotImplementedError("torch aligns a causal mask top-left when Tq != Tk; the kernels align bottom-right")
    if scale is None:
        scale = hd ** -0.5
    pad = Tk > 32
    # [B, heads, T, hd] -> [B, T, heads, hd]: a view; the bf16 cast writes the kernel layout
    q4, k4, v4 = (_to_kernel_layout(t.transpose(1, 2), pad) for t in (query, key, value))
    o = ops.gqa_swa_prefill(q4, k4, v4, q_valid, k_valid, float(scale), causal, -1, -1)
    return o.to(query.dtype).transpose(1, 2)


class FunctionalShim:
    """Stands in for the name `F` (torch.nn.functional) of a reference module: everything is passed through except
    `scaled_dot_product_attention`."""

    scaled_dot_product_attention = staticmethod(sdpa_drop_in)

    def __getattr__(self, name):
        return getattr(_F, name)


_installed = {}


def install(module) -> None:
    """Reroute `module.F.scaled_dot_product_attention` (module: an imported reference module) to the kernels."""
    if module.__name__ in _installed:
        return
    _installed[module.__name__] = (module, module.F)
    module.F = FunctionalShim()


def uninstall() -> None:
    for module, saved in _installed.values():
        module.F = saved
    _installed.clear()
