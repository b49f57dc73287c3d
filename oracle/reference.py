"""Importer for the UNMODIFIED reference modules (TEST INFRASTRUCTURE; authoring container only).

/root/reference is read-only and does not exist on the GPU box: nothing under `tests -m gpu`, `smoke()` or
`bench.py` may call into this file at run time.  It is used by `oracle/gen_golden.py` (to produce the committed
fixtures) and by CPU tests that are skipped when the reference tree is absent.
"""
from __future__ import annotations

import contextlib
import os
import sys
from typing import Any, Dict, List

REFERENCE_ROOT = os.environ.get("VATS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "optimized_attention.py"))


def _ensure_path() -> None:
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the mount is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def llm():
    """reference src/optimized_attention.py (Attention, AttentionBlock, KVCache, RoPE)"""
    _ensure_path()
    import src.optimized_attention as m
    return m


def vit2d():
    """reference src/transformers/vision/vit_2d/optimized_attention.py"""
    _ensure_path()
    import src.transformers.vision.vit_2d.optimized_attention as m
    return m


def vit3d():
    """reference src/transformers/vision/vit_3d/optimized_attention.py"""
    _ensure_path()
    import src.transformers.vision.vit_3d.optimized_attention as m
    return m


def imagegen_cross():
    """reference src/autoregressive_image_gen/autoregressive_transformer/attention/cross_attention.py"""
    _ensure_path()
    import src.autoregressive_image_gen.autoregressive_transformer.attention.cross_attention as m
    return m


def site(name: str):
    """One of the remaining attention call sites of the reference by short name."""
    _ensure_path()
    import importlib
    return importlib.import_module({
        "imagegen_self": "src.autoregressive_image_gen.autoregressive_transformer.attention.optimized_attention",
        "text_encoder": "src.autoregressive_image_gen.text_encoder.encoder_attention",
        "videogen_self": "src.autoregressive_video_gen.autoregressive_transformer.attention.optimized_attention",
        "videogen_cross": "src.autoregressive_video_gen.autoregressive_transformer.attention.cross_attention",
    }[name])


@contextlib.contextmanager
def capture_sdpa(module) -> "contextlib.AbstractContextManager[List[Dict[str, Any]]]":
    """Record every `F.scaled_dot_product_attention` call the reference module makes: the tensors entering the
    third-party call (post qk-norm, post RoPE, K/V already expanded to H heads) and its result."""
    import torch.nn.functional as F

    calls: List[Dict[str, Any]] = []
    real = F.scaled_dot_product_attention

    class _Shim:
        def __getattr__(self, name):
            return getattr(F, name)

        @staticmethod
        def scaled_dot_product_attention(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False,
                                         scale=None, enable_gqa=False):
            out = real(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale,
                       enable_gqa=enable_gqa)
            calls.append({
                "q": query.detach().clone(), "k": key.detach().clone(), "v": value.detach().clone(),
                "attn_mask": None if attn_mask is None else attn_mask.detach().clone(),
                "is_causal": bool(is_causal), "scale": scale, "out": out.detach().clone(),
            })
            return out

    saved = module.F
    module.F = _Shim()
    try:
        yield calls
    finally:
        module.F = saved
