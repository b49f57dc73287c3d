"""Swap the drop-in modules into an importable copy of the reference — no reference source change.

    import sys; sys.path.insert(0, "/path/to/vats-multimodal-lm")
    from vats_multimodal_lm_b200.integration import patch_reference
    patch_reference()                      # before OR after the reference's model files were imported
    from src.transformers.nlp.model import AutoregressiveTextTransformer      # now built from the drop-ins

The reference's model files bind the attention classes by name at import
(`from src.optimized_attention import AttentionBlock, KVCache`, src/transformers/nlp/model.py:12;
vit_2d/model.py:12; vit_3d/model.py:12; autoregressive_*/.../model.py:12), so besides replacing the attributes of the
defining modules `patch_reference` rebinds the same names in every already-imported reference module.
The reference's other attention modules (image-gen causal self-attention, text encoder, video-gen factorized self- and
cross-attention) keep their classes; `patch_reference` reroutes their one `F.scaled_dot_product_attention` call to the
kernels (`sdpa_adapter.FunctionalShim`).  `unpatch_reference()` restores the originals.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, List, Tuple

from . import modules as _m
from . import sdpa_adapter as _sdpa

# defining module of the reference -> {attribute: drop-in}
_TARGETS = {
    "src.optimized_attention": {
        "Attention": _m.Attention, "AttentionBlock": _m.AttentionBlock, "KVCache": _m.KVCache, "RoPE": _m.RoPE,
    },
    "src.transformers.vision.vit_2d.optimized_attention": {
        "SpatialAttention": _m.SpatialAttention, "SpatialAttentionBlock": _m.SpatialAttentionBlock,
    },
    "src.transformers.vision.vit_3d.optimized_attention": {
        "SpatioTemporalAttention": _m.SpatioTemporalAttention,
        "SpatioTemporalAttentionBlock": _m.SpatioTemporalAttentionBlock,
    },
    "src.autoregressive_image_gen.autoregressive_transformer.attention.cross_attention": {
        "CrossAttention": _m.CrossAttention, "CrossAttentionBlock": _m.CrossAttentionBlock,
    },
}

_saved: List[Tuple[object, str, object]] = []   # (module, attribute, original)


def patch_reference(strict: bool = False) -> Dict[str, List[str]]:
    """Replace the reference's attention classes with the drop-ins.  Returns {module name: [patched attributes]}.
    Modules of the reference that cannot be imported are skipped unless `strict`."""
    if _saved:
        return {}
    done: Dict[str, List[str]] = {}
    swaps = {}  # id(original class) -> (original, drop-in)
    for mod_name, attrs in _TARGETS.items():
        try:
            mod = importlib.import_module(mod_name)
        except Exception:
            if strict:
                raise
            continue
        for attr, repl in attrs.items():
            orig = getattr(mod, attr, None)
            if orig is None or orig is repl:
                continue
            swaps[id(orig)] = (orig, repl)
            _saved.append((mod, attr, orig))
            setattr(mod, attr, repl)
            done.setdefault(mod_name, []).append(attr)
    # the remaining attention call sites keep their own classes; only their SDPA call is rerouted (sdpa_adapter.py)
    for mod_name in _sdpa.SDPA_CALL_SITES:
        try:
            mod = importlib.import_module(mod_name)
        except Exception:
            if strict:
                raise
            continue
        _sdpa.install(mod)
        done.setdefault(mod_name, []).append("F.scaled_dot_product_attention")
    # names already bound elsewhere by `from ... import X`
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name == "src" or name.startswith(("src.", "training.", "scripts.", "tests."))):
            continue
        for attr, val in list(vars(mod).items()):
            hit = swaps.get(id(val))
            if hit is not None and val is hit[0]:
                _saved.append((mod, attr, val))
                setattr(mod, attr, hit[1])
                done.setdefault(name, []).append(attr)
    return done


def unpatch_reference() -> None:
    _sdpa.uninstall()
    while _saved:
        mod, attr, orig = _saved.pop()
        setattr(mod, attr, orig)
