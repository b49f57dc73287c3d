"""vats_multimodal_lm_b200 — B200-native (sm_100a) GQA + sliding-window attention core for
S-VATS31/vats-multimodal-lm: hand-written CUDA kernels behind a C-ABI library, `torch.library` custom ops, and
drop-in modules that keep the reference's signatures.  No CPU path: importing works anywhere, computing needs a B200.
"""
from . import _ffi, ops  # noqa: F401  (ops registers torch.ops.vats.*)
from .modules import (  # noqa: F401
    Attention, AttentionBlock, KVCache, RMSNorm, RoPE, RoPE2D, RoPE3D, SpatialAttention, SpatialAttentionBlock,
    SpatioTemporalAttention, SpatioTemporalAttentionBlock, CrossAttention, CrossAttentionBlock,
    get_default_window_mode, set_default_window_mode,
)
from .ops import gqa_swa_decode, gqa_swa_prefill  # noqa: F401

__version__ = "0.1.0"
