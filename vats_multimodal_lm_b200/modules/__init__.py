"""Drop-in `nn.Module`s with the reference's constructor / forward signatures."""
from ._common import WINDOW_MODES, get_default_window_mode, set_default_window_mode  # noqa: F401
from .llm import Attention, AttentionBlock, KVCache, RMSNorm, RoPE  # noqa: F401
from .vit2d import RoPE as RoPE2D, SpatialAttention, SpatialAttentionBlock  # noqa: F401
from .vit3d import RoPE3D, SpatioTemporalAttention, SpatioTemporalAttentionBlock  # noqa: F401
from .cross import CrossAttention, CrossAttentionBlock  # noqa: F401
