"""CPU: host-side logic of the drop-in modules — signatures, parameter names, error conventions, the pre-core
producers (projections, qk-norm, RoPE 1D/2D/3D) against the golden fixtures, KVCache bookkeeping."""
import inspect
import math

import pytest
import torch

from conftest import load_golden
import vats_multimodal_lm_b200 as vl
from vats_multimodal_lm_b200.modules import llm as L, vit2d as V2, vit3d as V3
from vats_multimodal_lm_b200.modules._common import apply_qk_norm


def _unexpand(x, G):
    H = x.size(1)
    return x[:, :: H // G].permute(0, 2, 1, 3).contiguous()


def test_signatures_match_reference():
    # reference src/optimized_attention.py:302-311, 355-367
    assert list(inspect.signature(vl.Attention.__init__).parameters)[1:8] == [
        "d_model", "num_heads", "query_groups", "theta", "softmax_scale", "use_proj_bias", "use_qkv_proj"]
    assert list(inspect.signature(vl.Attention.forward).parameters)[1:] == [
        "x", "left_window", "right_window", "causal", "padding_mask", "kv_cache", "layer_idx", "use_cache", "use_mqa",
        "use_qk_norm"]
    # reference :746-757
    assert list(inspect.signature(vl.AttentionBlock.__init__).parameters)[1:] == [
        "d_model", "num_heads", "query_groups", "softmax_scale", "use_proj_bias", "use_qkv_proj", "dropout", "theta",
        "eps"]
    # reference :179-186
    assert list(inspect.signature(vl.KVCache.__init__).parameters)[1:6] == [
        "max_batch_size", "max_seq_len", "num_heads", "head_dim", "num_layers"]
    # reference vit_2d/optimized_attention.py:214-226, 580-587
    assert list(inspect.signature(vl.SpatialAttention.__init__).parameters)[1:11] == [
        "d_model", "num_heads", "query_groups", "rope_theta", "target_size", "patch_size", "softmax_scale",
        "use_windowed_attn", "use_proj_bias", "use_fused_proj"]
    assert list(inspect.signature(vl.SpatialAttention.forward).parameters)[1:] == [
        "x", "use_mqa", "use_qk_norm", "left_window", "right_window"]
    # reference vit_3d/optimized_attention.py:30-37, 617-625
    assert list(inspect.signature(vl.SpatioTemporalAttention.__init__).parameters)[1:6] == [
        "d_model", "num_heads", "query_groups", "rope_theta", "patch_size"]
    assert list(inspect.signature(vl.SpatioTemporalAttention.forward).parameters)[1:] == [
        "x", "grid_size", "use_mqa", "use_qk_norm", "window_size", "padding_mask"]


def test_state_dict_names_and_shapes():
    a = vl.Attention(1440, 24, 8, 10000.0, 60 ** -0.5)
    sd = a.state_dict()
    assert sd["w_qkv.weight"].shape == (2400, 1440) and sd["w_o.weight"].shape == (1440, 1440)
    assert {"rope.inv_freq", "rope.cos_cache", "rope.sin_cache"} <= set(sd)
    a2 = vl.Attention(64, 4, 2, 10000.0, 0.25, use_qkv_proj=False)
    assert {"w_q.weight", "w_k.weight", "w_v.weight", "w_o.weight"} <= set(a2.state_dict())
    s = vl.SpatialAttention(768, 16, 8, 10000.0, 384, 16, 48 ** -0.5, False, False, True)
    assert s.state_dict()["qkv_proj.weight"].shape == (768 + 2 * 8 * 48, 768)
    assert "o_proj.weight" in s.state_dict() and "rope.inv_freq" in s.state_dict()
    t = vl.SpatioTemporalAttention(2112, 32, 8, 10000.0, (2, 16, 16))
    assert t.state_dict()["w_qkv.weight"].shape == (2112 + 2 * 8 * 66, 2112)
    assert {"rope.freqs_t", "rope.freqs_h", "rope.freqs_w"} <= set(t.state_dict())


def test_error_conventions():
    with pytest.raises(ValueError):
        vl.Attention(100, 3, 1, 1e4, 1.0)          # d_model % num_heads
    with pytest.raises(ValueError):
        vl.Attention(64, 4, 3, 1e4, 1.0)           # num_heads % query_groups
    with pytest.raises(ValueError):
        vl.RoPE(7, 1e4)
    with pytest.raises(ValueError):
        vl.RoPE2D(6, 64, 16, 1e4)
    with pytest.raises(ValueError):
        vl.RoPE3D(8, 1e4, (2, 16, 16))
    a = vl.Attention(64, 4, 2, 1e4, 0.25)
    with pytest.raises(ValueError):
        a(torch.zeros(2, 5, 32), -1, -1)           # wrong d_model
    out, cache = a(torch.zeros(2, 0, 64), -1, -1)  # T == 0 (reference :405-407)
    assert out.shape == (2, 0, 64) and cache is None
    with pytest.raises(ValueError):                # bad padding mask shape (reference :669-672)
        a(torch.zeros(2, 5, 64), -1, -1, True, torch.ones(2, 4, dtype=torch.bool))


def test_cpu_forward_fails_loudly_not_silently():
    a = vl.Attention(64, 4, 2, 1e4, 0.25)
    with pytest.raises(RuntimeError, match="no CPU implementation|CUDA"):
        a(torch.randn(1, 4, 64), -1, -1)


@pytest.mark.parametrize("fname", ["llm_hd16_causal.pt", "llm_hd60_causal_window.pt", "llm_hd60_nonorm_unfused.pt",
                                   "llm_hd128_causal.pt"])
def test_llm_precore_matches_reference_capture(fname):
    fx = load_golden(fname)
    a = vl.Attention(**fx["ctor"])
    a.load_state_dict(fx["state_dict"])
    x = fx["x"]
    B, T, _ = x.shape
    H, G, hd = a.num_heads, a.query_groups, a.head_dim
    if a.use_qkv_proj:
        q, k, v = torch.split(a.w_qkv(x), [H * hd, G * hd, G * hd], dim=-1)
    else:
        q, k, v = a.w_q(x), a.w_k(x), a.w_v(x)
    q, k, v = q.view(B, T, H, hd), k.view(B, T, G, hd), v.view(B, T, G, hd)
    if fx["kwargs"]["use_qk_norm"]:
        q, k = apply_qk_norm(q, k)
    q, k = a.rope(q), a.rope(k)
    call = fx["sdpa_calls"][0]
    torch.testing.assert_close(q, call["q"].permute(0, 2, 1, 3), atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(k, _unexpand(call["k"], G), atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(v, _unexpand(call["v"], G), atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("fname", ["vit2d_hd72.pt", "vit2d_hd48_windowed.pt"])
def test_vit2d_precore_matches_reference_capture(fname):
    fx = load_golden(fname)
    m = vl.SpatialAttention(**fx["ctor"])
    m.load_state_dict(fx["state_dict"])
    q, k, v = m._setup_qkv(fx["x"], use_mqa=False, use_qk_norm=fx["kwargs"]["use_qk_norm"])
    call = fx["sdpa_calls"][0]
    G = m.query_groups
    torch.testing.assert_close(q, call["q"].permute(0, 2, 1, 3), atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(k, _unexpand(call["k"], G), atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(v, _unexpand(call["v"], G), atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("fname", ["vit3d_hd66.pt", "vit3d_hd66_pad.pt", "vit3d_hd60_grid.pt"])
def test_vit3d_spatial_precore_matches_reference_capture(fname):
    fx = load_golden(fname)
    m = vl.SpatioTemporalAttention(**fx["ctor"])
    m.load_state_dict(fx["state_dict"])
    grid = fx["kwargs"]["grid_size"]
    q, k, v = m._setup_qkv(fx["x"], False, True, grid, "spatial")
    call = fx["sdpa_calls"][0]
    G = m.query_groups
    torch.testing.assert_close(q, call["q"].permute(0, 2, 1, 3), atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(k, _unexpand(call["k"], G), atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(v, _unexpand(call["v"], G), atol=1e-6, rtol=1e-5)
    # temporal pass producers, fed with the reference's own spatial output
    B = fx["x"].size(0)
    sp = call["out"].permute(0, 2, 1, 3).reshape(B * grid[0], -1, m.d_model).view(B, grid[0], -1, m.d_model)
    q2, k2, v2 = (t.reshape(-1, *t.shape[2:]) for t in m._setup_qkv(sp, False, True, grid, "temporal"))  # [B,S,T,..] views
    call2 = fx["sdpa_calls"][1]
    torch.testing.assert_close(q2, call2["q"].permute(0, 2, 1, 3), atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(k2, _unexpand(call2["k"], G), atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(v2, _unexpand(call2["v"], G), atol=1e-5, rtol=1e-4)


def test_rope_offset_continues_positions():
    r = vl.RoPE(16, 10000.0)
    x = torch.randn(2, 6, 3, 16)
    full = r(x)
    torch.testing.assert_close(r(x[:, 4:], offset=4), full[:, 4:])


def test_kvcache_api():
    c = vl.KVCache(max_batch_size=4, max_seq_len=8, num_heads=2, head_dim=4, num_layers=3, dtype=torch.float32)
    assert c.cache is None and c.current_seq_len is None
    c.initialize(2)
    assert c.current_seq_len == 0 and c.cache[1]["k"].shape == (2, 8, 2, 4)
    k = torch.arange(2 * 3 * 2 * 4, dtype=torch.float32).view(2, 3, 2, 4)
    for layer in range(3):
        c.update(layer, k, k + 1)
    assert c.current_seq_len == 3                      # per-layer lengths: three layers, three tokens — not nine
    kk, vv = c.get(1, 3)
    assert torch.equal(kk, k) and torch.equal(vv, k + 1)
    assert c.get(1, 4) == (None, None)
    c.update(0, torch.ones(2, 7, 2, 4), torch.ones(2, 7, 2, 4))   # truncates at max_seq_len (reference :241-250)
    assert c.layer_seq_len(0) == 8
    c.update(0, torch.ones(2, 1, 2, 4), torch.ones(2, 1, 2, 4))   # full: silently dropped (reference :244-245)
    assert c.layer_seq_len(0) == 8
    with pytest.raises(ValueError):
        c.initialize(5)
    c.reset()
    assert c.cache is None and c.current_seq_len is None and c.batch_size is None


def test_rope_cache_loads_reference_style_state_dict():
    a = vl.Attention(64, 4, 2, 1e4, 0.25)
    sd = a.state_dict()
    sd["rope.cos_cache"] = torch.zeros(5, 8)
    sd["rope.sin_cache"] = torch.zeros(5, 8)
    a.load_state_dict(sd)  # must not raise on the shape change
    assert a.rope.cached_seq_len == 5
