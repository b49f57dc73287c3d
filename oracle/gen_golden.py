"""Generate tests/golden/*.pt by RUNNING THE UNMODIFIED REFERENCE (TEST INFRASTRUCTURE; authoring container only).

    python oracle/gen_golden.py            # rewrites tests/golden/

Each fixture holds: the reference module's constructor args and state_dict, the seeded input, the forward kwargs,
every `F.scaled_dot_product_attention` call the module made (inputs entering the third-party call and its output),
and the module output.  The fixtures pin the oracle (tests/test_oracle_golden.py) and, on the GPU, the drop-in
modules and the op (tests/test_gpu_modules.py).  Reference revision: the tree mounted at /root/reference.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import reference as ref  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _pack_calls(calls):
    packed = []
    for c in calls:
        packed.append({k: (v if not torch.is_tensor(v) else v.contiguous()) for k, v in c.items()})
    return packed


def llm_case(name, seed, d_model, H, G, theta, scale, B, T, left, right, causal, pad, use_qk_norm=True, fused=True):
    m = ref.llm()
    torch.manual_seed(seed)
    attn = m.Attention(d_model, H, G, theta, scale, False, fused)
    x = torch.randn(B, T, d_model)
    padding_mask = None
    if pad:
        padding_mask = torch.rand(B, T) > 0.3
        padding_mask[:, 0] = True
    with torch.no_grad(), ref.capture_sdpa(m) as calls:
        out, _ = attn(x, left, right, causal, padding_mask, None, None, False, False, use_qk_norm)
    torch.save({
        "kind": "llm", "ctor": dict(d_model=d_model, num_heads=H, query_groups=G, theta=theta, softmax_scale=scale,
                                    use_proj_bias=False, use_qkv_proj=fused),
        "state_dict": {k: v.clone() for k, v in attn.state_dict().items()},
        "x": x, "kwargs": dict(left_window=left, right_window=right, causal=causal, use_qk_norm=use_qk_norm),
        "padding_mask": padding_mask, "sdpa_calls": _pack_calls(calls), "out": out,
    }, os.path.join(OUT, name + ".pt"))
    print(name, tuple(out.shape), len(calls), "sdpa call(s)")


def vit2d_case(name, seed, d_model, H, G, target, patch, B, use_qk_norm=True, windowed=False):
    m = ref.vit2d()
    torch.manual_seed(seed)
    hd = d_model // H
    attn = m.SpatialAttention(d_model, H, G, 10000.0, target, patch, 1.0 / hd ** 0.5, windowed, False, True)
    T = (target // patch) ** 2
    x = torch.randn(B, T, d_model)
    with torch.no_grad(), ref.capture_sdpa(m) as calls:
        out = attn(x, False, use_qk_norm, 4, 4)
    torch.save({
        "kind": "vit2d", "ctor": dict(d_model=d_model, num_heads=H, query_groups=G, rope_theta=10000.0,
                                      target_size=target, patch_size=patch, softmax_scale=1.0 / hd ** 0.5,
                                      use_windowed_attn=windowed, use_proj_bias=False, use_fused_proj=True),
        "state_dict": {k: v.clone() for k, v in attn.state_dict().items()},
        "x": x, "kwargs": dict(use_mqa=False, use_qk_norm=use_qk_norm, left_window=4, right_window=4),
        "sdpa_calls": _pack_calls(calls), "out": out,
    }, os.path.join(OUT, name + ".pt"))
    print(name, tuple(out.shape), len(calls), "sdpa call(s)")


def vit3d_case(name, seed, d_model, H, G, grid, B, pad):
    m = ref.vit3d()
    torch.manual_seed(seed)
    attn = m.SpatioTemporalAttention(d_model, H, G, 10000.0, (2, 16, 16))
    gt, gh, gw = grid
    x = torch.randn(B, gt, gh * gw, d_model)
    padding_mask = None
    if pad:
        padding_mask = torch.rand(B, gt * gh * gw) > 0.25
        padding_mask[:, ::gt] = True  # keep at least one valid key per re-viewed row
        padding_mask[:, : gh * gw : 1][:, 0] = True
    with torch.no_grad(), ref.capture_sdpa(m) as calls:
        out = attn(x, grid, False, True, (-1, -1), padding_mask)
    torch.save({
        "kind": "vit3d", "ctor": dict(d_model=d_model, num_heads=H, query_groups=G, rope_theta=10000.0,
                                      patch_size=(2, 16, 16)),
        "state_dict": {k: v.clone() for k, v in attn.state_dict().items()},
        "x": x, "kwargs": dict(grid_size=grid, use_mqa=False, use_qk_norm=True, window_size=(-1, -1)),
        "padding_mask": padding_mask, "sdpa_calls": _pack_calls(calls), "out": out,
    }, os.path.join(OUT, name + ".pt"))
    print(name, tuple(out.shape), len(calls), "sdpa call(s)")


def prepare_case(name, seed, H, G, hd, theta, P, B, use_qk_norm=True):
    """Pre-core producers of the LLM module (src/optimized_attention.py:463-474): apply_qk_norm then RoPE, run by the
    reference on a [B, P+1, heads, hd] tensor; the fixture keeps the LAST position (index P) — what a cached decode
    step has to produce for its new token — plus the cos / sin rows the reference used."""
    m = ref.llm()
    from utils.attention_utils import apply_qk_norm
    torch.manual_seed(seed)
    rope = m.RoPE(hd, theta)
    q = torch.randn(B, P + 1, H, hd)
    k = torch.randn(B, P + 1, G, hd)
    with torch.no_grad():
        qn, kn = apply_qk_norm(q, k) if use_qk_norm else (q, k)
        qr, kr = rope(qn), rope(kn)
    torch.save({
        "kind": "prepare", "H": H, "G": G, "hd": hd, "theta": theta, "position": P, "use_qk_norm": use_qk_norm,
        "q_in": q[:, P].clone(), "k_in": k[:, P].clone(), "q_out": qr[:, P].clone(), "k_out": kr[:, P].clone(),
        "cos_row": rope.cos_cache[P].clone(), "sin_row": rope.sin_cache[P].clone(),
    }, os.path.join(OUT, name + ".pt"))
    print(name, tuple(qr[:, P].shape))


def prepare_seq_case(name, seed, H, G, hd, theta, T, B, use_qk_norm=True):
    """The same producers over a whole sequence (positions 0 .. T-1), as the reference's prefill runs them."""
    m = ref.llm()
    from utils.attention_utils import apply_qk_norm
    torch.manual_seed(seed)
    rope = m.RoPE(hd, theta)
    q = torch.randn(B, T, H, hd)
    k = torch.randn(B, T, G, hd)
    with torch.no_grad():
        qn, kn = apply_qk_norm(q, k) if use_qk_norm else (q, k)
        qr, kr = rope(qn), rope(kn)
    torch.save({"kind": "prepare_seq", "H": H, "G": G, "hd": hd, "theta": theta, "T": T, "use_qk_norm": use_qk_norm,
                "q_in": q, "k_in": k, "q_out": qr, "k_out": kr}, os.path.join(OUT, name + ".pt"))
    print(name, tuple(qr.shape))


def cross_case(name, seed, d_model, H, B, Tq, Tk, pad):
    """Image-generation cross-attention block (queries = image tokens, keys / values = text tokens, key padding)."""
    m = ref.imagegen_cross()
    torch.manual_seed(seed)
    hd = d_model // H
    blk = m.CrossAttentionBlock(d_model, H, 1.0 / hd ** 0.5, False, 1e-7, 0.0).eval()
    x = torch.randn(B, Tq, d_model)
    text = torch.randn(B, Tk, d_model)
    padding_mask = None
    if pad:
        padding_mask = torch.rand(B, Tk) > 0.3
        padding_mask[:, 0] = True
    with torch.no_grad(), ref.capture_sdpa(m) as calls:
        out = blk(x, text, padding_mask)
    torch.save({
        "kind": "cross", "ctor": dict(d_model=d_model, num_heads=H, softmax_scale=1.0 / hd ** 0.5, use_proj_bias=False,
                                      eps=1e-7, dropout=0.0),
        "state_dict": {k: v.clone() for k, v in blk.state_dict().items()},
        "x": x, "text": text, "padding_mask": padding_mask, "sdpa_calls": _pack_calls(calls), "out": out,
    }, os.path.join(OUT, name + ".pt"))
    print(name, tuple(out.shape), len(calls), "sdpa call(s)")


def site_case(name, kind, seed, build, run):
    """A remaining attention call site of the reference (image-gen self-attention, text encoder, video-gen factorized
    self- / cross-attention): the module is run unmodified and every SDPA call (inputs, mask, output) is kept — the
    fixtures pin `vats_multimodal_lm_b200.sdpa_adapter.sdpa_drop_in` on the exact tensors those call sites produce."""
    m = ref.site(kind)
    torch.manual_seed(seed)
    mod = build(m).eval()
    with torch.no_grad(), ref.capture_sdpa(m) as calls:
        out = run(mod)
    torch.save({"kind": "site_" + kind, "sdpa_calls": _pack_calls(calls), "out": out},
               os.path.join(OUT, name + ".pt"))
    print(name, tuple(out.shape), len(calls), "sdpa call(s)",
          [(tuple(c["q"].shape), None if c["attn_mask"] is None else tuple(c["attn_mask"].shape), c["is_causal"]) for c in calls])


def site_cases():
    B = 2
    pm = lambda T, seed: (torch.rand(B, T, generator=torch.Generator().manual_seed(seed)) > 0.3)

    def keep_first(mask):
        mask = mask.clone()
        mask[:, 0] = True
        return mask

    # image-generation causal self-attention: key padding & causal triangle (materialised AND), and plain is_causal
    site_case("site_imagegen_self_pad", "imagegen_self", 601,
              lambda m: m.CausalSelfAttention(128, 4, 2, 10000.0, 32 ** -0.5, False, True, False, False),
              lambda mod: mod(torch.randn(B, 64, 128), False, True, True, -1, -1, False, keep_first(pm(64, 1))))
    site_case("site_imagegen_self_ntk", "imagegen_self", 602,
              lambda m: m.CausalSelfAttention(192, 4, 4, 10000.0, 48 ** -0.5, False, False, False, True, 2.0),
              lambda mod: mod(torch.randn(B, 36, 192), False, True, True, -1, -1, False, None))
    # text encoder: bidirectional, key padding (an .expand() view)
    site_case("site_text_encoder_pad", "text_encoder", 603,
              lambda m: m.Attention(128, 4, 2, 10000.0, 32 ** -0.5, False, True),
              lambda mod: mod(torch.randn(B, 50, 128), False, True, keep_first(pm(50, 2))))
    site_case("site_text_encoder_mqa", "text_encoder", 604,
              lambda m: m.Attention(128, 4, 1, 10000.0, 32 ** -0.5, False, True),
              lambda mod: mod(torch.randn(B, 40, 128), True, True, None))
    # video-generation factorized causal self-attention: query-row padding x causal triangle, spatial then temporal
    site_case("site_videogen_self_pad", "videogen_self", 605,
              lambda m: m.CausalFactorizedAttention(128, 4, 2, 10000.0, 32 ** -0.5, False, True, False, False),
              lambda mod: mod(torch.randn(B, 4, 9, 128), False, True, True, -1, -1, False,
                              torch.ones(B, 36, dtype=torch.bool) & (torch.rand(B, 36) > 0.2)))
    site_case("site_videogen_self", "videogen_self", 606,
              lambda m: m.CausalFactorizedAttention(128, 4, 2, 10000.0, 32 ** -0.5, False, True, False, False),
              lambda mod: mod(torch.randn(B, 4, 9, 128), False, True, True, -1, -1, False, None))
    # video-generation factorized cross-attention: text keys with key padding
    site_case("site_videogen_cross_pad", "videogen_cross", 607,
              lambda m: m.FactorizedCrossAttention(128, 4, 2, 32 ** -0.5, False),
              lambda mod: mod(torch.randn(B, 4, 9, 128), torch.randn(B, 12, 128), False, True, keep_first(pm(12, 3))))


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--sites" in sys.argv:
        site_cases()
        return
    cross_case("cross_hd16_pad", 501, 128, 8, 2, 150, 16, True)
    cross_case("cross_hd64", 502, 128, 2, 1, 40, 70, False)
    prepare_seq_case("prepareseq_hd60_t40", 403, 6, 2, 60, 10000.0, 40, 2)
    # pre-core step of a decode token (qk-norm + RoPE at position P)
    prepare_case("prepare_hd128_p300", 401, 8, 2, 128, 10000.0, 300, 3)
    prepare_case("prepare_hd60_p17_nonorm", 402, 6, 2, 60, 10000.0, 17, 2, use_qk_norm=False)
    # LLM — xsmall-like geometry (hd = 16, softmax_scale = sqrt(16) as in model_args_xsmall.py:31)
    llm_case("llm_hd16_causal", 101, 64, 4, 2, 10000.0, 4.0, 2, 16, 128, 0, True, False)
    llm_case("llm_hd16_causal_pad", 102, 64, 4, 2, 10000.0, 4.0, 3, 16, 128, 0, True, True)
    llm_case("llm_hd16_noncausal_pad", 103, 64, 4, 1, 10000.0, 4.0, 2, 16, -1, -1, False, True)
    # LLM — medium head geometry (hd = 60, H/G = 3) at a reduced width; window smaller than T (dropped by the reference)
    llm_case("llm_hd60_causal_window", 104, 180, 3, 1, 10000.0, 60 ** -0.5, 2, 48, 8, 0, True, False)
    llm_case("llm_hd60_nonorm_unfused", 105, 180, 3, 1, 10000.0, 60 ** -0.5, 1, 33, -1, -1, True, False,
             use_qk_norm=False, fused=False)
    # LLM — large head geometry (hd = 128, H/G = 4), one 128-token block + remainder
    llm_case("llm_hd128_causal", 106, 256, 2, 1, 10000.0, 128 ** -0.5, 1, 160, 64, 0, True, True)
    # ViT-2D — hd = 72 (medium) and hd = 48 (xsmall)
    vit2d_case("vit2d_hd72", 201, 144, 2, 1, 64, 16, 3)
    vit2d_case("vit2d_hd48_windowed", 202, 96, 2, 1, 96, 16, 2, windowed=True)
    # ViT-3D — hd = 66 (large), spatial 3x3 / temporal 4, with and without key padding
    vit3d_case("vit3d_hd66", 301, 132, 2, 1, (4, 3, 3), 2, False)
    vit3d_case("vit3d_hd66_pad", 302, 132, 2, 1, (4, 3, 3), 2, True)
    vit3d_case("vit3d_hd60_grid", 303, 120, 2, 2, (2, 6, 6), 1, True)
    site_cases()


if __name__ == "__main__":
    main()
