"""Pin the oracle: oracle.sdpa / oracle.mask against every golden fixture produced by the unmodified reference
(oracle/gen_golden.py), plus the reference's own numeric property (tests/transformers/nlp/attention_tests.py:111-131).
CPU only."""
import math

import pytest
import torch

from conftest import cross_golden_files, golden_files, load_golden, prepare_golden_files, prepare_seq_golden_files
from oracle import expand_kv, mask_predicate, sdpa_explicit


def _unexpand(x, G):
    """[N,H,T,hd] (reference layout, K/V repeat_interleave'd) -> [N,T,G,hd] keeping one head per group."""
    H = x.size(1)
    return x[:, :: H // G].permute(0, 2, 1, 3).contiguous()


def _cases():
    out = []
    for f in golden_files():
        fx = load_golden(f)
        for ci, call in enumerate(fx["sdpa_calls"]):
            out.append(pytest.param(f, ci, id=f"{f[:-3]}-call{ci}"))
    return out


@pytest.mark.parametrize("fname,ci", _cases())
def test_oracle_matches_reference_sdpa_call(fname, ci):
    fx = load_golden(fname)
    call = fx["sdpa_calls"][ci]
    G = fx["ctor"]["query_groups"]
    q = call["q"].permute(0, 2, 1, 3).contiguous()           # [N,Tq,H,hd]
    N, Tq, H, hd = q.shape
    # the reference expanded K/V with repeat_interleave: check that, then fold back to G heads
    k = _unexpand(call["k"], G)
    v = _unexpand(call["v"], G)
    assert torch.equal(expand_kv(k, H).permute(0, 2, 1, 3), call["k"])
    assert torch.equal(expand_kv(v, H).permute(0, 2, 1, 3), call["v"])
    Tk = k.size(1)
    scale = call["scale"] if call["scale"] is not None else 1.0 / math.sqrt(hd)

    # rebuild the mask the reference handed to SDPA from the predicate
    if fx["kind"] == "llm":
        causal = fx["kwargs"]["causal"]
        qv = fx.get("padding_mask")
        mask = mask_predicate(N, Tq, Tk, causal, -1, -1, q_valid=qv)      # executable path: window dropped
        if call["attn_mask"] is not None:
            assert torch.equal(call["attn_mask"][:, 0], mask), "LLM mask build (reference :668-706) != predicate"
        else:
            assert call["is_causal"] == causal
    elif fx["kind"] == "vit2d":
        assert call["attn_mask"] is None and not call["is_causal"]
        mask = None
    else:  # vit3d: key padding, re-viewed per pass (reference :264-277)
        pm = fx.get("padding_mask")
        if pm is None:
            assert call["attn_mask"] is None
            mask = None
        else:
            kvalid = call["attn_mask"][:, 0, 0, :]
            gt = fx["kwargs"]["grid_size"][0]
            B = fx["x"].size(0)
            want = pm.reshape(B * gt, -1) if ci == 0 else pm.reshape(-1, gt)
            assert torch.equal(kvalid, want)
            mask = mask_predicate(N, Tq, Tk, False, -1, -1, k_valid=kvalid)

    o = sdpa_explicit(q, k, v, mask, scale)                                   # [N,Tq,H,hd]
    ref = call["out"].permute(0, 2, 1, 3)
    ref = torch.nan_to_num(ref, nan=0.0)  # belt and braces: torch>=2.5 already returns 0 for dead rows
    torch.testing.assert_close(o, ref, atol=2e-6, rtol=1e-5)


def test_reference_dead_rows_are_zero():
    """Fully masked query rows: the reference's SDPA (torch >= 2.5) returns zeros, and so does the oracle."""
    fx = load_golden("llm_hd16_causal_pad.pt")
    call = fx["sdpa_calls"][0]
    dead = ~call["attn_mask"].any(dim=-1)  # [B,H,T]
    assert dead.any()
    assert torch.all(call["out"][dead] == 0)


def test_reference_prefix_consistency_property_holds_for_oracle():
    """reference tests/transformers/nlp/attention_tests.py:111-131 restated on the core: with a causal mask the
    output at position t-1 of the full sequence equals the last position of the length-t prefix (atol 1e-3)."""
    torch.manual_seed(42)
    B, T, H, G, hd = 8, 16, 16, 2, 16
    q, k, v = torch.randn(B, T, H, hd), torch.randn(B, T, G, hd), torch.randn(B, T, G, hd)
    full = sdpa_explicit(q, k, v, mask_predicate(B, T, T, True, -1, 0), 4.0)
    for t in range(1, T):
        part = sdpa_explicit(q[:, :t], k[:, :t], v[:, :t], mask_predicate(B, t, t, True, -1, 0), 4.0)
        assert torch.allclose(part[:, -1], full[:, t - 1], atol=1e-3)


@pytest.mark.parametrize("Tq,Tk,causal,left,right", [
    (8, 8, True, -1, -1), (8, 8, True, 2, 0), (8, 8, False, 2, 1), (1, 9, True, 4, 0), (4, 9, True, 0, 0),
    (9, 4, True, 3, 0), (5, 5, False, 0, 0), (6, 6, False, -1, 2), (7, 7, True, 100, 0),
])
def test_predicate_against_python_loops(Tq, Tk, causal, left, right):
    """Pure-Python statement of SURVEY.md §8a-0 on small cases."""
    g = torch.Generator().manual_seed(Tq * 100 + Tk)
    qv = torch.rand(2, Tq, generator=g) > 0.3
    kv = torch.rand(2, Tk, generator=g) > 0.3
    m = mask_predicate(2, Tq, Tk, causal, left, right, qv, kv)
    off = Tk - Tq
    for n in range(2):
        for i in range(Tq):
            for j in range(Tk):
                ok = bool(qv[n, i]) and bool(kv[n, j])
                ok = ok and (not causal or j <= i + off)
                ok = ok and (left < 0 or j >= i + off - left)
                ok = ok and (right < 0 or j <= i + off + right)
                assert bool(m[n, i, j]) == ok


@pytest.mark.parametrize("name", prepare_golden_files())
def test_prepare_oracle_matches_reference_producers(name):
    """qk-norm + RoPE of the reference (apply_qk_norm, RoPE.forward) at one position == oracle.decode_prepare_explicit,
    and the oracle's cos/sin tables equal the reference's cache rows bit for bit."""
    from oracle import decode_prepare_explicit, rope_tables
    fx = load_golden(name)
    P, hd, G = fx["position"], fx["hd"], fx["G"]
    cos, sin = rope_tables(hd, fx["theta"], P + 1)
    assert torch.equal(cos[P], fx["cos_row"]) and torch.equal(sin[P], fx["sin_row"])
    B = fx["q_in"].size(0)
    lens = torch.full((B,), P + 1, dtype=torch.int32)
    kc = torch.zeros(B, P + 1, G, hd)
    vc = torch.zeros(B, P + 1, G, hd)
    v_in = torch.randn(B, G, hd)
    q_out, kc2, vc2 = decode_prepare_explicit(fx["q_in"], fx["k_in"], v_in, kc, vc, lens, cos, sin, fx["use_qk_norm"])
    assert torch.allclose(q_out, fx["q_out"], atol=2e-6, rtol=0)
    assert torch.allclose(kc2[:, P], fx["k_out"], atol=2e-6, rtol=0)
    assert torch.equal(vc2[:, P], v_in) and torch.count_nonzero(kc2[:, :P]) == 0


@pytest.mark.parametrize("name", prepare_seq_golden_files())
def test_prefill_prepare_oracle_matches_reference_producers(name):
    """apply_qk_norm + RoPE.forward of the reference over a whole sequence == oracle.prefill_prepare_explicit."""
    from oracle import prefill_prepare_explicit, rope_tables
    fx = load_golden(name)
    cos, sin = rope_tables(fx["hd"], fx["theta"], fx["T"])
    v = torch.randn_like(fx["k_in"])
    q, k, v2 = prefill_prepare_explicit(fx["q_in"], fx["k_in"], v, cos, sin, 0, fx["use_qk_norm"])
    assert torch.allclose(q, fx["q_out"], atol=2e-6, rtol=0) and torch.allclose(k, fx["k_out"], atol=2e-6, rtol=0)
    assert torch.equal(v2, v)


@pytest.mark.parametrize("name", cross_golden_files())
def test_oracle_matches_reference_cross_attention_call(name):
    """Image-gen cross-attention (reference cross_attention.py:73-101): key-padding mask == predicate with k_valid,
    SDPA output == oracle.sdpa_explicit with the reference's softmax_scale."""
    fx = load_golden(name)
    (call,) = fx["sdpa_calls"]
    q = call["q"].permute(0, 2, 1, 3).contiguous()
    k = call["k"].permute(0, 2, 1, 3).contiguous()
    v = call["v"].permute(0, 2, 1, 3).contiguous()
    N, Tq, H, hd = q.shape
    Tk = k.size(1)
    assert call["scale"] == fx["ctor"]["softmax_scale"] and not call["is_causal"]
    pm = fx["padding_mask"]
    mask = mask_predicate(N, Tq, Tk, False, -1, -1, k_valid=pm)
    if pm is None:
        assert call["attn_mask"] is None
    else:
        assert torch.equal(call["attn_mask"].expand(N, 1, Tq, Tk)[:, 0], mask)
    out = sdpa_explicit(q, k, v, mask, call["scale"])
    assert torch.allclose(out, call["out"].permute(0, 2, 1, 3), atol=2e-6, rtol=0)
