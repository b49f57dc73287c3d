"""CPU: the drop-ins at the reference's REAL call sites (needs the reference tree; skipped where it is absent, e.g.
on the GPU box).

  * `integration.patch_reference()` swaps the classes with no reference source change; the
    unmodified `AutoregressiveTextTransformer`, `ImageEncoderTransformer` and `VideoTransformer` are then BUILT from the
    drop-ins and load a state-dict produced by the unpatched reference model.
  * the unmodified `AutoregressiveTokenGenerator._generate` (src/transformers/nlp/inference/generate.py:35-243) is run
    end to end through the drop-in `Attention` / `KVCache`: prefill with `use_cache=True` and a [B,T] mask, the uncached
    re-forward of step 0, then T=1 cached steps with a [B,1] mask — the H-head `KVCache` both call sites build
    (model.py:148-154, generate.py:27-33) must be accepted and every cached step must go through
    `decode_prepare` + `gqa_swa_decode`.  The ops are replaced by oracle stand-ins here (host logic only; the same
    sequence runs on the real kernels in tests/test_gpu_generate.py).
"""
import os
import sys

import pytest
import torch

import vats_multimodal_lm_b200 as vl
from vats_multimodal_lm_b200 import integration, ops
from vats_multimodal_lm_b200.modules import llm as L
from oracle import (decode_explicit, decode_prepare_explicit, mask_predicate, prefill_prepare_explicit,
                    sdpa_explicit)

REF = os.environ.get("VATS_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present")


@pytest.fixture()
def reference_on_path(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                       # utils/setup_logger.py creates ./logs at import
    monkeypatch.setenv("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    yield
    integration.unpatch_reference()
    sys.path.remove(REF)


class OpLog:
    def __init__(self):
        self.calls = []

    def count(self, name):
        return sum(1 for c in self.calls if c[0] == name)


@pytest.fixture()
def oracle_ops(monkeypatch):
    """Stand-ins for the four ops the LLM drop-in calls, computed by the CPU oracle (bf16-rounded like the kernels)."""
    log = OpLog()

    def prefill(q, k, v, q_valid, k_valid, scale, causal, left, right, kernel=0, logit_bound=0.0):
        log.calls.append(("prefill", tuple(q.shape), tuple(k.shape)))
        m = mask_predicate(q.size(0), q.size(1), k.size(1), causal, left, right, q_valid, k_valid)
        return sdpa_explicit(q, k, v, m, scale).to(torch.bfloat16)

    def decode(q, kc, vc, seq_lens, scale, left):
        log.calls.append(("decode", tuple(q.shape), tuple(kc.shape), seq_lens.tolist()))
        return decode_explicit(q, kc, vc, seq_lens, scale, left).to(torch.bfloat16)

    def decode_prepare(q, k, v, kc, vc, seq_lens, cos, sin, qk_norm, eps):
        log.calls.append(("decode_prepare", tuple(q.shape), tuple(kc.shape), seq_lens.tolist()))
        qo, kc2, vc2 = decode_prepare_explicit(q, k, v, kc, vc, seq_lens, cos, sin, qk_norm, eps)
        for b in range(q.size(0)):
            p = int(seq_lens[b]) - 1
            if p >= 0:
                kc[b, p] = kc2[b, p].to(kc.dtype)
                vc[b, p] = vc2[b, p].to(vc.dtype)
        return qo.to(torch.bfloat16)

    def prefill_prepare_views(q, k, v, cos, sin, pos0, qk_norm, eps=1e-6):
        log.calls.append(("prefill_prepare", tuple(q.shape), pos0))
        return tuple(t.to(torch.bfloat16) for t in prefill_prepare_explicit(q, k, v, cos, sin, pos0, qk_norm, eps))

    monkeypatch.setattr(ops, "gqa_swa_prefill", prefill)
    monkeypatch.setattr(ops, "gqa_swa_decode", decode)
    monkeypatch.setattr(ops, "decode_prepare", decode_prepare)
    monkeypatch.setattr(ops, "prefill_prepare_views", prefill_prepare_views)
    monkeypatch.setattr(L, "_on_gpu", lambda x: True)
    return log


def _llm_args():
    from configs.transformers.nlp.model_args.model_args_xsmall import ModelArgs
    return ModelArgs(d_model=64, num_heads=4, query_groups=2, d_ffn=256, num_layers=2, vocab_size=97, max_seq_len=40,
                     left_window=6, softmax_scale=0.25, dropout=0.0)


def test_patch_builds_reference_models_from_dropins_and_loads_reference_state(reference_on_path):
    torch.manual_seed(0)
    # --- unpatched reference models -> state dicts
    from src.transformers.nlp.model import AutoregressiveTextTransformer
    import src.optimized_attention as ref_attn
    ref_llm = AutoregressiveTextTransformer(_llm_args())
    assert type(ref_llm.layers[0].attn_block) is ref_attn.AttentionBlock and not isinstance(
        ref_llm.layers[0].attn_block, vl.AttentionBlock)
    from src.transformers.vision.vit_2d.model import ImageEncoderTransformer
    from configs.transformers.vision.vit_2d.model_args.model_args_xsmall import ModelArgs as A2
    a2 = A2(target_size=64, d_model=96, num_heads=4, query_groups=2, d_ffn=384, num_layers=2)
    ref_v2 = ImageEncoderTransformer(a2)
    from src.transformers.vision.vit_3d.model import VideoTransformer
    from configs.transformers.vision.vit_3d.model_args.model_args_xsmall import ModelArgs as A3
    a3 = A3(target_size=(32, 32), max_frames=4, d_model=96, num_heads=4, query_groups=2, d_ffn=384, num_layers=2,
            num_classes=10)
    ref_v3 = VideoTransformer(a3)

    # --- swap the classes (model files already imported: names are rebound there too)
    done = integration.patch_reference(strict=True)
    assert "src.transformers.nlp.model" in done and "KVCache" in done["src.transformers.nlp.model"]
    new_llm = AutoregressiveTextTransformer(_llm_args())
    assert isinstance(new_llm.layers[0].attn_block, vl.AttentionBlock)
    assert isinstance(new_llm.kv_cache, vl.KVCache) and new_llm.kv_cache.num_heads == 4   # H heads, as model.py builds it
    new_llm.load_state_dict(ref_llm.state_dict(), strict=True)
    new_v2 = ImageEncoderTransformer(a2)
    assert any(isinstance(m, vl.SpatialAttention) for m in new_v2.modules())
    new_v2.load_state_dict(ref_v2.state_dict(), strict=True)
    new_v3 = VideoTransformer(a3)
    assert any(isinstance(m, vl.SpatioTemporalAttention) for m in new_v3.modules())
    new_v3.load_state_dict(ref_v3.state_dict(), strict=True)

    integration.unpatch_reference()
    import src.transformers.nlp.model as ref_model_mod
    assert ref_model_mod.KVCache is ref_attn.KVCache and ref_attn.KVCache is not vl.KVCache


@pytest.mark.parametrize("with_eos", [False, True])
def test_unmodified_generate_runs_on_dropins_and_reaches_the_decode_ops(reference_on_path, oracle_ops, with_eos):
    integration.patch_reference(strict=True)
    from src.transformers.nlp.inference.generate import AutoregressiveTokenGenerator
    torch.manual_seed(1)
    args = _llm_args()
    gen = AutoregressiveTokenGenerator(args)
    assert isinstance(gen.kv_cache, vl.KVCache) and isinstance(gen.model.kv_cache, vl.KVCache)
    B, T, new = 3, 5, 7
    ids = torch.randint(1, args.vocab_size, (B, T))
    ids[1, 3:] = 0                                        # padded prompt (pad_token_id = 0)

    steps = []
    orig_forward = gen.model.forward

    def recording_forward(input_ids, padding_mask=None, use_cache=False):
        out = orig_forward(input_ids=input_ids, padding_mask=padding_mask, use_cache=use_cache)
        steps.append((tuple(input_ids.shape), bool(use_cache), padding_mask.clone(), out[0][:, -1].clone()))
        return out

    gen.model.forward = recording_forward
    eos = None
    if with_eos:   # make sequence 0 finish after its second generated token
        probe = gen._generate(ids, new, temperature=0.0, pad_token_id=0, eos_token_id=None, use_cache=True)
        eos = int(probe[0, T + 1])
        steps.clear()
        oracle_ops.calls.clear()
    out = gen._generate(ids, new, temperature=0.0, pad_token_id=0, eos_token_id=eos, use_cache=True)
    assert out.shape == (B, T + new)

    # call pattern of generate.py:96-137
    assert steps[0][:2] == ((B, T), True) and steps[1][:2] == ((B, T), False)
    cached = steps[2:]
    assert len(cached) == new - 1 and all(s[0] == (B, 1) and s[1] for s in cached)
    n_layers = args.num_layers
    assert oracle_ops.count("decode") == n_layers * (new - 1) == oracle_ops.count("decode_prepare")
    dec = [c for c in oracle_ops.calls if c[0] == "decode"]
    assert dec[0][2] == (B, args.max_seq_len, args.query_groups, args.d_model // args.num_heads)   # G heads stored
    assert [c[3][2] for c in dec[::n_layers]] == list(range(T + 1, T + new))                        # positions advance
    if with_eos:
        assert any(c[3][0] == 0 for c in dec), "a finished sequence must reach the decode op as seq_len 0"
        assert all(c[3][0] > 0 for c in oracle_ops.calls if c[0] == "decode_prepare"), "k/v are appended regardless"

    # cached steps == an uncached forward of the same tokens (causal => position p predicts from the prefix)
    gen.model.forward = orig_forward
    mask = torch.ones(B, T + new - 1, dtype=torch.bool)
    mask[:, :T] = ids != 0
    with torch.no_grad():
        full, _, _ = gen.model(input_ids=out[:, :-1], padding_mask=mask, use_cache=False)
    for s, (_, _, pm, logits) in enumerate(cached):
        live = pm[:, 0]
        ref = full[:, T + s]
        assert live.any()
        torch.testing.assert_close(logits[live], ref[live], atol=3e-2, rtol=3e-2)

    # a second call starts from an empty cache although generate.py resets ITS OWN cache object, not the model's
    assert gen.model.kv_cache.current_seq_len is None
    out2 = gen._generate(ids, new, temperature=0.0, pad_token_id=0, eos_token_id=eos, use_cache=True)
    assert torch.equal(out, out2)


def test_reference_attention_test_sequence_with_h_head_cache(oracle_ops):
    """tests/transformers/nlp/attention_tests.py:73-102 of the reference: initialize(B), two cached multi-token calls."""
    torch.manual_seed(2)
    a = vl.Attention(64, 4, 2, 10000.0, 0.25)
    cache = vl.KVCache(max_batch_size=4, max_seq_len=32, num_heads=4, head_dim=16, num_layers=1)   # num_heads = H
    cache.initialize(batch_size=2)
    x = torch.randn(2, 6, 64)
    o1, c1 = a(x, -1, -1, True, None, cache, 0, True)
    assert o1.shape == x.shape and c1["k"].shape == (2, 6, 2, 16)
    x2 = torch.randn(2, 4, 64)
    o2, _ = a(x2, -1, -1, True, None, cache, 0, True)
    k_total, v_total = cache.get(layer_idx=0, seq_len=cache.current_seq_len)
    assert cache.current_seq_len == 10 and k_total.shape == (2, 10, 2, 16) and cache.kv_heads == 2
    # the chunked result equals one pass over the ten tokens
    full, _ = a(torch.cat([x, x2], 1), -1, -1, True)
    torch.testing.assert_close(torch.cat([o1, o2], 1), full, atol=2e-2, rtol=2e-2)
    with pytest.raises(ValueError, match="overflow"):
        a(torch.randn(2, 30, 64), -1, -1, True, None, cache, 0, True)
    # head stride is padded to 8 elements for head dims TMA cannot address (60 -> 64)
    c60 = vl.KVCache(2, 8, 24, 60, 1)
    c60.initialize(1)
    c60.bind(8)
    assert c60.cache[0]["k"].shape == (1, 8, 8, 60) and c60.cache[0]["k"].stride() == (8 * 8 * 64, 8 * 64, 64, 1)


def test_rope_tables_stay_fp32_after_module_cast():
    a = vl.Attention(64, 4, 2, 10000.0, 0.25)
    ref_cos, ref_sin = a.rope.get_cos_sin_cache(12)
    ref_cos, ref_sin = ref_cos.clone(), ref_sin.clone()
    a = a.to(torch.bfloat16)
    cos, sin = a.rope.get_cos_sin_cache(12)
    assert cos.dtype == torch.float32 and sin.dtype == torch.float32
    torch.testing.assert_close(cos, ref_cos, atol=1e-6, rtol=0)
    sd = a.state_dict()
    sd["rope.cos_cache"] = ref_cos.to(torch.bfloat16)
    sd["rope.sin_cache"] = ref_sin.to(torch.bfloat16)
    a.load_state_dict(sd)
    assert a.rope.cos_cache.dtype == torch.float32
