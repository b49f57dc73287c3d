"""CPU: the call-level drop-in for `F.scaled_dot_product_attention` (vats_multimodal_lm_b200/sdpa_adapter.py) that
serves the reference's remaining attention call sites: mask decomposition, and — with the op
replaced by the oracle (host logic only; the kernels run the same fixtures in tests/test_gpu_modules.py) — agreement
with what the unmodified reference modules computed (tests/golden/site_*.pt, produced by oracle/gen_golden.py)."""
import os
import sys

import pytest
import torch

from conftest import GOLDEN, load_golden
from oracle import mask_predicate, sdpa_explicit
from vats_multimodal_lm_b200 import integration, ops, sdpa_adapter

SITE_FILES = sorted(f for f in os.listdir(GOLDEN) if f.startswith("site_"))
REF = os.environ.get("VATS_REFERENCE", "/root/reference")


@pytest.fixture()
def oracle_op(monkeypatch):
    def prefill(q, k, v, q_valid, k_valid, scale, causal, left, right, kernel=0, logit_bound=0.0):
        m = mask_predicate(q.size(0), q.size(1), k.size(1), causal, left, right, q_valid, k_valid)
        return sdpa_explicit(q, k, v, m, scale).to(torch.bfloat16)
    monkeypatch.setattr(ops, "gqa_swa_prefill", prefill)


def test_decompose_mask_forms():
    B, T = 3, 7
    g = torch.Generator().manual_seed(0)
    kv = torch.rand(B, T, generator=g) > 0.3
    qv = torch.rand(B, T, generator=g) > 0.3
    kv[:, 0] = qv[:, 0] = True
    tril = torch.ones(T, T, dtype=torch.bool).tril()
    # key padding as the reference expands it (text encoder :286, image-gen :243): a stride-0 view
    q_, k_, c_ = sdpa_adapter.decompose_mask(kv[:, None, None, :].expand(B, 4, T, T), B, T, T)
    assert q_ is None and torch.equal(k_, kv) and c_ is False
    # query-row padding (video-gen :186-189)
    q_, k_, c_ = sdpa_adapter.decompose_mask(qv[:, None, :, None].expand(B, 1, T, T), B, T, T)
    assert torch.equal(q_, qv) and k_ is None and c_ is False
    # materialised products
    for m, causal in [((qv[:, :, None] & kv[:, None, :]), False), ((kv[:, None, :] & tril[None]), True),
                      ((qv[:, :, None] & tril[None]), True), (tril[None].expand(B, T, T).clone(), True)]:
        q_, k_, c_ = sdpa_adapter.decompose_mask(m[:, None].expand(B, 4, T, T).contiguous(), B, T, T)
        rebuilt = torch.ones(B, T, T, dtype=torch.bool)
        if q_ is not None:
            rebuilt &= q_[:, :, None]
        if k_ is not None:
            rebuilt &= k_[:, None, :]
        if c_:
            rebuilt &= tril[None]
        assert c_ is causal and torch.equal(rebuilt, m)
    with pytest.raises(NotImplementedError):          # an arbitrary mask has no kernel path and no fallback
        sdpa_adapter.decompose_mask((torch.rand(B, 1, T, T, generator=g) > 0.5), B, T, T)
    with pytest.raises(NotImplementedError):
        sdpa_adapter.decompose_mask(torch.zeros(B, 1, T, T), B, T, T)     # additive float mask


@pytest.mark.parametrize("fname", SITE_FILES)
def test_drop_in_reproduces_the_reference_sdpa_calls(fname, oracle_op, monkeypatch):
    monkeypatch.setattr(sdpa_adapter, "_to_kernel_layout", lambda t, pad=True: t.to(torch.bfloat16))
    fx = load_golden(fname)
    assert fx["sdpa_calls"]
    for c in fx["sdpa_calls"]:
        out = sdpa_adapter.sdpa_drop_in(c["q"], c["k"], c["v"], attn_mask=c["attn_mask"], is_causal=c["is_causal"],
                                        scale=c["scale"])
        ref = torch.nan_to_num(c["out"], nan=0.0)      # rows without any allowed key: zeros here, NaN in older torch
        assert out.shape == ref.shape and out.dtype == c["q"].dtype
        torch.testing.assert_close(out, ref, atol=2e-2, rtol=0)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present")
def test_patched_reference_modules_route_through_the_op(oracle_op, monkeypatch, tmp_path):
    """INTEGRATION.md §1 applied to the live reference: after `patch_reference()` the four remaining attention modules
    run their own code and reach the op for the attention arithmetic; outputs match the unpatched run (fixtures)."""
    monkeypatch.setattr(sdpa_adapter, "_to_kernel_layout", lambda t, pad=True: t.to(torch.bfloat16))
    monkeypatch.chdir(tmp_path)
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    calls = []
    real = ops.gqa_swa_prefill
    monkeypatch.setattr(ops, "gqa_swa_prefill", lambda *a, **k: (calls.append(a[0].shape), real(*a, **k))[1])
    try:
        done = integration.patch_reference(strict=True)
        assert all(site in done for site in sdpa_adapter.SDPA_CALL_SITES)
        import importlib
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import gen_golden as gg
        from oracle import reference as refmod
        results = {}

        def site_case(name, kind, seed, build, run):
            m = refmod.site(kind)
            torch.manual_seed(seed)
            mod = build(m).eval()
            with torch.no_grad():
                results[name] = run(mod)
        monkeypatch.setattr(gg, "site_case", site_case)
        gg.site_cases()
        assert len(calls) >= len(results)
        for name, out in results.items():
            ref = torch.nan_to_num(load_golden(name + ".pt")["out"], nan=0.0)
            torch.testing.assert_close(torch.nan_to_num(out, nan=0.0), ref, atol=3e-2, rtol=0), name
    finally:
        integration.unpatch_reference()
        sys.path.remove(REF)
