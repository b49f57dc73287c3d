"""GPU, BASELINE.json's full sizes: oracle on sampled units / rows (the full CPU oracle would take minutes), plus
size-independent properties (split invariance, shard == whole, two kernels agree)."""
import math

import pytest
import torch

from gpu_util import check_close, err_stats
from oracle import decode_explicit, sdpa_explicit
from vats_multimodal_lm_b200 import ops

pytestmark = pytest.mark.gpu


def _randn_bf16(shape, seed, normalize, device="cuda"):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
    if normalize:
        x = torch.nn.functional.normalize(x, dim=-1)
    return x.bfloat16()


def test_cfg2_decode_full_size_sampled_oracle_and_batch_invariance():
    """B=64, context 8192, window 4096, H=32, G=8, hd=128 (1.07 GB of K/V touched)."""
    B, S, H, G, hd, left = 64, 8192, 32, 8, 128, 4096
    kc = _randn_bf16((B, S, G, hd), 1, True)
    vc = _randn_bf16((B, S, G, hd), 2, False)
    q = _randn_bf16((B, H, hd), 3, True)
    lens = torch.full((B,), S, dtype=torch.int32, device="cuda")
    lens[5] = 4097
    lens[6] = 17
    scale = hd ** -0.5
    o = ops.gqa_swa_decode(q, kc, vc, lens, scale, left)
    torch.cuda.synchronize()
    assert torch.isfinite(o.float()).all()
    for b in (0, 5, 6, 63):
        ref = decode_explicit(q[b:b + 1].cpu(), kc[b:b + 1].cpu(), vc[b:b + 1].cpu(), lens[b:b + 1].cpu(), scale, left)
        check_close(o[b:b + 1], ref, f"cfg2 sequence {b}")
    # property: a sequence's result does not depend on its batch-mates (the shard == whole property of §8e)
    o_half = ops.gqa_swa_decode(q[32:], kc[32:], vc[32:], lens[32:], scale, left)
    # the split-K plan depends on the batch size, so the last bf16 bit may round differently
    assert torch.allclose(o_half.float(), o[32:].float(), atol=2e-3, rtol=2e-2)


def _sampled_rows_oracle(q, k, v, rows, scale, left):
    """fp32 oracle for selected (n, i) query rows of a causal + left-window prefill with Tq == Tk."""
    outs = []
    for (n, i) in rows:
        lo = max(0, i - left) if left >= 0 else 0
        qq = q[n:n + 1, i:i + 1].cpu()
        kk = k[n:n + 1, lo:i + 1].cpu()
        vv = v[n:n + 1, lo:i + 1].cpu()
        outs.append(sdpa_explicit(qq, kk, vv, None, scale)[0, 0])
    return torch.stack(outs)


def test_cfg5_long_prefill_one_sequence_sampled_oracle():
    """One sequence of the 32k / window 4096 / H=32 / G=8 / hd=128 config (the per-GPU unit at 8-way sharding)."""
    N, T, H, G, hd, left = 1, 32768, 32, 8, 128, 4096
    q = _randn_bf16((N, T, H, hd), 11, True)
    k = _randn_bf16((N, T, G, hd), 12, True)
    v = _randn_bf16((N, T, G, hd), 13, False)
    scale = hd ** -0.5
    o = ops.gqa_swa_prefill(q, k, v, None, None, scale, True, left, 0, ops.KERNEL_TCGEN05)
    torch.cuda.synchronize()
    assert torch.isfinite(o.float()).all()
    rows = [(0, i) for i in (0, 1, 127, 128, 4095, 4096, 4097, 4223, 8191, 20000, 32767)]
    ref = _sampled_rows_oracle(q, k, v, rows, scale, left)
    got = torch.stack([o[n, i] for (n, i) in rows])
    check_close(got, ref, "cfg5 sampled rows")
    # property: the two kernels agree on a slice (queries 30000.. against the same keys, bottom-right aligned)
    sl = slice(30000, 30512)
    a = ops.gqa_swa_prefill(q[:, sl], k[:, :30512], v[:, :30512], None, None, scale, True, left, 0, ops.KERNEL_SIMT)
    max_abs, rel = err_stats(o[:, sl], a.float().cpu())
    assert max_abs <= 2e-2 and rel <= 1e-2


def test_cfg3_cfg4_vit_shapes_sampled_oracle():
    for (N, T, H, G, hd, tag) in [(256, 196, 16, 8, 72, "cfg3"), (512, 196, 32, 8, 66, "cfg4a"),
                                  (12544, 8, 32, 8, 66, "cfg4b")]:
        q = _randn_bf16((N, T, H, hd), 21, True)
        k = _randn_bf16((N, T, G, hd), 22, True)
        v = _randn_bf16((N, T, G, hd), 23, False)
        scale = hd ** -0.5
        o = ops.gqa_swa_prefill(q, k, v, None, None, scale, False, -1, -1, 0)
        torch.cuda.synchronize()
        assert torch.isfinite(o.float()).all(), tag
        for n in (0, N // 2, N - 1):
            ref = sdpa_explicit(q[n:n + 1].cpu(), k[n:n + 1].cpu(), v[n:n + 1].cpu(), None, scale)
            check_close(o[n:n + 1], ref, f"{tag} sequence {n}")


def test_cfg1_llm_default_prefill():
    """LLM medium geometry (H=24, G=8, hd=60), batch 1, T in {32, 384, 4096}, causal, left 384."""
    H, G, hd, left = 24, 8, 60, 384
    scale = hd ** -0.5
    for T in (32, 384, 4096):
        q = _randn_bf16((1, T, H, hd), 31, True)
        k = _randn_bf16((1, T, G, hd), 32, True)
        v = _randn_bf16((1, T, G, hd), 33, False)
        o = ops.gqa_swa_prefill(q, k, v, None, None, scale, True, left, 0, 0)
        rows = [(0, i) for i in sorted({0, T // 3, T - 1, min(T - 1, 385), min(T - 1, 127)})]
        ref = _sampled_rows_oracle(q, k, v, rows, scale, left)
        got = torch.stack([o[n, i] for (n, i) in rows])
        check_close(got, ref, f"cfg1 T={T}")
