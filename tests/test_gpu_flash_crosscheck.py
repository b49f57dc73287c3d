"""GPU, test-only third-party pin: sliding-window, Tq != Tk alignment and single-query decode semantics against
flash-attn (the library whose `window_size=(left, right)` the reference's intended call names,
src/optimized_attention.py:628-635; SURVEY §8c allows it as an optional cross-check on the GPU box).  The reference can
execute none of these behaviours itself, so this is the only external implementation they can be pinned to.  The
package never imports flash_attn; the whole file is skipped where it is missing or has no kernel for this GPU.
"""
import pytest
import torch

from conftest import make_qkv
from gpu_util import check_close, oracle_prefill
from vats_multimodal_lm_b200 import ops
from oracle import mask_predicate

pytestmark = pytest.mark.gpu

try:
    from flash_attn import flash_attn_func
except Exception:  # pragma: no cover - depends on the box
    flash_attn_func = None


def _flash(q, k, v, scale, causal, left, right):
    if flash_attn_func is None:
        pytest.skip("flash_attn is not importable here")
    try:
        return flash_attn_func(q, k, v, dropout_p=0.0, softmax_scale=scale, causal=causal, window_size=(left, right))
    except RuntimeError as e:  # e.g. no kernel image for sm_100 in this build
        pytest.skip(f"flash_attn cannot run on this GPU: {str(e)[:120]}")


CASES = [  # N, Tq, Tk, H, G, hd, causal, left, right
    (2, 300, 300, 4, 2, 64, True, 100, 0),        # LLM: causal + left window
    (2, 300, 300, 4, 2, 64, True, 0, 0),          # window of exactly the diagonal
    (1, 1000, 1000, 8, 2, 128, True, 384, 0),     # band across several KV tiles
    (2, 77, 333, 4, 4, 64, True, 50, 0),          # Tq != Tk: bottom-right alignment (chunked prefill)
    (2, 77, 333, 4, 2, 64, True, -1, 0),
    (3, 1, 500, 8, 2, 128, True, 128, 0),         # single-query decode step
    (3, 1, 500, 8, 2, 128, False, 128, -1),       # ... for which causal does not matter
    (2, 196, 196, 4, 2, 64, False, 30, 30),       # ViT: symmetric window, non-causal
    (2, 196, 196, 4, 2, 64, False, 5, 60),
    (2, 160, 200, 4, 2, 32, False, 17, 3),        # Tq != Tk non-causal window
    (2, 64, 64, 4, 2, 64, False, -1, -1),
]


@pytest.mark.parametrize("N,Tq,Tk,H,G,hd,causal,left,right", CASES)
def test_outputs_match_flash_attn(N, Tq, Tk, H, G, hd, causal, left, right):
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=Tq * 7 + Tk)
    scale = hd ** -0.5
    ref_fa = _flash(q.cuda(), k.cuda(), v.cuda(), scale, causal, left, right)
    o = ops.gqa_swa_prefill(q.cuda(), k.cuda(), v.cuda(), None, None, scale, causal, left, 0 if causal else right, 0)
    torch.cuda.synchronize()
    # two bf16-output kernels: each is within the stated tolerance of the fp32 oracle, and of each other
    check_close(o, ref_fa.float().cpu(), "vs flash_attn")
    ref = oracle_prefill(q, k, v, scale, causal, left, 0 if causal else right)
    check_close(o, ref, "vs oracle")
    check_close(ref_fa, ref, "flash_attn vs oracle")


@pytest.mark.parametrize("Tq,Tk,causal,left,right", [
    (128, 128, True, 20, 0), (128, 128, True, 0, 0), (128, 128, False, 7, 9), (128, 128, False, -1, 3),
    (128, 128, False, 3, -1), (50, 128, True, 10, 0), (50, 128, False, 10, 4), (1, 128, True, 31, 0),
    (1, 97, True, -1, 0), (128, 40, True, 5, 0), (128, 40, False, 2, 2), (100, 100, True, 200, 0),
])
def test_mask_pattern_is_bit_exact_with_flash_attn(Tq, Tk, causal, left, right):
    """q = 0 makes every allowed key equally likely and V = identity reads the probabilities out: row i of the output
    is 1/count on the allowed keys and 0 elsewhere.  The allowed SET must be identical in flash-attn, in our kernel,
    in the device predicate (ops.attn_mask) and in the oracle — this pins inclusivity and the bottom-right alignment."""
    hd = 128
    q = torch.zeros(1, Tq, 2, hd, dtype=torch.bfloat16, device="cuda")
    k = torch.randn(1, Tk, 1, hd, device="cuda").bfloat16()
    v = torch.zeros(1, Tk, 1, hd, dtype=torch.bfloat16, device="cuda")
    v[0, torch.arange(Tk), 0, torch.arange(Tk)] = 1.0
    fa = _flash(q, k, v, 1.0, causal, left, right)
    o = ops.gqa_swa_prefill(q, k, v, None, None, 1.0, causal, left, right, 0)
    torch.cuda.synchronize()
    pat_fa = (fa[0, :, 0, :Tk] > 0).cpu()
    pat_us = (o[0, :, 0, :Tk] > 0).cpu()
    pred = mask_predicate(1, Tq, Tk, causal, left, right)[0]
    dev_pred = ops.attn_mask(None, None, 1, Tq, Tk, causal, left, right).bool().cpu()[0]
    assert torch.equal(pat_us, pred), "kernel vs oracle predicate"
    assert torch.equal(dev_pred, pred), "device predicate vs oracle predicate"
    # flash-attn leaves rows without any allowed key undefined-ish (zeros); compare the rows that have one
    live = pred.any(-1)
    assert torch.equal(pat_fa[live], pred[live]), "flash_attn vs oracle predicate"
