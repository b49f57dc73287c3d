"""GPU parity of prefill_tc64_kernel (the tile kernel with 64-key steps and double-buffered S, what AUTO takes for long
KV loops on TMA-addressable tensors) against the CPU oracle and the 128-key-step tile kernel: every mask form, one to
many KV steps (odd and even counts: the two S buffers, the prologue and the drain phase), a missing second head of a
pair, padding masks, the lazy-rescale path (peaky logits), bounded-logit softmax."""
import math

import pytest
import torch

from conftest import make_qkv
from gpu_util import check_close, oracle_prefill, run_prefill
from vats_multimodal_lm_b200 import _ffi, ops

pytestmark = pytest.mark.gpu
TC64 = ops.KERNEL_TC64

# (N, Tq, Tk, H, G, hd)
SHAPES = [
    (3, 64, 64, 2, 2, 32),        # one KV step
    (2, 128, 128, 2, 1, 128),     # two steps: the prologue issues everything
    (2, 130, 190, 4, 2, 64),      # three steps, Tq != Tk
    (2, 300, 300, 4, 2, 64),      # five steps, partial last block
    (1, 1000, 1000, 4, 2, 128),   # many steps, hd 128 (two swizzle regions)
    (1, 700, 700, 3, 1, 64),      # hpg = 3: the second pair has one head only
    (2, 77, 250, 4, 4, 72),       # MHA, hd 72
    (1, 2048, 2048, 8, 2, 16),    # hd 16, 32 steps
]
MASKS = [(True, -1, -1), (True, 100, 0), (True, 0, 0), (False, -1, -1), (False, 37, 11), (False, -1, 5)]


@pytest.mark.parametrize("causal,left,right", MASKS)
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_tc64_matches_oracle(shape, causal, left, right):
    N, Tq, Tk, H, G, hd = shape
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=sum(shape))
    scale = 1.0 / math.sqrt(hd)
    o = run_prefill(q, k, v, scale, causal, left, right, kernel=TC64)
    assert _ffi.last_kernel() == "prefill_tc64"
    check_close(o, oracle_prefill(q, k, v, scale, causal, left, right), f"{shape} causal={causal} window=({left},{right})")


def test_tc64_padding_masks():
    N, Tq, Tk, H, G, hd = 3, 300, 260, 4, 2, 64
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=11)
    g = torch.Generator().manual_seed(12)
    qv = torch.rand(N, Tq, generator=g) > 0.3
    kv = torch.rand(N, Tk, generator=g) > 0.3
    kv[0, :] = True
    kv[-1, 40:] = False
    scale = hd ** -0.5
    for causal, left, right, uq, uk in [(True, -1, 0, True, False), (False, -1, -1, False, True),
                                        (True, 64, 0, True, True), (False, 20, 20, True, True)]:
        o = run_prefill(q, k, v, scale, causal, left, right, qv if uq else None, kv if uk else None, kernel=TC64)
        ref = oracle_prefill(q, k, v, scale, causal, left, right, qv if uq else None, kv if uk else None)
        check_close(o, ref, f"pad uq={uq} uk={uk} causal={causal}")
        if uq:
            assert torch.all(o.cpu()[~qv] == 0)


@pytest.mark.parametrize("T", [128, 192, 200, 1000])
def test_tc64_rescale_path_and_drain_phase(T):
    """Un-normalised q / k with a large scale: the running maximum jumps by more than the lazy-rescale threshold between
    KV steps, in the last step too (which has to wait for the drain phase before it touches O)."""
    N, H, G, hd = 2, 4, 2, 64
    q, k, v = make_qkv(N, T, T, H, G, hd, seed=5 + T, unit_norm=False)
    k[:, -3:] *= 4.0           # the largest logits sit in the last step
    k[:, T // 2] *= 3.0
    for scale in (1.0 / 8.0, 4.0 / 8.0):
        for causal in (True, False):
            o = run_prefill(q, k, v, scale, causal, -1, 0 if causal else -1, kernel=TC64)
            check_close(o, oracle_prefill(q, k, v, scale, causal, -1, 0 if causal else -1), f"T={T} scale={scale} causal={causal}")


def test_tc64_bounded_logits_and_agreement_with_the_128_key_kernel():
    N, T, H, G, hd = 2, 1500, 8, 2, 128
    q, k, v = make_qkv(N, T, T, H, G, hd, seed=21)
    scale = hd ** -0.5
    ref = oracle_prefill(q, k, v, scale, True, 512, 0)
    o64 = run_prefill(q, k, v, scale, True, 512, 0, kernel=TC64)
    ob = run_prefill(q, k, v, scale, True, 512, 0, kernel=TC64, logit_bound=1.0)
    o128 = run_prefill(q, k, v, scale, True, 512, 0, kernel=ops.KERNEL_TCGEN05)
    assert _ffi.last_kernel() in ("prefill_tc", "prefill_tc64")
    check_close(o64, ref, "exact")
    check_close(ob, ref, "bounded")
    assert (o64.float() - o128.float()).abs().max().item() <= 2e-2


def test_tc64_many_items_ring_wraparound_and_auto_choice():
    """More work items than SMs (every ring / barrier phase flips many times); AUTO picks the kernel for long KV loops."""
    N, T, H, G, hd = 6, 2048, 8, 2, 64
    q, k, v = make_qkv(N, T, T, H, G, hd, seed=8)
    o = run_prefill(q, k, v, hd ** -0.5, True, -1, 0)
    assert _ffi.last_kernel() == "prefill_tc64"
    ref = oracle_prefill(q[:2], k[:2], v[:2], hd ** -0.5, True, -1, 0)
    check_close(o[:2], ref, "auto, first sequences")
    o128 = run_prefill(q, k, v, hd ** -0.5, True, -1, 0, kernel=ops.KERNEL_TCGEN05)
    assert (o.float() - o128.float()).abs().max().item() <= 2e-2
    # rows TMA cannot address keep the 128-key-step kernel (repack route)
    q, k, v = make_qkv(1, 1500, 1500, 4, 2, 60, seed=9)
    run_prefill(q, k, v, 60 ** -0.5, True, -1, 0)
    assert _ffi.last_kernel() in ("prefill_tc", "prefill_tc64")
