"""GPU parity of the resident-K/V kernel (prefill_mid_kernel: <= 256 keys, K/V of a KV group fetched once for all of
its heads and query blocks, single-pass softmax) against the CPU oracle — every mask form, both staging modes (TMA
and cp.async for rows TMA cannot address), head packing on and off, ragged shapes, and agreement with the other
kernels."""
import math

import pytest
import torch

from conftest import make_qkv
from gpu_util import check_close, oracle_prefill, run_prefill
from vats_multimodal_lm_b200 import _ffi, ops

pytestmark = pytest.mark.gpu
MID = ops.KERNEL_MID

# (N, Tq, Tk, H, G, hd)
SHAPES = [
    (3, 196, 196, 4, 2, 72),     # ViT-2D geometry (cfg3): hd 72 -> two swizzle regions, 2 heads packed per tile
    (2, 196, 196, 8, 2, 66),     # ViT-3D spatial (cfg4a), DENSE hd 66: cp.async staging, 4 heads packed
    (2, 200, 200, 6, 2, 60),     # hpg = 3: no packing, one tile per head; dense hd 60 (8-byte rows)
    (1, 128, 128, 2, 1, 128),    # exactly one tile, hd 128
    (2, 256, 256, 4, 2, 128),    # the largest K/V that fits: 256 keys x hd 128
    (1, 40, 40, 2, 2, 16),       # hd 16, a single partial tile
    (5, 33, 33, 8, 1, 64),       # MQA, 8 heads packed, n_pad 48
    (2, 77, 250, 4, 4, 64),      # Tq != Tk (bottom-right alignment), MHA
    (3, 300, 100, 4, 2, 32),     # more queries than keys: fully masked rows under a causal mask
    (2, 1, 200, 8, 2, 64),       # a single query token
    (1, 1000, 64, 16, 1, 64),    # cross-attention-like: many queries, few keys, 16 heads packed
    (2, 130, 17, 64, 1, 32),     # 64 heads per group: packs of 32
]
MASKS = [(True, -1, -1), (True, 100, 0), (True, 0, 0), (False, -1, -1), (False, 37, 11), (False, -1, 5)]


@pytest.mark.parametrize("causal,left,right", MASKS)
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_mid_matches_oracle(shape, causal, left, right):
    N, Tq, Tk, H, G, hd = shape
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=sum(shape))
    scale = 1.0 / math.sqrt(hd)
    o = run_prefill(q, k, v, scale, causal, left, right, kernel=MID)
    assert _ffi.last_kernel() == "prefill_mid"
    ref = oracle_prefill(q, k, v, scale, causal, left, right)
    check_close(o, ref, f"{shape} causal={causal} window=({left},{right})")


@pytest.mark.parametrize("shape", [(3, 196, 196, 4, 2, 72), (2, 196, 196, 8, 2, 66), (4, 150, 90, 6, 2, 64)],
                         ids=lambda s: "x".join(map(str, s)))
def test_mid_padding_masks(shape):
    N, Tq, Tk, H, G, hd = shape
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=11)
    g = torch.Generator().manual_seed(12)
    qv = torch.rand(N, Tq, generator=g) > 0.3
    kv = torch.rand(N, Tk, generator=g) > 0.3
    kv[0, :] = True
    kv[-1, 40:] = False
    scale = 1.0 / math.sqrt(hd)
    for causal, left, right, uq, uk in [(True, -1, 0, True, False), (False, -1, -1, False, True),
                                        (True, 64, 0, True, True), (False, 20, 20, True, True)]:
        o = run_prefill(q, k, v, scale, causal, left, right, qv if uq else None, kv if uk else None, kernel=MID)
        ref = oracle_prefill(q, k, v, scale, causal, left, right, qv if uq else None, kv if uk else None)
        check_close(o, ref, f"{shape} pad uq={uq} uk={uk} causal={causal}")
        if uq:
            assert torch.all(o.cpu()[~qv] == 0)


def test_mid_module_layout_hd66_uses_tma_and_matches_dense():
    """The drop-in modules hand hd 66 over with the head stride padded to 72 (TMA-addressable); a dense tensor goes
    through the cp.async staging — same arithmetic, bit-identical results."""
    N, T, H, G, hd = 4, 196, 8, 2, 66
    q, k, v = make_qkv(N, T, T, H, G, hd, seed=3)

    def padded(x):
        buf = torch.zeros(*x.shape[:-1], 72, dtype=x.dtype, device="cuda")
        buf[..., :hd] = x.cuda()
        return buf[..., :hd]

    scale = hd ** -0.5
    o_dense = ops.gqa_swa_prefill(q.cuda(), k.cuda(), v.cuda(), None, None, scale, False, -1, -1, MID)
    o_pad = ops.gqa_swa_prefill(padded(q), padded(k), padded(v), None, None, scale, False, -1, -1, MID)
    torch.cuda.synchronize()
    assert torch.equal(o_dense, o_pad)
    check_close(o_dense, oracle_prefill(q, k, v, scale, False, -1, -1), "hd66")


def test_mid_cpasync_staging_without_workspace_matches_repack_route():
    """A C-ABI caller that passes no scratch gets the in-kernel cp.async staging (kLdg instantiation); the op layer hands
    over scratch and gets repack + TMA.  Both must agree bit for bit (same arithmetic) and with the oracle."""
    for (N, T, H, G, hd) in [(3, 196, 8, 2, 66), (2, 150, 6, 2, 60), (2, 77, 4, 4, 66)]:
        q, k, v = make_qkv(N, T, T, H, G, hd, seed=hd + T)
        dq, dk, dv = q.cuda(), k.cuda(), v.cuda()
        scale = hd ** -0.5
        o_ws = ops.gqa_swa_prefill(dq, dk, dv, None, None, scale, True, 50, 0, MID)
        o_raw = torch.empty_like(o_ws)
        s3 = lambda t: tuple(t.stride()[:3])
        _ffi.prefill(dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), o_raw.data_ptr(), None, None, N, T, T, H, G, hd,
                     s3(dq), s3(dk), s3(dv), s3(o_raw), scale, True, 50, 0, torch.cuda.current_stream().cuda_stream, MID)
        torch.cuda.synchronize()
        assert _ffi.last_kernel() == "prefill_mid"
        assert torch.equal(o_ws, o_raw)
        check_close(o_raw, oracle_prefill(q, k, v, scale, True, 50, 0), f"cp.async staging hd={hd}")


def test_mid_unnormalised_peaky_logits():
    N, T, H, G, hd = 2, 200, 4, 2, 64
    q, k, v = make_qkv(N, T, T, H, G, hd, seed=5, unit_norm=False)
    for scale in (1.0 / 8.0, 4.0 / 8.0):
        o = run_prefill(q, k, v, scale, True, -1, 0, kernel=MID)
        check_close(o, oracle_prefill(q, k, v, scale, True, -1, 0), f"unnormalised scale={scale}")


def test_mid_many_items_exercise_ring_wraparound():
    """More (sequence, KV group) items than SMs x ring depth: every mbarrier phase flips several times."""
    N, T, H, G, hd = 700, 50, 4, 2, 32
    q, k, v = make_qkv(N, T, T, H, G, hd, seed=8)
    o = run_prefill(q, k, v, hd ** -0.5, False, -1, -1, kernel=MID)
    check_close(o, oracle_prefill(q, k, v, hd ** -0.5, False, -1, -1), "many items")
    N, T, H, G, hd = 600, 196, 2, 2, 72     # one tile per item and odd tile counts per CTA
    q, k, v = make_qkv(N, T, T, H, G, hd, seed=9)
    o = run_prefill(q, k, v, hd ** -0.5, False, -1, -1, kernel=MID)
    check_close(o, oracle_prefill(q, k, v, hd ** -0.5, False, -1, -1), "many items, hd 72")


def test_auto_picks_mid_for_vit_shapes_and_kernels_agree():
    N, T, H, G, hd = 10, 196, 16, 8, 72     # 80 (sequence, KV group) items: enough for AUTO to pick the kernel
    q, k, v = make_qkv(N, T, T, H, G, hd, seed=21)
    scale = hd ** -0.5
    o_auto = run_prefill(q, k, v, scale, False, -1, -1)
    assert _ffi.last_kernel() == "prefill_mid"
    o_tc = run_prefill(q, k, v, scale, False, -1, -1, kernel=ops.KERNEL_TCGEN05)
    assert _ffi.last_kernel() == "prefill_tc"
    ref = oracle_prefill(q, k, v, scale, False, -1, -1)
    check_close(o_auto, ref, "auto")
    check_close(o_tc, ref, "tc")
    assert (o_auto.float() - o_tc.float()).abs().max().item() <= 2e-2
    # beyond 256 keys the tile kernel takes over
    q, k, v = make_qkv(1, 300, 300, 4, 2, 64, seed=22)
    run_prefill(q, k, v, 0.125, True, -1, 0)
    assert _ffi.last_kernel() == "prefill_tc"


def test_mid_strided_views_of_fused_qkv():
    N, T, H, G, hd = 2, 170, 6, 2, 64
    g = torch.Generator().manual_seed(31)
    qkv = torch.randn(N, T, (H + 2 * G) * hd, generator=g).bfloat16()
    dq, dk, dv = torch.split(qkv.cuda(), [H * hd, G * hd, G * hd], dim=-1)
    o = ops.gqa_swa_prefill(dq.view(N, T, H, hd), dk.view(N, T, G, hd), dv.view(N, T, G, hd), None, None, 0.05, True,
                            90, 0, MID)
    q, k, v = torch.split(qkv, [H * hd, G * hd, G * hd], dim=-1)
    ref = oracle_prefill(q.view(N, T, H, hd), k.view(N, T, G, hd), v.view(N, T, G, hd), 0.05, True, 90, 0)
    check_close(o, ref, "fused qkv views")


@pytest.mark.parametrize("kernel", [MID, ops.KERNEL_TCGEN05], ids=["mid", "tcgen05"])
@pytest.mark.parametrize("shape", [(3, 196, 196, 4, 2, 72), (2, 196, 196, 8, 2, 66), (2, 256, 256, 4, 2, 128),
                                   (2, 77, 250, 4, 4, 64), (1, 700, 700, 4, 2, 64)], ids=lambda s: "x".join(map(str, s)))
def test_bounded_logits_skip_the_row_maximum_and_give_the_same_softmax(shape, kernel):
    """logit_bound = 1.0 is what the modules pass behind qk-norm (unit-norm q, k): the kernels use the bound instead of
    the row maximum (softmax is shift-invariant).  Same results as the exact-maximum path, same oracle tolerance, for
    every mask form incl. fully masked rows and key padding."""
    N, Tq, Tk, H, G, hd = shape
    if kernel == MID and Tk > 256:
        pytest.skip("resident-K/V kernel: <= 256 keys")
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=sum(shape))      # unit-norm q, k
    g = torch.Generator().manual_seed(5)
    kv = torch.rand(N, Tk, generator=g) > 0.3
    kv[:, 0] = True
    scale = 1.0 / math.sqrt(hd)
    for causal, left, right, use_kv in [(False, -1, -1, False), (True, 60, 0, False), (True, -1, 0, True), (False, 9, 30, True)]:
        kvm = kv if use_kv else None
        o_b = run_prefill(q, k, v, scale, causal, left, right, None, kvm, kernel=kernel, logit_bound=1.0)
        o_e = run_prefill(q, k, v, scale, causal, left, right, None, kvm, kernel=kernel)
        ref = oracle_prefill(q, k, v, scale, causal, left, right, None, kvm)
        check_close(o_b, ref, f"bounded {shape} causal={causal} ({left},{right}) kv={use_kv}")
        assert (o_b.float() - o_e.float()).abs().max().item() <= 2e-2   # both within the tolerance of the same oracle


# (N, Tq, Tk, H, G, hd): sequences whose last tile holds only a few (token, head) rows
RESIDUAL_SHAPES = [
    (5, 196, 196, 4, 2, 72),      # cfg3 geometry: 2 heads packed, 3 full tiles + 4 tokens x 2 heads
    (3, 196, 196, 8, 2, 66),      # cfg4a geometry, DENSE hd 66: repack + TMA, 6 full tiles + 4 tokens x 4 heads
    (3, 196, 196, 8, 2, 72),      # the same with TMA-addressable rows
    (7, 33, 33, 8, 1, 64),        # 8 heads packed: 2 full tiles + 1 token; n_pad 48 (odd 16-key block)
    (4, 130, 130, 2, 2, 48),      # no packing: one full tile + 2 tokens, hd 48 in a 64-wide tile
    (4, 140, 200, 4, 4, 60),      # Tq != Tk, dense hd 60 (repack), 12 residual rows
    (2, 72, 256, 4, 2, 80),       # 256 keys, hd 80, 64-token tiles: 1 full tile + 8 tokens x 2 heads
    (300, 68, 100, 4, 2, 64),     # many items per CTA
]


@pytest.mark.parametrize("bound", [0.0, 1.0], ids=["exact", "bounded"])
@pytest.mark.parametrize("shape", RESIDUAL_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_mid_residual_query_tokens(shape, bound):
    """Unmasked geometries whose last tile is almost empty, exact and bounded-logit softmax: every row — the residual
    ones included — against the oracle and against the tile kernel."""
    N, Tq, Tk, H, G, hd = shape
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=sum(shape))
    scale = 1.0 / math.sqrt(hd)
    o = run_prefill(q, k, v, scale, False, -1, -1, kernel=MID, logit_bound=bound)
    assert _ffi.last_kernel() == "prefill_mid"
    ref = oracle_prefill(q, k, v, scale, False, -1, -1)
    check_close(o, ref, f"{shape} bound={bound}")
    check_close(o[:, -4:], ref[:, -4:], f"{shape} last tokens")
    o_tc = run_prefill(q, k, v, scale, False, -1, -1, kernel=ops.KERNEL_TCGEN05)
    assert (o.float() - o_tc.float()).abs().max().item() <= 2e-2


def test_repack_chunk_kernel_strided_sources():
    """The repack route (rows TMA cannot address) with dense and strided hd-66 / hd-60 sources: views of a fused QKV
    projection, token-strided views, 4-byte aligned bases."""
    N, T, H, G, hd = 3, 150, 8, 2, 66
    g = torch.Generator().manual_seed(77)
    fused = torch.nn.functional.normalize(torch.randn(N, T, (H + 2 * G) * hd + 2, generator=g), dim=-1).bfloat16().cuda()
    body = fused[..., 2:]   # 4-byte aligned base
    q = body[..., : H * hd].unflatten(-1, (H, hd))
    k = body[..., H * hd: (H + G) * hd].unflatten(-1, (G, hd))
    v = body[..., (H + G) * hd:].unflatten(-1, (G, hd))
    scale = hd ** -0.5
    for kernel in (MID, ops.KERNEL_TCGEN05):
        o = ops.gqa_swa_prefill(q, k, v, None, None, scale, False, -1, -1, kernel)
        torch.cuda.synchronize()
        assert _ffi.last_launch_count() == 2   # repack + the TMA-fed kernel
        check_close(o, oracle_prefill(q.cpu(), k.cpu(), v.cpu(), scale, False, -1, -1), f"fused view kernel={kernel}")
