"""GPU parity tests proper: CUDA kernels (through torch.ops -> ctypes -> C-ABI) vs the CPU oracle on identical
seeded inputs, both prefill kernels, decode, and the bit-exact mask predicate."""
import math

import pytest
import torch

from conftest import make_qkv
from gpu_util import check_close, err_stats, oracle_prefill, run_prefill
from oracle import decode_explicit, mask_predicate
from vats_multimodal_lm_b200 import _ffi, ops

pytestmark = pytest.mark.gpu

TC, SIMT, AUTO = ops.KERNEL_TCGEN05, ops.KERNEL_SIMT, ops.KERNEL_AUTO

# (N, Tq, Tk, H, G, hd)
SHAPES = [
    (2, 300, 300, 4, 2, 64),     # 2.3 q-blocks, ragged tail
    (1, 128, 128, 2, 1, 128),    # exactly one tile, hd 128 (two swizzle regions)
    (1, 129, 257, 4, 4, 128),    # Tq != Tk (bottom-right aligned), H == G (second M-tile idle)
    (3, 196, 196, 4, 2, 72),     # ViT-2D geometry: hd 72 -> 80, split map zero-fills the padding
    (2, 196, 196, 8, 2, 66),     # ViT-3D spatial geometry: hd 66, merged map + Q fix-up
    (2, 200, 200, 6, 2, 60),     # LLM medium geometry: hd 60, H/G = 3 (odd: one idle tile per group)
    (1, 40, 40, 2, 2, 16),       # hd 16
    (1, 700, 700, 3, 1, 48),     # MQA-like G = 1, H/G = 3, hd 48
    (2, 384, 1024, 8, 2, 128),   # chunked prefill against a longer cache
]
MASKS = [
    # causal, left, right
    (True, -1, -1), (True, 100, 0), (True, 0, 0), (False, -1, -1), (False, 37, 11), (True, 5000, 0),
]


@pytest.mark.parametrize("kernel", [TC, SIMT], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("causal,left,right", MASKS)
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_prefill_matches_oracle(shape, causal, left, right, kernel):
    N, Tq, Tk, H, G, hd = shape
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=hash(shape) % 1000)
    scale = 1.0 / math.sqrt(hd)
    o = run_prefill(q, k, v, scale, causal, left, right, kernel=kernel)
    ref = oracle_prefill(q, k, v, scale, causal, left, right)
    check_close(o, ref, f"{shape} causal={causal} window=({left},{right}) kernel={kernel}")


@pytest.mark.parametrize("kernel", [TC, SIMT], ids=["tcgen05", "simt"])
def test_prefill_unnormalised_inputs_and_large_scale(kernel):
    """N(0,1) q,k (no qk-norm) with the xsmall LLM's softmax_scale = 4.0: peaky softmax, exercises the online
    rescaling path."""
    N, Tq, Tk, H, G, hd = 2, 520, 520, 4, 2, 64
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=5, unit_norm=False)
    for scale in (1.0 / 8.0, 4.0 / 8.0):
        o = run_prefill(q, k, v, scale, True, -1, 0, kernel=kernel)
        ref = oracle_prefill(q, k, v, scale, True, -1, 0)
        check_close(o, ref, f"unnormalised scale={scale} kernel={kernel}")


@pytest.mark.parametrize("kernel", [TC, SIMT], ids=["tcgen05", "simt"])
def test_prefill_increasing_logits_force_rescale(kernel):
    """Keys whose logits grow along the sequence: the running maximum rises in every KV tile."""
    N, T, H, G, hd = 1, 640, 2, 1, 64
    g = torch.Generator().manual_seed(3)
    q = torch.ones(N, T, H, hd) * 0.5
    k = (torch.arange(T, dtype=torch.float32)[None, :, None, None] / T * 8.0).expand(N, T, G, hd).contiguous()
    v = torch.randn(N, T, G, hd, generator=g)
    q, k, v = q.bfloat16(), k.bfloat16(), v.bfloat16()
    o = run_prefill(q, k, v, 1.0, False, -1, -1, kernel=kernel)
    ref = oracle_prefill(q, k, v, 1.0, False, -1, -1)
    check_close(o, ref, f"increasing logits kernel={kernel}")


@pytest.mark.parametrize("kernel", [TC, SIMT], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("shape", [(3, 260, 260, 4, 2, 64), (2, 196, 196, 4, 1, 66)], ids=lambda s: "x".join(map(str, s)))
def test_prefill_padding_masks(shape, kernel):
    N, Tq, Tk, H, G, hd = shape
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=11)
    g = torch.Generator().manual_seed(12)
    qv = torch.rand(N, Tq, generator=g) > 0.3
    kv = torch.rand(N, Tk, generator=g) > 0.3
    kv[0, :] = True
    kv[-1, 130:] = False           # a whole KV tile without valid keys
    scale = 1.0 / math.sqrt(hd)
    for causal, left, right, uq, uk in [(True, -1, 0, True, False), (False, -1, -1, False, True),
                                        (True, 64, 0, True, True), (False, 20, 20, True, True)]:
        o = run_prefill(q, k, v, scale, causal, left, right, qv if uq else None, kv if uk else None, kernel=kernel)
        ref = oracle_prefill(q, k, v, scale, causal, left, right, qv if uq else None, kv if uk else None)
        check_close(o, ref, f"{shape} pad uq={uq} uk={uk} causal={causal} kernel={kernel}")
        if uq:  # masked query rows are exactly zero, never NaN
            assert torch.all(o.cpu()[~qv] == 0)


@pytest.mark.parametrize("kernel", [TC, SIMT], ids=["tcgen05", "simt"])
def test_prefill_fully_masked_rows_are_zero(kernel):
    """Tq > Tk with causal: the first Tq - Tk query rows see no key at all (off < 0)."""
    N, Tq, Tk, H, G, hd = 1, 300, 140, 2, 2, 64
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=21)
    o = run_prefill(q, k, v, 0.125, True, -1, 0, kernel=kernel)
    ref = oracle_prefill(q, k, v, 0.125, True, -1, 0)
    check_close(o, ref, f"Tq>Tk kernel={kernel}")
    assert torch.all(o.cpu()[:, : Tq - Tk] == 0)


@pytest.mark.parametrize("kernel", [TC, SIMT], ids=["tcgen05", "simt"])
def test_prefill_strided_views_of_fused_qkv(kernel):
    """q, k, v as views into one [N, T, (H+2G)*hd] projection output, as the modules produce them."""
    N, T, H, G, hd = 2, 270, 6, 2, 64
    g = torch.Generator().manual_seed(31)
    qkv = torch.randn(N, T, (H + 2 * G) * hd, generator=g).bfloat16()
    dqkv = qkv.cuda()
    dq, dk, dv = torch.split(dqkv, [H * hd, G * hd, G * hd], dim=-1)
    o = ops.gqa_swa_prefill(dq.view(N, T, H, hd), dk.view(N, T, G, hd), dv.view(N, T, G, hd), None, None, 0.05, True,
                            90, 0, kernel)
    q, k, v = torch.split(qkv, [H * hd, G * hd, G * hd], dim=-1)
    ref = oracle_prefill(q.view(N, T, H, hd), k.view(N, T, G, hd), v.view(N, T, G, hd), 0.05, True, 90, 0)
    check_close(o, ref, f"fused qkv views kernel={kernel}")


def test_prefill_head_major_layout_tcgen05():
    """[N, H, T, hd] storage viewed as [N, T, H, hd] (head stride > token stride): split tensor map."""
    N, T, H, G, hd = 2, 260, 4, 2, 128
    q, k, v = make_qkv(N, T, T, H, G, hd, seed=41)
    dq = q.permute(0, 2, 1, 3).contiguous().cuda().permute(0, 2, 1, 3)
    dk = k.permute(0, 2, 1, 3).contiguous().cuda().permute(0, 2, 1, 3)
    dv = v.permute(0, 2, 1, 3).contiguous().cuda().permute(0, 2, 1, 3)
    o = ops.gqa_swa_prefill(dq, dk, dv, None, None, 0.09, True, -1, 0, TC)
    ref = oracle_prefill(q, k, v, 0.09, True, -1, 0)
    check_close(o, ref, "head-major layout")


def test_two_kernels_agree_and_auto_plan():
    N, Tq, Tk, H, G, hd = 2, 333, 333, 8, 2, 128
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=51)
    a = run_prefill(q, k, v, 0.09, True, 77, 0, kernel=TC)
    b = run_prefill(q, k, v, 0.09, True, 77, 0, kernel=SIMT)
    max_abs, rel = err_stats(a, b.float().cpu())
    assert max_abs <= 2e-2 and rel <= 1e-2
    s = lambda t: t.stride()[:3]
    dq, dk, dv = q.cuda(), k.cuda(), v.cuda()
    assert _ffi.prefill_plan(N, Tq, Tk, H, G, hd, s(dq), s(dk), s(dv), s(dq), dq.data_ptr(), dk.data_ptr(),
                             dv.data_ptr()) == TC   # (N*G = 4 items: too few for the resident-K/V kernel)
    # ViT-3D temporal pass (8 tokens) and hd = 6 go to the CUDA-core kernel
    assert _ffi.prefill_plan(64, 8, 8, 32, 8, 66, (8 * 32 * 66, 32 * 66, 66), (8 * 8 * 66, 8 * 66, 66),
                             (8 * 8 * 66, 8 * 66, 66), (8 * 32 * 66, 32 * 66, 66), dq.data_ptr(), dk.data_ptr(),
                             dv.data_ptr()) == SIMT


def test_vit3d_temporal_and_tiny_head_dim_simt():
    for (N, T, H, G, hd) in [(500, 8, 32, 8, 66), (7, 5, 124, 2, 6), (3, 33, 4, 4, 130)]:
        q, k, v = make_qkv(N, T, T, H, G, hd, seed=61)
        g = torch.Generator().manual_seed(62)
        kv = torch.rand(N, T, generator=g) > 0.3
        kv[:, 0] = True
        o = run_prefill(q, k, v, 1 / math.sqrt(hd), False, -1, -1, None, kv, kernel=AUTO)
        ref = oracle_prefill(q, k, v, 1 / math.sqrt(hd), False, -1, -1, None, kv)
        check_close(o, ref, f"simt {(N, T, H, G, hd)}")


# (N, Tq, Tk, H, G, hd): the one-CTA-per-sequence kernel (Tk <= 32): bulk (dense, odd hd/2) and cp.async staging, every
# keys-per-pass / heads-per-thread instantiation, Tq != Tk, more sequences than persistent CTAs
# (the last three: more query tokens than fit one stage -> several work items per sequence; image-gen cross-attention
#  geometry: thousands of image tokens against 16 text tokens)
SHORT_SHAPES = [
    (3, 700, 8, 32, 8, 66), (2, 3000, 16, 8, 8, 16), (2, 1000, 31, 6, 3, 34),
    (700, 8, 8, 32, 8, 66), (5, 8, 8, 8, 2, 64), (4, 16, 16, 6, 3, 34), (3, 32, 32, 4, 4, 20), (9, 5, 8, 12, 4, 26),
    (6, 8, 3, 16, 2, 66), (2, 30, 17, 3, 3, 10), (11, 1, 1, 8, 8, 66), (3, 12, 12, 10, 5, 18), (2, 7, 7, 6, 6, 128),
]


@pytest.mark.parametrize("causal,left,right", [(False, -1, -1), (True, -1, 0), (True, 3, 0), (False, 2, 1)])
@pytest.mark.parametrize("shape", SHORT_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_short_sequence_kernel_matches_oracle(shape, causal, left, right):
    N, Tq, Tk, H, G, hd = shape
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=71)
    g = torch.Generator().manual_seed(72)
    kv = torch.rand(N, Tk, generator=g) > 0.25
    qv = torch.rand(N, Tq, generator=g) > 0.2
    scale = 1 / math.sqrt(hd)
    for (q_valid, k_valid) in [(None, None), (qv, kv)]:
        ref = oracle_prefill(q, k, v, scale, causal, left, right, q_valid, k_valid)
        o = run_prefill(q, k, v, scale, causal, left, right, q_valid, k_valid, kernel=SIMT)
        check_close(o, ref, f"short dense {shape}")
        # the same values seen through a padded, non-dense layout (head stride rounded up to 8, as the modules do)
        hp = (hd + 7) // 8 * 8 + 8
        qp, kp, vp = (torch.zeros(*t.shape[:-1], hp, dtype=t.dtype) for t in (q, k, v))
        qp[..., :hd], kp[..., :hd], vp[..., :hd] = q, k, v
        dq, dk, dv = (t.cuda()[..., :hd] for t in (qp, kp, vp))
        o2 = ops.gqa_swa_prefill(dq, dk, dv, None if q_valid is None else q_valid.cuda(),
                                 None if k_valid is None else k_valid.cuda(), scale, causal, left, right, SIMT)
        torch.cuda.synchronize()
        check_close(o2, ref, f"short strided {shape}")
        assert torch.equal(o.cpu(), o2.cpu()), "the two staging paths must give identical bits"



def test_empty_inputs():
    dev = "cuda"
    z = lambda *s: torch.zeros(*s, dtype=torch.bfloat16, device=dev)
    assert ops.gqa_swa_prefill(z(2, 0, 4, 64), z(2, 5, 2, 64), z(2, 5, 2, 64), None, None, 1.0, True, -1, -1).shape == (2, 0, 4, 64)
    o = ops.gqa_swa_prefill(z(2, 3, 4, 64) + 1, z(2, 0, 2, 64), z(2, 0, 2, 64), None, None, 1.0, False, -1, -1)
    assert o.shape == (2, 3, 4, 64) and torch.all(o == 0)   # no keys: zero rows


MASK_CASES = [
    (2, 1, 1, True, -1, -1), (2, 7, 7, True, 0, 0), (2, 64, 64, True, 3, 0), (1, 130, 130, False, 5, 9),
    (3, 1, 300, True, 128, 0), (2, 300, 1, True, -1, 0), (2, 100, 257, True, 100, 0), (2, 257, 100, True, 256, 0),
    (1, 129, 129, False, -1, 0), (1, 129, 129, False, 0, -1), (2, 50, 50, True, 49, 0), (2, 50, 50, True, 50, 0),
    (2, 50, 50, True, 51, 0), (2, 50, 50, False, 1 << 30, 1 << 30),
]


@pytest.mark.parametrize("N,Tq,Tk,causal,left,right", MASK_CASES)
def test_mask_predicate_bit_exact(N, Tq, Tk, causal, left, right):
    g = torch.Generator().manual_seed(Tq * 7 + Tk)
    qv = torch.rand(N, Tq, generator=g) > 0.25
    kv = torch.rand(N, Tk, generator=g) > 0.25
    for uq, uk in [(False, False), (True, False), (False, True), (True, True)]:
        got = ops.attn_mask(qv.cuda() if uq else None, kv.cuda() if uk else None, N, Tq, Tk, causal, left, right)
        want = mask_predicate(N, Tq, Tk, causal, left, right, qv if uq else None, kv if uk else None)
        assert torch.equal(got.cpu().bool(), want)


def _mk_cache(B, S, G, hd, H, seed):
    g = torch.Generator().manual_seed(seed)
    kc = torch.nn.functional.normalize(torch.randn(B, S, G, hd, generator=g), dim=-1).bfloat16()
    vc = torch.randn(B, S, G, hd, generator=g).bfloat16()
    q = torch.nn.functional.normalize(torch.randn(B, H, hd, generator=g), dim=-1).bfloat16()
    return q, kc, vc


DECODE_CASES = [
    # B, S_max, H, G, hd, left, seq_lens
    (4, 600, 32, 8, 128, 256, [600, 1, 257, 300]),
    (3, 1500, 8, 2, 128, -1, [1500, 777, 2]),
    (5, 300, 24, 8, 60, 100, [300, 101, 100, 99, 0]),       # hd 60 (64-bit loads), H/G = 3, an empty sequence
    (2, 400, 16, 2, 16, 0, [400, 9]),                       # left = 0: the query sees only itself; H/G = 8
    (3, 520, 4, 4, 64, 4096, [520, 519, 1]),                # H == G
    (2, 260, 6, 1, 48, 50, [260, 30]),                      # MQA, H/G = 6 -> two head batches
    (2, 300, 8, 4, 66, 128, [300, 150]),                    # hd 66: 32-bit loads
    (1, 9000, 32, 8, 128, 4096, [8192]),                    # BASELINE decode geometry, one sequence
]


@pytest.mark.parametrize("B,S,H,G,hd,left,lens", DECODE_CASES)
def test_decode_matches_oracle(B, S, H, G, hd, left, lens):
    q, kc, vc = _mk_cache(B, S, G, hd, H, seed=S + hd)
    sl = torch.tensor(lens, dtype=torch.int32)
    scale = 1.0 / math.sqrt(hd)
    o = ops.gqa_swa_decode(q.cuda(), kc.cuda(), vc.cuda(), sl.cuda(), scale, left)
    torch.cuda.synchronize()
    ref = decode_explicit(q, kc, vc, sl, scale, left)
    check_close(o, ref, f"decode {(B, S, H, G, hd, left)}")
    for b, L in enumerate(lens):
        if L == 0:
            assert torch.all(o[b] == 0)


def test_decode_unnormalised_peaky():
    B, S, H, G, hd = 2, 1100, 8, 2, 128
    g = torch.Generator().manual_seed(9)
    q = torch.randn(B, H, hd, generator=g).bfloat16()
    kc = torch.randn(B, S, G, hd, generator=g).bfloat16()
    vc = torch.randn(B, S, G, hd, generator=g).bfloat16()
    sl = torch.tensor([1100, 640], dtype=torch.int32)
    o = ops.gqa_swa_decode(q.cuda(), kc.cuda(), vc.cuda(), sl.cuda(), 0.3, 800)
    ref = decode_explicit(q, kc, vc, sl, 0.3, 800)
    check_close(o, ref, "decode peaky")


def test_decode_equals_prefill_last_row():
    """A decode step is the last row of a bottom-right aligned prefill over the same cache."""
    B, S, H, G, hd, left = 2, 700, 8, 2, 128, 300
    q, kc, vc = _mk_cache(B, S, G, hd, H, seed=77)
    sl = torch.full((B,), S, dtype=torch.int32)
    od = ops.gqa_swa_decode(q.cuda(), kc.cuda(), vc.cuda(), sl.cuda(), 0.09, left)
    op = ops.gqa_swa_prefill(q[:, None].cuda(), kc.cuda(), vc.cuda(), None, None, 0.09, True, left, 0, SIMT)
    max_abs, rel = err_stats(od, op[:, 0].float().cpu())
    assert max_abs <= 1e-2 and rel <= 5e-3


def test_launch_count_reported():
    # TMA-addressable cache -> mma.sync kernel, splits merged by the last CTA: one launch
    q, kc, vc = _mk_cache(2, 600, 2, 64, 4, seed=1)
    ops.gqa_swa_decode(q.cuda(), kc.cuda(), vc.cuda(), torch.tensor([600, 600], dtype=torch.int32).cuda(), 0.1, -1)
    assert _ffi.last_launch_count() == 1
    # hd 60 is not TMA-addressable -> CUDA-core kernel: split kernel + combine, or one launch for a single split
    q, kc, vc = _mk_cache(2, 600, 2, 60, 4, seed=1)
    ops.gqa_swa_decode(q.cuda(), kc.cuda(), vc.cuda(), torch.tensor([600, 600], dtype=torch.int32).cuda(), 0.1, -1)
    assert _ffi.last_launch_count() == 2
    ops.gqa_swa_decode(q.cuda(), kc.cuda(), vc.cuda(), torch.tensor([600, 600], dtype=torch.int32).cuda(), 0.1, 100)
    assert _ffi.last_launch_count() == 1


def test_decode_workspace_is_reusable_across_calls():
    """The split counters in the workspace are left at zero: repeated calls (and a different window) stay correct."""
    B, S, H, G, hd = 3, 2000, 16, 4, 128
    q, kc, vc = _mk_cache(B, S, G, hd, H, seed=5)
    dq, dk, dv = q.cuda(), kc.cuda(), vc.cuda()
    for lens, left in [([2000, 1500, 900], -1), ([2000, 1500, 900], -1), ([1999, 3, 1000], 700), ([2000, 2000, 0], 64)]:
        sl = torch.tensor(lens, dtype=torch.int32)
        o = ops.gqa_swa_decode(dq, dk, dv, sl.cuda(), 0.09, left)
        check_close(o, decode_explicit(q, kc, vc, sl, 0.09, left), f"reuse {lens} {left}")


def test_decode_stale_rows_past_sequence_end_do_not_leak():
    """Cache rows beyond seq_len may hold anything (here NaN): they must not reach the output."""
    B, S, H, G, hd = 2, 300, 8, 2, 128
    q, kc, vc = _mk_cache(B, S, G, hd, H, seed=6)
    lens = torch.tensor([137, 290], dtype=torch.int32)
    kd, vd = kc.clone(), vc.clone()
    for b, L in enumerate(lens.tolist()):
        kd[b, L:] = float("nan")
        vd[b, L:] = float("nan")
    o = ops.gqa_swa_decode(q.cuda(), kd.cuda(), vd.cuda(), lens.cuda(), 0.09, -1)
    check_close(o, decode_explicit(q, kc, vc, lens, 0.09, -1), "stale rows")


# ---- fused decode pre-core step (qk-norm + RoPE + cache append), SURVEY §8f rank 1
from conftest import load_golden, prepare_golden_files  # noqa: E402
from oracle import decode_prepare_explicit, rope_tables  # noqa: E402


@pytest.mark.parametrize("name", prepare_golden_files())
def test_decode_prepare_matches_reference_fixture(name):
    """The kernel against what the unmodified reference produced (apply_qk_norm + RoPE at position P)."""
    fx = load_golden(name)
    P, hd, G, H = fx["position"], fx["hd"], fx["G"], fx["H"]
    B = fx["q_in"].size(0)
    cos, sin = rope_tables(hd, fx["theta"], P + 8)
    kc = torch.zeros(B, P + 8, G, hd, dtype=torch.bfloat16, device="cuda")
    vc = torch.zeros_like(kc)
    lens = torch.full((B,), P + 1, dtype=torch.int32, device="cuda")
    v_in = torch.randn(B, G, hd, generator=torch.Generator().manual_seed(5))
    q_out = ops.decode_prepare(fx["q_in"].cuda(), fx["k_in"].cuda(), v_in.cuda(), kc, vc, lens, cos.cuda(), sin.cuda(),
                               fx["use_qk_norm"], 1e-6)
    torch.cuda.synchronize()
    # one bf16 rounding of fp32 values: relative error <= 2^-8
    assert torch.allclose(q_out.float().cpu(), fx["q_out"], atol=4e-3 * fx["q_out"].abs().max().item(), rtol=4e-3)
    assert torch.equal(q_out.cpu(), fx["q_out"].bfloat16()) or \
        (q_out.float().cpu() - fx["q_out"]).abs().max() <= 2 ** -8 * fx["q_out"].abs().max()
    assert (kc[:, P].float().cpu() - fx["k_out"]).abs().max() <= 2 ** -8 * fx["k_out"].abs().max()
    assert torch.equal(vc[:, P].cpu(), v_in.bfloat16())
    assert torch.count_nonzero(kc[:, :P]) == 0 and torch.count_nonzero(kc[:, P + 1:]) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("B,H,G,hd,S", [(5, 32, 8, 128, 600), (3, 24, 8, 60, 97), (4, 6, 6, 66, 40), (2, 8, 1, 256, 33)])
def test_decode_prepare_matches_oracle_and_feeds_decode(B, H, G, hd, S, dtype):
    g = torch.Generator().manual_seed(31)
    q = torch.randn(B, H, hd, generator=g).to(dtype)
    k = torch.randn(B, G, hd, generator=g).to(dtype)
    v = torch.randn(B, G, hd, generator=g).to(dtype)
    kc = torch.nn.functional.normalize(torch.randn(B, S, G, hd, generator=g), dim=-1).bfloat16()
    vc = torch.randn(B, S, G, hd, generator=g).bfloat16()
    lens = torch.randint(1, S + 1, (B,), generator=g).int()
    lens[0] = S          # last slot
    lens[-1] = 0         # empty sequence: nothing is written
    cos, sin = rope_tables(hd, 10000.0, S)
    dk, dv = kc.cuda().clone(), vc.cuda().clone()
    q_out = ops.decode_prepare(q.cuda(), k.cuda(), v.cuda(), dk, dv, lens.cuda(), cos.cuda(), sin.cuda(), True, 1e-6)
    torch.cuda.synchronize()
    q_ref, k_ref, v_ref = decode_prepare_explicit(q, k, v, kc, vc, lens, cos, sin, True)
    live = lens > 0
    assert (q_out.float().cpu()[live] - q_ref[live]).abs().max() <= 2 ** -8 + 1e-6   # |q| <= 1 after the norm
    assert (dk.float().cpu() - k_ref).abs().max() <= 2 ** -8 + 1e-6
    assert torch.equal(dv.cpu(), v_ref.bfloat16())
    # the decode kernel consumes what the prepare kernel wrote
    o = ops.gqa_swa_decode(q_out, dk, dv, lens.cuda(), hd ** -0.5, 64)
    torch.cuda.synchronize()
    ref = decode_explicit(q_out.cpu(), dk.cpu(), dv.cpu(), lens, hd ** -0.5, 64)
    check_close(o[live.cuda()], ref[live], "decode after prepare")
    # no rotation / no norm variants
    dk2, dv2 = kc.cuda().clone(), vc.cuda().clone()
    q2 = ops.decode_prepare(q.cuda(), k.cuda(), v.cuda(), dk2, dv2, lens.cuda(), None, None, False, 1e-6)
    q2_ref, k2_ref, _ = decode_prepare_explicit(q, k, v, kc, vc, lens, None, None, False)
    assert torch.equal(q2.cpu()[live], q2_ref.bfloat16()[live]) and torch.equal(dk2.cpu(), k2_ref.bfloat16())


def test_decode_random_geometries_back_to_back():
    """Many decode calls of different geometry on one stream: workspace reuse across shapes and split counts, empty
    sequences, every TMA head dim — exercises the consumer / flush-warp hand-off across item boundaries."""
    g = torch.Generator().manual_seed(123)
    for it in range(24):
        B = int(torch.randint(1, 70, (1,), generator=g))
        G = int(torch.randint(1, 5, (1,), generator=g))
        H = G * int(torch.randint(1, 9, (1,), generator=g))
        hd = [16, 32, 64, 128][it % 4]
        S = int(torch.randint(1, 2500, (1,), generator=g))
        left = [-1, 0, 17, 300, 4096][it % 5]
        kc = torch.nn.functional.normalize(torch.randn(B, S, G, hd, generator=g), dim=-1).bfloat16()
        vc = torch.randn(B, S, G, hd, generator=g).bfloat16()
        q = torch.nn.functional.normalize(torch.randn(B, H, hd, generator=g), dim=-1).bfloat16()
        lens = torch.randint(0, S + 1, (B,), generator=g).int()
        o = ops.gqa_swa_decode(q.cuda(), kc.cuda(), vc.cuda(), lens.cuda(), hd ** -0.5, left)
        ref = decode_explicit(q, kc, vc, lens, hd ** -0.5, left)
        check_close(o, ref, f"decode {(B, H, G, hd, S, left)}")


@pytest.mark.parametrize("N,Tq,Tk", [(3, 1, 300), (5, 4, 700), (2, 15, 129), (4, 2, 64)])
def test_few_query_tokens_against_many_keys(N, Tq, Tk):
    """Chunked-prefill tails / speculative-token verification: Tq << Tk, bottom-right aligned causal window.  AUTO picks
    the tensor-core kernel (Tk >= 32) and both kernels agree with the oracle."""
    H, G, hd = 8, 2, 64
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=81)
    ref = oracle_prefill(q, k, v, hd ** -0.5, True, 100, 0)
    for kern in (AUTO, TC, SIMT):
        o = run_prefill(q, k, v, hd ** -0.5, True, 100, 0, kernel=kern)
        check_close(o, ref, f"few-query {(N, Tq, Tk)} kernel {kern}")
    s = lambda t: tuple(t.stride()[:3])
    dq, dk, dv = q.cuda(), k.cuda(), v.cuda()
    assert _ffi.prefill_plan(N, Tq, Tk, H, G, hd, s(dq), s(dk), s(dv), s(dq), dq.data_ptr(), dk.data_ptr(),
                             dv.data_ptr()) == TC   # (N*G = 4 items: too few for the resident-K/V kernel)


# ---- fused prefill pre-core producers (qk-norm + RoPE + bf16 + TMA-addressable layout), SURVEY §8f rank 1
from conftest import prepare_seq_golden_files  # noqa: E402
from oracle import prefill_prepare_explicit  # noqa: E402


@pytest.mark.parametrize("name", prepare_seq_golden_files())
def test_prefill_prepare_matches_reference_fixture(name):
    fx = load_golden(name)
    cos, sin = rope_tables(fx["hd"], fx["theta"], fx["T"])
    v = torch.randn_like(fx["k_in"])
    q, k, v2 = ops.prefill_prepare_views(fx["q_in"].cuda(), fx["k_in"].cuda(), v.cuda(), cos.cuda(), sin.cuda(), 0,
                                         fx["use_qk_norm"])
    torch.cuda.synchronize()
    assert (q.float().cpu() - fx["q_out"]).abs().max() <= 2 ** -8 * fx["q_out"].abs().max()
    assert (k.float().cpu() - fx["k_out"]).abs().max() <= 2 ** -8 * fx["k_out"].abs().max()
    assert torch.equal(v2.cpu(), v.bfloat16())
    assert q.stride(2) % 8 == 0 and q.stride(3) == 1      # TMA-addressable head stride (hd 60 -> 64)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("N,T,H,G,hd,pos0", [(2, 300, 8, 2, 128, 0), (3, 77, 6, 2, 60, 11), (5, 8, 8, 4, 66, 0),
                                             (1, 1, 4, 4, 16, 500)])
def test_prefill_prepare_matches_oracle_and_feeds_attention(N, T, H, G, hd, pos0, dtype):
    g = torch.Generator().manual_seed(41)
    # strided views of a fused projection output, as the modules hand them over
    qkv = torch.randn(N, T, (H + 2 * G) * hd, generator=g).to(dtype)
    q, k, v = (t.view(N, T, -1, hd) for t in torch.split(qkv, [H * hd, G * hd, G * hd], dim=-1))
    cos, sin = rope_tables(hd, 10000.0, pos0 + T)
    dqkv = qkv.cuda()
    cq, ck, cv = (t.view(N, T, -1, hd) for t in torch.split(dqkv, [H * hd, G * hd, G * hd], dim=-1))
    dq, dk, dv = ops.prefill_prepare_views(cq, ck, cv, cos.cuda(), sin.cuda(), pos0, True)
    torch.cuda.synchronize()
    q_ref, k_ref, v_ref = prefill_prepare_explicit(q, k, v, cos, sin, pos0, True)
    assert (dq.float().cpu() - q_ref).abs().max() <= 2 ** -8 + 1e-6
    assert (dk.float().cpu() - k_ref).abs().max() <= 2 ** -8 + 1e-6
    assert torch.equal(dv.cpu(), v_ref.bfloat16())
    o = ops.gqa_swa_prefill(dq, dk, dv, None, None, hd ** -0.5, True, 64, 0, AUTO)
    torch.cuda.synchronize()
    ref = oracle_prefill(dq.cpu(), dk.cpu(), dv.cpu(), hd ** -0.5, True, 64, 0)
    check_close(o, ref, "attention after prefill_prepare")


# ---- fused output gather (vats_attn_prefill_gather): several "ranks" emulated on one GPU — every launch stores its
#      block of the gathered tensor into ALL copies (here: plain device buffers instead of peer mappings)
@pytest.mark.parametrize("split", ["batch", "groups"])
def test_prefill_gather_writes_every_copy(split):
    N, T, H, G, hd = 4, 300, 8, 4, 64
    q, k, v = make_qkv(N, T, T, H, G, hd, seed=91)
    dq, dk, dv = q.cuda(), k.cuda(), v.cuda()
    scale, causal, left = hd ** -0.5, True, 120
    full = ops.gqa_swa_prefill(dq, dk, dv, None, None, scale, causal, left, 0, TC)
    world = 2
    copies = [torch.full((N, T, H, hd), float("nan"), dtype=torch.bfloat16, device="cuda") for _ in range(world)]
    ptrs = [c.data_ptr() for c in copies]
    hpg = H // G
    for rank in range(world):
        if split == "batch":
            b0, b1, g0, g1 = rank * 2, rank * 2 + 2, 0, G
        else:
            b0, b1, g0, g1 = 0, N, rank * 2, rank * 2 + 2
        ql, kl, vl = dq[b0:b1, :, g0 * hpg:g1 * hpg], dk[b0:b1, :, g0:g1], dv[b0:b1, :, g0:g1]
        ops.gqa_swa_prefill_gather(ql, kl, vl, copies[rank], ptrs, rank, b0, g0 * hpg, None, None, scale, causal, left, 0)
        assert _ffi.last_kernel() == "prefill_tc"
    torch.cuda.synchronize()
    for c in copies:
        assert torch.equal(c, full)


def test_prefill_gather_rejects_unaddressable_output():
    N, T, H, G, hd = 1, 200, 4, 2, 60       # dense hd 60: rows only 8-byte aligned -> no TMA tile stores
    q, k, v = make_qkv(N, T, T, H, G, hd, seed=92)
    out = torch.zeros(N, T, H, hd, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(_ffi.VatsAttnError) as e:
        ops.gqa_swa_prefill_gather(q.cuda(), k.cuda(), v.cuda(), out, [out.data_ptr()], 0, 0, 0, None, None, 0.1, True,
                                   -1, 0)
    assert e.value.code == 2
