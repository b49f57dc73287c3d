"""GPU: backward of the attention core (vats_attn_prefill_backward behind `register_autograd`) against torch.autograd
through the fp32 oracle on the same bf16-rounded inputs — SURVEY.md §8f rank 4.  The reference trains through these
modules (training/transformers/nlp/loops/training_loop.py:54-65) and its own attention test back-propagates
(tests/transformers/nlp/attention_tests.py `test_gradients`).

Stated tolerance for gradients (bf16 operands, bf16 P / dS fed to the tensor cores, fp32 accumulation, bf16 outputs):
relative L2 error <= 2e-2 per tensor and max-abs error <= 2e-2 * max|reference gradient| + 1e-3."""
import math

import pytest
import torch

from conftest import make_qkv
import vats_multimodal_lm_b200 as vl
from vats_multimodal_lm_b200 import ops
from oracle import mask_predicate, sdpa_explicit

pytestmark = pytest.mark.gpu


def _ref_grads(q, k, v, dout, scale, causal, left, right, q_valid=None, k_valid=None):
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    m = mask_predicate(q.size(0), q.size(1), k.size(1), causal, left, right, q_valid, k_valid)
    o = sdpa_explicit(qf, kf, vf, m, scale)
    o.backward(dout.float())
    return qf.grad, kf.grad, vf.grad


def _check(got, ref, what):
    got, ref = got.float().cpu(), ref.float()
    assert torch.isfinite(got).all(), what
    rel = (got - ref).norm().item() / max(ref.norm().item(), 1e-12)
    max_abs = (got - ref).abs().max().item()
    lim = 2e-2 * ref.abs().max().item() + 1e-3
    assert rel <= 2e-2 and max_abs <= lim, f"{what}: rel_l2={rel:.3e} max_abs={max_abs:.3e} (limit {lim:.3e})"


SHAPES = [  # N, Tq, Tk, H, G, hd
    (2, 200, 200, 4, 2, 64),
    (1, 128, 128, 2, 1, 128),
    (2, 77, 250, 4, 4, 64),      # Tq != Tk
    (3, 196, 196, 4, 2, 72),     # ViT-2D geometry
    (2, 196, 196, 8, 2, 66),     # ViT-3D spatial geometry (hd 66 -> padded to 80 in the kernels)
    (2, 150, 150, 6, 2, 60),     # LLM medium geometry, H/G = 3
    (1, 300, 140, 2, 2, 32),     # more queries than keys: fully masked rows under the causal mask
    (4, 8, 8, 8, 2, 66),         # ViT-3D temporal geometry
    (1, 40, 40, 2, 2, 16),
]
MASKS = [(True, -1, -1), (True, 50, 0), (False, -1, -1), (False, 20, 9)]


@pytest.mark.parametrize("causal,left,right", MASKS)
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_backward_matches_autograd_of_oracle(shape, causal, left, right):
    N, Tq, Tk, H, G, hd = shape
    q, k, v = make_qkv(N, Tq, Tk, H, G, hd, seed=sum(shape))
    g = torch.Generator().manual_seed(7)
    dout = torch.randn(N, Tq, H, hd, generator=g).bfloat16()
    scale = 1.0 / math.sqrt(hd)
    dq_r, dk_r, dv_r = _ref_grads(q, k, v, dout, scale, causal, left, right)
    dq_, dk_, dv_ = (t.cuda().requires_grad_(True) for t in (q, k, v))
    o = ops.gqa_swa_prefill(dq_, dk_, dv_, None, None, scale, causal, left, right, 0)
    o.backward(dout.cuda())
    _check(dq_.grad, dq_r, f"dq {shape}")
    _check(dk_.grad, dk_r, f"dk {shape}")
    _check(dv_.grad, dv_r, f"dv {shape}")


def test_backward_with_padding_masks_and_strided_views():
    N, T, H, G, hd = 3, 170, 6, 2, 64
    g = torch.Generator().manual_seed(31)
    qkv = torch.randn(N, T, (H + 2 * G) * hd, generator=g).bfloat16()
    qv = torch.rand(N, T, generator=g) > 0.25
    kv = torch.rand(N, T, generator=g) > 0.25
    kv[:, 0] = True
    dout = torch.randn(N, T, H, hd, generator=g).bfloat16()
    q, k, v = (t.reshape(N, T, -1, hd) for t in torch.split(qkv, [H * hd, G * hd, G * hd], dim=-1))
    dq_r, dk_r, dv_r = _ref_grads(q, k, v, dout, 0.1, True, 60, 0, qv, kv)
    dqkv = qkv.cuda().requires_grad_(True)
    dq_, dk_, dv_ = (t.reshape(N, T, -1, hd) for t in torch.split(dqkv, [H * hd, G * hd, G * hd], dim=-1))
    o = ops.gqa_swa_prefill(dq_, dk_, dv_, qv.cuda(), kv.cuda(), 0.1, True, 60, 0, 0)
    o.backward(dout.cuda())
    ref = torch.cat([dq_r.reshape(N, T, -1), dk_r.reshape(N, T, -1), dv_r.reshape(N, T, -1)], dim=-1)
    _check(dqkv.grad, ref, "d(qkv) through strided views")
    # masked query rows get exactly zero dq; masked keys exactly zero dk / dv
    gq = dqkv.grad[..., :H * hd].reshape(N, T, H, hd).cpu()
    gk = dqkv.grad[..., H * hd:(H + G) * hd].reshape(N, T, G, hd).cpu()
    assert torch.all(gq[~qv] == 0) and torch.all(gk[~kv] == 0)


def test_unnormalised_inputs_large_scale():
    N, T, H, G, hd = 1, 260, 4, 2, 64
    q, k, v = make_qkv(N, T, T, H, G, hd, seed=5, unit_norm=False)
    dout = torch.randn(N, T, H, hd, generator=torch.Generator().manual_seed(1)).bfloat16()
    dq_r, dk_r, dv_r = _ref_grads(q, k, v, dout, 0.125, True, -1, 0)
    a, b, c = (t.cuda().requires_grad_(True) for t in (q, k, v))
    ops.gqa_swa_prefill(a, b, c, None, None, 0.125, True, -1, 0, 0).backward(dout.cuda())
    # peaky softmax: bf16 rounding of P / dS dominates -> same relative tolerance, on larger gradients
    _check(a.grad, dq_r, "dq unnormalised")
    _check(b.grad, dk_r, "dk unnormalised")
    _check(c.grad, dv_r, "dv unnormalised")


def test_llm_module_trains_like_the_reference_test_gradients():
    """tests/transformers/nlp/attention_tests.py::test_gradients of the reference: loss = out.sum(); every parameter
    gets a gradient.  Also compared with the gradients of an fp32 PyTorch restatement of the same layer."""
    torch.manual_seed(3)
    d_model, H, G = 256, 4, 2
    hd = d_model // H
    attn = vl.Attention(d_model, H, G, 10000.0, hd ** -0.5).cuda()
    x = torch.randn(2, 70, d_model, device="cuda", requires_grad=True)
    out, _ = attn(x, 30, 0, True, None)
    out.sum().backward()
    for name, p in attn.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
    assert x.grad is not None

    # fp32 reference of the same layer on the CPU
    ref = vl.Attention(d_model, H, G, 10000.0, hd ** -0.5)
    ref.load_state_dict(attn.state_dict())
    xr = x.detach().cpu().clone().requires_grad_(True)
    qkv = ref.w_qkv(xr)
    q, k, v = torch.split(qkv, [H * hd, G * hd, G * hd], dim=-1)
    q, k, v = q.view(2, 70, H, hd), k.view(2, 70, G, hd), v.view(2, 70, G, hd)
    q, k = torch.nn.functional.normalize(q, dim=-1, eps=1e-6), torch.nn.functional.normalize(k, dim=-1, eps=1e-6)
    q, k = ref.rope(q), ref.rope(k)
    o = sdpa_explicit(q, k, v, mask_predicate(2, 70, 70, True, 30, 0), hd ** -0.5)
    ref.w_o(o.reshape(2, 70, d_model)).sum().backward()
    _check(x.grad, xr.grad, "dx")
    _check(attn.w_qkv.weight.grad, ref.w_qkv.weight.grad, "d w_qkv")
    _check(attn.w_o.weight.grad, ref.w_o.weight.grad, "d w_o")


def test_vit_modules_backward_smoke():
    torch.manual_seed(4)
    m = vl.SpatialAttention(288, 4, 2, 10000.0, 64, 16, 72 ** -0.5, False, False, True).cuda()
    x = torch.randn(3, 16, 288, device="cuda", requires_grad=True)
    m(x, False, True, -1, -1).square().mean().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
