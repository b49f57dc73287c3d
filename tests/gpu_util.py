"""Helpers for the GPU parity tests: run the op through torch.ops (-> ctypes -> C-ABI), compare with the oracle."""
import torch

from oracle import decode_explicit, mask_predicate, sdpa_explicit
from vats_multimodal_lm_b200 import ops

# Stated tolerance (bf16 operands and bf16 output, fp32 accumulation, vs the fp32 oracle on the same bf16-rounded
# inputs): max-abs error <= 2e-2 and relative L2 error <= 1e-2 for |v| ~ N(0,1).
MAX_ABS_TOL = 2e-2
REL_L2_TOL = 1e-2


def err_stats(out, ref):
    out = out.float().cpu()
    ref = ref.float()
    diff = (out - ref).abs()
    max_abs = diff.max().item() if diff.numel() else 0.0
    denom = ref.norm().item()
    rel_l2 = (out - ref).norm().item() / denom if denom > 0 else (out - ref).norm().item()
    return max_abs, rel_l2


def check_close(out, ref, what=""):
    assert torch.isfinite(out.float()).all(), f"{what}: non-finite values in the output"
    max_abs, rel_l2 = err_stats(out, ref)
    assert max_abs <= MAX_ABS_TOL and rel_l2 <= REL_L2_TOL, f"{what}: max_abs={max_abs:.3e} rel_l2={rel_l2:.3e}"
    return max_abs, rel_l2


def run_prefill(q, k, v, scale, causal, left, right, q_valid=None, k_valid=None, kernel=ops.KERNEL_AUTO,
                device="cuda", logit_bound=0.0):
    dq = q.to(device)
    dk = k.to(device)
    dv = v.to(device)
    qv = None if q_valid is None else q_valid.to(device)
    kv = None if k_valid is None else k_valid.to(device)
    o = ops.gqa_swa_prefill(dq, dk, dv, qv, kv, scale, causal, left, right, kernel, logit_bound)
    torch.cuda.synchronize()
    return o


def oracle_prefill(q, k, v, scale, causal, left, right, q_valid=None, k_valid=None):
    N, Tq, Tk = q.size(0), q.size(1), k.size(1)
    mask = mask_predicate(N, Tq, Tk, causal, left, right, q_valid, k_valid)
    return sdpa_explicit(q, k, v, mask, scale)
