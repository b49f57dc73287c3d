"""Image-gen cross-attention geometry (reference cross_attention.py main(): Tq = 72*144 image tokens, Tk = 16 text tokens,
32 heads, hd 16): time the kernel AUTO picks.  VATS_PREFILL_SHORT=0 selects the generic warp kernel for comparison."""
import sys, torch
sys.path.insert(0, ".")
from vats_multimodal_lm_b200 import ops
for B in (2, 16):
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(B, 10368, 32, 16, generator=g, device="cuda").bfloat16()
    k = torch.randn(B, 16, 32, 16, generator=g, device="cuda").bfloat16()
    v = torch.randn(B, 16, 32, 16, generator=g, device="cuda").bfloat16()
    kv = torch.rand(B, 16, device="cuda") > 0.3; kv[:, 0] = True
    for kern, name in ((0, "auto"), (1, "tcgen05")):
        f = lambda: ops.gqa_swa_prefill(q, k, v, None, kv, 0.25, False, -1, -1, kern)
        for _ in range(3): f()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(20): f()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print(f"cross-attn B={B} {name}: {ms:.4f} ms, {2 * q.numel() * 2 / ms / 1e6:.0f} GB/s")
