"""Is the bounded-logit path really taken?  Inputs that violate the bound must change the result."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vats_multimodal_lm_b200 import ops, _ffi
g = torch.Generator(device="cuda").manual_seed(1)
for (N, T, H, G, hd) in [(80, 196, 4, 2, 64), (1, 700, 4, 2, 64)]:
    q = (torch.randn(N, T, H, hd, generator=g, device="cuda") * 3).bfloat16()
    k = (torch.randn(N, T, G, hd, generator=g, device="cuda") * 3).bfloat16()
    v = torch.randn(N, T, G, hd, generator=g, device="cuda").bfloat16()
    o0 = ops.gqa_swa_prefill(q, k, v, None, None, 1.0, False, -1, -1, 0, 0.0)
    o1 = ops.gqa_swa_prefill(q, k, v, None, None, 1.0, False, -1, -1, 0, 1.0)
    torch.cuda.synchronize()
    print(_ffi.last_kernel(), "finite exact:", bool(torch.isfinite(o0.float()).all()), "finite bounded:",
          bool(torch.isfinite(o1.float()).all()), "max diff", (o0.float() - o1.float()).abs().nan_to_num(99).max().item())
