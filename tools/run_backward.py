#!/usr/bin/env python
"""One training step of the core (forward + backward through autograd) a few times, for ncu launch lists / timing:
python tools/run_backward.py [N T H G hd left] [--time]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vats_multimodal_lm_b200 import ops  # noqa: E402
from vats_multimodal_lm_b200.modules._common import _to_kernel_layout  # noqa: E402

nums = [int(a) for a in sys.argv[1:] if a.lstrip("-").isdigit()]
N, T, H, G, hd, left = nums if len(nums) == 6 else (8, 2048, 24, 8, 60, 384)
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *s: torch.randn(*s, generator=g, device="cuda")
q = _to_kernel_layout(torch.nn.functional.normalize(mk(N, T, H, hd), dim=-1).bfloat16()).detach().requires_grad_(True)
k = _to_kernel_layout(torch.nn.functional.normalize(mk(N, T, G, hd), dim=-1).bfloat16()).detach().requires_grad_(True)
v = _to_kernel_layout(mk(N, T, G, hd).bfloat16()).detach().requires_grad_(True)
do = mk(N, T, H, hd).bfloat16()
for _ in range(3):
    o = ops.gqa_swa_prefill(q, k, v, None, None, hd ** -0.5, True, left, 0, 0)
    o.backward(do)
    q.grad = k.grad = v.grad = None
torch.cuda.synchronize()
if "--time" in sys.argv:
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(10):
        e[0].record()
        o = ops.gqa_swa_prefill(q, k, v, None, None, hd ** -0.5, True, left, 0, 0)
        e[1].record()
        o.backward(do)
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]) / 10
        tb += e[1].elapsed_time(e[2]) / 10
        q.grad = k.grad = v.grad = None
    print(f"forward {tf:.4f} ms  backward {tb:.4f} ms")
print("ok")
