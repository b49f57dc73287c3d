"""Latency probe: time small prefill problems to expose the per-item critical path."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vats_multimodal_lm_b200 import ops

def rnd(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(shape, generator=g, device="cuda"), dim=-1).bfloat16()

def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts)//2] * 1e3

cases = []
for hd in (64, 72, 80, 128):
    for T in (128, 196, 256):
        cases += [(37, T, 16, 8, hd, False), (148, T, 16, 8, hd, False)]
for (N, T, H, G, hd, causal) in cases:
    q, k, v = rnd((N, T, H, hd), 1), rnd((N, T, G, hd), 2), rnd((N, T, G, hd), 3)
    us = timeit(lambda: ops.gqa_swa_prefill(q, k, v, None, None, hd ** -0.5, causal, -1, 0 if causal else -1, 0))
    items = N * G * ((H // G + 1) // 2) * ((T + 127) // 128)
    print(f"N={N:4d} T={T:5d} hd={hd:3d} causal={causal!s:5}: {us:8.1f} us  items={items:5d} items/SM={items/148:6.2f}")
