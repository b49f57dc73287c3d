#!/usr/bin/env python
"""Run one named workload a few times (for ncu captures): python tools/run_workload.py cfg5 [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vats_multimodal_lm_b200 import ops  # noqa: E402


def rnd(shape, seed, norm):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(shape, generator=g, device="cuda")
    if norm:
        x = torch.nn.functional.normalize(x, dim=-1)
    return x.bfloat16()


W = {
    # name: (N, T, H, G, hd, causal, left)
    "cfg1": (1, 4096, 24, 8, 60, True, 384),
    "cfg3": (256, 196, 16, 8, 72, False, -1),
    "cfg4a": (512, 196, 32, 8, 66, False, -1),
    "cfg4b": (12544, 8, 32, 8, 66, False, -1),
    "cfg5": (1, 32768, 32, 8, 128, True, 4096),
    "cfg5b2": (2, 32768, 32, 8, 128, True, 4096),
    "cfg1s": (1, 384, 24, 8, 60, True, 384),
    "cfg1t": (1, 32, 24, 8, 60, True, 384),
}


def main():
    name = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    if name in ("cfg2", "cfg2m"):
        B, S, H, G, hd, left = (64, 8192, 32, 8, 128, 4096) if name == "cfg2" else (64, 8192, 24, 8, 60, 4096)
        kc, vc, q = rnd((B, S, G, hd), 1, True), rnd((B, S, G, hd), 2, False), rnd((B, H, hd), 3, True)
        if hd % 8:   # the KVCache module keeps rows of hd rounded up to 8 elements (TMA-addressable)
            def pad(x):
                buf = torch.zeros(*x.shape[:-1], (hd + 7) // 8 * 8, dtype=x.dtype, device="cuda")
                buf[..., :hd] = x
                return buf[..., :hd]
            kc, vc = pad(kc), pad(vc)
        lens = torch.full((B,), S, dtype=torch.int32, device="cuda")
        for _ in range(reps):
            o = ops.gqa_swa_decode(q, kc, vc, lens, hd ** -0.5, left)
        if "--time" in sys.argv:   # the cache (>= 0.5 GB) is larger than L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(50):
                e0.record()
                o = ops.gqa_swa_decode(q, kc, vc, lens, hd ** -0.5, left)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            from vats_multimodal_lm_b200 import _ffi
            print(name, "ms/call", tot / 50, "kernel", _ffi.last_kernel())
    else:
        N, T, H, G, hd, causal, left = W[name]
        q, k, v = rnd((N, T, H, hd), 1, True), rnd((N, T, G, hd), 2, True), rnd((N, T, G, hd), 3, False)
        if "--module-layout" in sys.argv and hd % 8 != 0:   # head stride padded to 8 elements, as the modules produce
            def pad(x):
                buf = torch.zeros(*x.shape[:-1], (hd + 7) // 8 * 8, dtype=x.dtype, device="cuda")
                buf[..., :hd] = x
                return buf[..., :hd]
            q, k, v = pad(q), pad(k), pad(v)
        bound = 1.0 if "--bound" in sys.argv else 0.0   # unit-norm q, k: what the modules pass behind qk-norm
        f = lambda: ops.gqa_swa_prefill(q, k, v, None, None, hd ** -0.5, causal, left, 0 if causal else -1, 0, bound)
        for _ in range(reps):
            o = f()
        if "--time" in sys.argv:   # CUDA-event timing of 20 further calls (inputs > L2 or L2 flushed by the next call's data)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if "--flush" in sys.argv else None
            tot = 0.0
            for _ in range(20):
                if flush is not None:
                    flush.zero_()
                e0.record()
                o = f()
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            from vats_multimodal_lm_b200 import _ffi
            print(name, "ms/call", tot / 20, "kernel", _ffi.last_kernel())
    torch.cuda.synchronize()
    print(name, "ok", float(o.float().abs().mean()))


if __name__ == "__main__":
    main()
