"""Timeline of block 0 / thread 0 of the short-sequence kernel (needs the -DVATS_ENABLE_TRACE build):
VATS_ATTN_LIB=.../libvats_attn_trace.so python tools/trace_short.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vats_multimodal_lm_b200 import ops, _ffi
N, T, H, G, hd = 12544, 8, 32, 8, 66
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda s: torch.nn.functional.normalize(torch.randn(s, generator=g, device="cuda"), dim=-1).bfloat16()
q, k, v = mk((N, T, H, hd)), mk((N, T, G, hd)), mk((N, T, G, hd))
f = lambda: ops.gqa_swa_prefill(q, k, v, None, None, hd ** -0.5, False, -1, -1, 0)
for _ in range(3): f()
torch.cuda.synchronize()
cap = 4000
buf = torch.zeros(4 * 2 * cap, dtype=torch.int64, device="cuda")
_ffi.load().vats_attn_debug_set_trace(buf.data_ptr(), cap)
f(); torch.cuda.synchronize()
_ffi.load().vats_attn_debug_set_trace(None, 0)
b = buf.cpu().view(-1, 2).tolist()
recs = [(c, t) for t, c in b if c]
t0 = recs[0][0]
names = {1: "iter start", 2: "data ready", 3: "kv widened+sync", 4: "compute done (thread 0)", 5: "all done (sync)", 6: "store issued"}
prev = t0
for c, t in recs[40:100]:
    print(f"{c - t0:9d} (+{c - prev:6d})  {names.get(t, t)}")
    prev = c
