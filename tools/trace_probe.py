"""Dump the prefill kernel's block-0 timeline for a small problem: python tools/trace_probe.py N T hd [max_records]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vats_multimodal_lm_b200 import ops, _ffi
N, T, hd = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
H, G = 16, 8
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda s: torch.nn.functional.normalize(torch.randn(s, generator=g, device="cuda"), dim=-1).bfloat16()
q, k, v = mk((N, T, H, hd)), mk((N, T, G, hd)), mk((N, T, G, hd))
f = lambda: ops.gqa_swa_prefill(q, k, v, None, None, hd ** -0.5, False, -1, -1, 0)
for _ in range(3): f()
torch.cuda.synchronize()
cap = 2000
NROLES = 4
buf = torch.zeros(NROLES * 2 * cap, dtype=torch.int64, device="cuda")
_ffi.load().vats_attn_debug_set_trace(buf.data_ptr(), cap)
f(); torch.cuda.synchronize()
_ffi.load().vats_attn_debug_set_trace(None, 0)
b = buf.cpu().view(NROLES, cap, 2).tolist()
recs = []
for role in range(NROLES):
    for tag, clk in b[role]:
        if clk: recs.append((clk, role, tag))
recs.sort()
t0 = recs[0][0]
rolename = ["prod", "mma ", "smx0", "smx1"]
names = {0x100: "item start (q ready)", 0x101: "k0 ready", 0x140: "item done (last PV issued)", 0x230: "wait o_full", 0x231: "o_full ok",
         0x240: "epilogue done", 0x300: "q issued"}
lim = int(sys.argv[4]) if len(sys.argv) > 4 else 150
skip = int(sys.argv[5]) if len(sys.argv) > 5 else 0
for (c, role, tag) in recs[skip:skip + lim]:
    base = tag & ~0xf
    nm = names.get(tag) or {0x110: "v ready j=", 0x120: "p0 ready j=", 0x130: "p1 ready j=", 0x200: "wait s j=", 0x210: "s ready j=",
                            0x220: "p written j=", 0x250: "  S in regs j=", 0x260: "  max/rescale done j=", 0x270: "  exp done j=", 0x310: "k slot free j="}.get(base, hex(base)) + str(tag & 0xf)
    print(f"{c - t0:9d}  {rolename[role]}  {nm}")
