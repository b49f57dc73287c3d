"""Dump the resident-K/V kernel's block-0 timeline: python tools/trace_mid.py N T H G hd [first] [count]
(needs the -DVATS_ENABLE_TRACE build: VATS_ATTN_LIB=.../libvats_attn_trace.so)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vats_multimodal_lm_b200 import ops, _ffi
N, T, H, G, hd = (int(x) for x in sys.argv[1:6])
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda s: torch.nn.functional.normalize(torch.randn(s, generator=g, device="cuda"), dim=-1).bfloat16()
q, k, v = mk((N, T, H, hd)), mk((N, T, G, hd)), mk((N, T, G, hd))
if "--pad" in sys.argv and hd % 8:
    def pad(x):
        buf = torch.zeros(*x.shape[:-1], (hd + 7) // 8 * 8, dtype=x.dtype, device="cuda"); buf[..., :hd] = x; return buf[..., :hd]
    q, k, v = pad(q), pad(k), pad(v)
bound = 1.0 if "--bound" in sys.argv else 0.0
f = lambda: ops.gqa_swa_prefill(q, k, v, None, None, hd ** -0.5, False, -1, -1, ops.KERNEL_MID, bound)
for _ in range(3): f()
torch.cuda.synchronize()
cap, NROLES = 4000, 4
buf = torch.zeros(NROLES * 2 * cap, dtype=torch.int64, device="cuda")
_ffi.load().vats_attn_debug_set_trace(buf.data_ptr(), cap)
f(); torch.cuda.synchronize()
_ffi.load().vats_attn_debug_set_trace(None, 0)
b = buf.cpu().view(NROLES, cap, 2).tolist()
recs = sorted((clk, role, tag) for role in range(NROLES) for tag, clk in b[role] if clk)
t0 = recs[0][0]
rolename = ["qprd", "mma ", "smxA", "epil"]
names = {0x100: "PV: wait p_full f=", 0x110: "PV: p (and v) ready f=", 0x120: "S: wait q f=", 0x130: "S: q,k ready f=", 0x140: "S: slot free, issue f=",
         0x200: "wait s f=", 0x210: "s ready f=", 0x220: "  max exchanged f=", 0x230: "p written f=", 0x240: "wait o f=", 0x250: "o ready f=",
         0x260: "stored f=", 0x300: "q issued f=", 0x280: "   S in regs f=", 0x290: "   local max done f=", 0x2a0: "     exp group ", 0x270: "    exp block "}
first = int(sys.argv[6]) if len(sys.argv) > 6 and sys.argv[6].isdigit() else 0
count = int(sys.argv[7]) if len(sys.argv) > 7 and sys.argv[7].isdigit() else 200
for c, role, tag in recs[first:first + count]:
    print(f"{c - t0:9d}  {rolename[role]}  {names.get(tag & ~0xf, hex(tag & ~0xf))}{tag & 0xf}")
print("total cycles", recs[-1][0] - t0, "records", len(recs))
