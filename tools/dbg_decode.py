import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vats_multimodal_lm_b200 import ops
B, S, H, G, hd, left = int(sys.argv[1]), int(sys.argv[2]), 32, 8, 128, int(sys.argv[3])
g = torch.Generator(device="cuda").manual_seed(1)
kc = torch.randn((B, S, G, hd), generator=g, device="cuda").bfloat16()
vc = torch.randn((B, S, G, hd), generator=g, device="cuda").bfloat16()
q = torch.randn((B, H, hd), generator=g, device="cuda").bfloat16()
lens = torch.full((B,), S, dtype=torch.int32, device="cuda")
for i in range(int(os.environ.get("ITERS", "3"))):
    o = ops.gqa_swa_decode(q, kc, vc, lens, hd ** -0.5, left)
    torch.cuda.synchronize()
    print("iter", i, "ok", float(o.float().abs().mean()), flush=True)
