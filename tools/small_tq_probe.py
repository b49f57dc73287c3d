import sys, torch
sys.path.insert(0, ".")
from vats_multimodal_lm_b200 import ops
def rnd(shape, seed, norm=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(shape, generator=g, device="cuda")
    return (torch.nn.functional.normalize(x, dim=-1) if norm else x).bfloat16()
for (N, Tq, Tk) in [(64, 4, 8192), (8, 8, 4096), (64, 15, 512), (256, 2, 64)]:
    q, k, v = rnd((N, Tq, 32, 128), 1), rnd((N, Tk, 8, 128), 2), rnd((N, Tk, 8, 128), 3, False)
    res = {}
    for name, kern in (("simt", 2), ("tc", 1)):
        f = lambda: ops.gqa_swa_prefill(q, k, v, None, None, 128 ** -0.5, True, 4096, 0, kern)
        for _ in range(3): o = f()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(10): o = f()
        b.record(); torch.cuda.synchronize()
        res[name] = (a.elapsed_time(b) / 10, o)
    err = (res["simt"][1].float() - res["tc"][1].float()).abs().max().item()
    print(f"N={N} Tq={Tq} Tk={Tk}: simt {res['simt'][0]:.3f} ms, tc {res['tc'][0]:.3f} ms, max diff {err:.2e}")
