"""HBM calibration next to MEASURED_PEAKS.json: read-only (torch reduction), copy, and fill rates on this GPU."""
import torch
x = torch.empty(2 * 1024 ** 3, dtype=torch.bfloat16, device="cuda").normal_()   # 4 GiB
y = torch.empty_like(x)
def timeit(f, reps=10):
    for _ in range(2): f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
n = x.numel() * 2
t = timeit(lambda: x.view(torch.int32).sum()); print(f"read-only (int32 sum): {n / t / 1e12:.2f} TB/s")
t = timeit(lambda: x.view(torch.float32).amax()); print(f"read-only (fp32 amax): {n / t / 1e12:.2f} TB/s")
t = timeit(lambda: y.copy_(x)); print(f"copy (read+write): {2 * n / t / 1e12:.2f} TB/s")
t = timeit(lambda: y.zero_()); print(f"fill (write-only): {n / t / 1e12:.2f} TB/s")
