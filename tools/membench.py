"""What this box's HBM does for the access mixes that matter here (torch kernels, CUDA events, best of 8):
copy (50 % reads / 50 % writes: the MEASURED_PEAKS.json number), write-only (fill), read-only (sum)."""
import torch
n = 1 << 28                      # 1 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty_like(a)


def best(f, reps=8):
    t = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        t = min(t, e0.elapsed_time(e1))
    return t * 1e-3


a.normal_()
print(f"copy   (1 GiB -> 1 GiB) : {2 * n * 4 / best(lambda: b.copy_(a)) / 1e9:7.1f} GB/s")
print(f"fill   (write only)     : {n * 4 / best(lambda: b.zero_()) / 1e9:7.1f} GB/s")
print(f"sum    (read only)      : {n * 4 / best(lambda: a.sum()) / 1e9:7.1f} GB/s")
c = torch.empty(n // 4, dtype=torch.float32, device="cuda")
print(f"1 read + 4 writes (a[:n/4] -> 4 quarters of b): "
      f"{5 * (n // 4) * 4 / best(lambda: [b[q * (n // 4):(q + 1) * (n // 4)].copy_(a[:n // 4]) for q in range(4)]) / 1e9:7.1f} GB/s (4 launches)")
