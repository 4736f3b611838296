"""Write-only HBM bandwidth (fill kernels) next to the copy bandwidth MEASURED_PEAKS.json quotes: the grid tile kernel
writes 12 B per point and reads almost nothing, so a pure-write figure is the tighter ceiling for it."""
import time

import torch

n = 3 * (1 << 30)  # 12 GiB as float32
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n // 4, dtype=torch.float32, device="cuda")
z = torch.empty(n // 4, dtype=torch.float32, device="cuda")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


t = timed(lambda: x.fill_(1.0))
print(f"fill   {4 * n / t / 1e9:8.1f} GB/s written")
t = timed(lambda: x.zero_())
print(f"memset {4 * n / t / 1e9:8.1f} GB/s written")
t = timed(lambda: z.copy_(y))
print(f"copy   {2 * 4 * (n // 4) / t / 1e9:8.1f} GB/s read + written")
