"""Pinned host -> device copy bandwidth of this box (the bound on bench.py's e2e line)."""
import torch, time
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): d.copy_(h, non_blocking=True)
b.record(); torch.cuda.synchronize()
print("H2D pinned GB/s:", 5 * n / a.elapsed_time(b) / 1e6)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d[: n // 2].copy_(h[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h[n // 2:], non_blocking=True)
torch.cuda.synchronize()
print("H2D pinned, two streams GB/s:", 5 * n / (time.perf_counter() - t0) / 1e9)
