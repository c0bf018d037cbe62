"""How fast can the host convert fp32 features to bf16 (torch CPU, all threads), into pinned memory?"""
import time, torch, os
torch.set_num_threads(os.cpu_count())
n = 512 * 564 * 768   # one 512-utterance slab
x = torch.randn(n).pin_memory()
y = torch.empty(n, dtype=torch.bfloat16).pin_memory()
for _ in range(2): y.copy_(x)
t0 = time.perf_counter()
for _ in range(5): y.copy_(x)
dt = (time.perf_counter() - t0) / 5
print(f"threads {torch.get_num_threads()}: fp32->bf16 of {n*4/1e9:.2f} GB in {dt*1e3:.1f} ms = {n*4/dt/1e9:.1f} GB/s of fp32 input")
# same while an H2D copy of another buffer is in flight
d = torch.empty(n, dtype=torch.float32, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(20): d.copy_(x, non_blocking=True)
t0 = time.perf_counter()
for _ in range(5): y.copy_(x)
dt2 = (time.perf_counter() - t0) / 5
torch.cuda.synchronize()
print(f"with concurrent H2D: {dt2*1e3:.1f} ms = {n*4/dt2/1e9:.1f} GB/s")
