"""How fast can the host convert fp32 features to bf16 into pinned memory: torch CPU copy_ vs the library's
hriemo_host_pack_bf16, alone and while an H2D copy is in flight."""
import time, torch, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import ops
print("cpus", os.cpu_count(), "torch threads", torch.get_num_threads())
n_utt, T, d = 512, 564, 768
x = torch.randn(n_utt, T, d).pin_memory()
y = torch.empty(n_utt * T * d, dtype=torch.bfloat16).pin_memory()
def bench(fn, label):
    for _ in range(2): fn()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    dt = (time.perf_counter() - t0) / 5
    print(f"  {label}: {dt*1e3:.1f} ms per 512-utterance slab = {n_utt / dt / 1e3:.1f} utt/ms, {x.numel()*4/dt/1e9:.1f} GB/s of fp32 input")
def run_all(tag):
    print(tag)
    bench(lambda: y.view(n_utt, T, d).copy_(x), "torch copy_")
    for nt in (4, 8, 16, 32, 64):
        bench(lambda: ops.host_pack_bf16(x, y, T, threads=nt), f"host_pack {nt:2d} threads")
run_all("alone")
d_buf = torch.empty(n_utt * T * d, dtype=torch.float32, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(400): d_buf.copy_(x.view(-1), non_blocking=True)
run_all("with concurrent H2D")
torch.cuda.synchronize()
