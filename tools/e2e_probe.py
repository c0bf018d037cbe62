"""Where does forward_from_host spend its time?  Per-slab copy / compute events."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import pipeline
from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder
dev = torch.device("cuda")
torch.manual_seed(0)
model = FusionWithEmotionDecoder().eval().to(dev)
B, Ta, Tt = 4096, 500, 64
ha = torch.empty((B, Ta, 768)).pin_memory(); ht = torch.empty((B, Tt, 768)).pin_memory()
ha.normal_(); ht.normal_()
for every in (0, 2, 3):
  for slab in (256, 512):
    for _ in range(2):
        pipeline.forward_from_host(model, ha, ht, device=dev, slab=slab, host_cast_every=every)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        pipeline.forward_from_host(model, ha, ht, device=dev, slab=slab, host_cast_every=every)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"host_cast_every {every} slab {slab}: {dt*1e3:.1f} ms  {B/dt:.0f} utt/s")
# staging path with a no-op model: pure copy pipeline
class Nop(torch.nn.Module):
    def forward(self, a, t, ma=None, mt=None):
        z = torch.zeros(a.shape[0], 4, device=a.device)
        return z, z[:, :1], z
for slab in (256, 512):
    pipeline.forward_from_host(Nop(), ha, ht, device=dev, slab=slab)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pipeline.forward_from_host(Nop(), ha, ht, device=dev, slab=slab)
    torch.cuda.synchronize(); print(f"copy+cast only, slab {slab}: {(time.perf_counter() - t0) * 1e3:.1f} ms")
da, dt_ = ha[:512].to(dev), ht[:512].to(dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(0, B, 512): model(da, dt_)
torch.cuda.synchronize(); print("compute only (8 slabs of 512):", (time.perf_counter() - t0) * 1e3, "ms")
