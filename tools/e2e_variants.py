import os, sys, time, torch
sys.path.insert(0, "hri-emo_b200")
from hriemo import pipeline
from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder
dev = torch.device("cuda"); torch.manual_seed(0)
model = FusionWithEmotionDecoder().eval().to(dev)
B, Ta, Tt = 4096, 500, 64
ha = torch.empty((B, Ta, 768)).pin_memory(); ht = torch.empty((B, Tt, 768)).pin_memory(); ha.normal_(); ht.normal_()
def t(**kw):
    for _ in range(2): pipeline.forward_from_host(model, ha, ht, device=dev, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): pipeline.forward_from_host(model, ha, ht, device=dev, **kw)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / 3 * 1e3
for kw in (dict(slab=512, host_cast_every=0), dict(slab=512, host_cast_every=0, out_device="cuda"), dict(slab=512, host_cast_every=2, out_device="cuda"),
           dict(slab=256, host_cast_every=2, out_device="cuda"), dict(slab=512, host_cast_every=2)):
    print(kw, f"{t(**kw):.1f} ms")
