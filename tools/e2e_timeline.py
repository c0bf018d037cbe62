"""Timeline of forward_from_host calls (its `trace` argument): per-slab copy and compute intervals.
    python tools/e2e_timeline.py [--slab N] [--every K] [--no-ramp] [--ragged] [--bucket]"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import pipeline
from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default

dev = torch.device("cuda")
torch.manual_seed(0)
model = FusionWithEmotionDecoder().eval().to(dev)
B, Ta, Tt = arg("--batch", 4096), 500, 64
slab, every = arg("--slab", 512), arg("--every", 2)
ramp, ragged, bucket = "--no-ramp" not in sys.argv, "--ragged" in sys.argv, "--bucket" in sys.argv
ha = torch.empty((B, Ta, 768)).pin_memory(); ht = torch.empty((B, Tt, 768)).pin_memory()
ha.normal_(); ht.normal_()
ma = mt = None
if ragged:
    g = torch.Generator().manual_seed(1)
    la = torch.randint(Ta // 2, Ta + 1, (B,), generator=g); lt = torch.randint(Tt // 2, Tt + 1, (B,), generator=g)
    ma = (torch.arange(Ta)[None] >= la[:, None]).pin_memory(); mt = (torch.arange(Tt)[None] >= lt[:, None]).pin_memory()
print(f"B={B} slab={slab} host_cast_every={every} ramp={ramp} ragged={ragged} bucket={bucket} threads={torch.get_num_threads()}")
for rep in range(3):
    trace = [] if rep == 2 else None
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    w0 = time.perf_counter()
    pipeline.forward_from_host(model, ha, ht, ma, mt, device=dev, slab=slab, host_cast_every=every, ramp=ramp, bucket=bucket, trace=trace)
    wall = time.perf_counter() - w0
    print(f"  call {rep}: wall {wall*1e3:.1f} ms = {B / wall:.0f} utt/s")
for r in trace:
    print(f"  slab {r['slab']:2d} n={r['n']:4d} T_a={r['T_a']:3d} {'host' if r['host_cast'] else 'fp32'}: copy {t0.elapsed_time(r['copy0']):6.1f} -> {t0.elapsed_time(r['copy1']):6.1f} ({r['copy0'].elapsed_time(r['copy1']):5.1f})"
          f"   compute {t0.elapsed_time(r['comp0']):6.1f} -> {t0.elapsed_time(r['comp1']):6.1f} ({r['comp0'].elapsed_time(r['comp1']):5.1f})")
