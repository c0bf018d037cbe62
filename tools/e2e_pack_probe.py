"""How long the host-side bf16 conversion of a slab takes INSIDE forward_from_host (two calls in flight, as bench.py's e2e
leg), how many slabs the host side gets, and the step time: python tools/e2e_pack_probe.py [threads]"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import pipeline
from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

threads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
torch.set_num_threads(threads)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = FusionWithEmotionDecoder().eval().to(dev)
B, T_a, T_t = 4096, 500, 64
ha = torch.empty(B, T_a, 768).pin_memory(); ht = torch.empty(B, T_t, 768).pin_memory()
ha.normal_(); ht.normal_()
packs = []
orig = pipeline.TwoEndedPlan.publish
def publish(self, slab, k, seconds):
    packs.append((slab, k, seconds))
    return orig(self, slab, k, seconds)
pipeline.TwoEndedPlan.publish = publish
def stream(n):
    pend = None
    for _ in range(n):
        nxt = pipeline.forward_from_host(model, ha, ht, device=dev, slab=512, out_device="cpu", wait=False)
        if pend is not None: pend.wait()
        pend = nxt
    pend.wait()
# the device-resident forward of the same batch on this box (what bench.py's `value` times), in slabs of 2048
def resident(n):
    for _ in range(n):
        for lo in range(0, B, 2048):
            model(xa[lo:lo + 2048], xt[lo:lo + 2048])
xa, xt = ha.to(dev), ht.to(dev)
resident(2)
torch.cuda.synchronize(); t0 = time.perf_counter(); resident(4); torch.cuda.synchronize()
print(f"device-resident forward: {(time.perf_counter() - t0) / 4 * 1e3:.1f} ms per step")
del xa, xt
torch.cuda.empty_cache()
# bench.py's sequence: one call at a time (3), then two warm-up calls of the stream and 5 timed ones from a synchronise
for _ in range(1): pipeline.forward_from_host(model, ha, ht, device=dev, slab=512, out_device="cpu")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): pipeline.forward_from_host(model, ha, ht, device=dev, slab=512, out_device="cpu")
torch.cuda.synchronize(); print(f"one call at a time: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per step")
stream(2)
for n in (5, 5, 10):
    pipeline.reset_stats()
    torch.cuda.synchronize(); t0 = time.perf_counter(); stream(n); torch.cuda.synchronize()
    print(f"stream of {n} from a synchronise: {(time.perf_counter() - t0) / n * 1e3:.1f} ms per step, host-cast slabs {pipeline.STATS['host_cast_slabs'] / n:.1f}")
stream(3)
for rep, host_reserve in enumerate([2, 0]):
    pipeline.HOST_RESERVE = host_reserve
    print("HOST_RESERVE", host_reserve)
    packs.clear(); pipeline.reset_stats()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    stream(8)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = dict(pipeline.STATS)
    print(f"threads {threads}: {8 * B / dt:.0f} utt/s, {dt / 8 * 1e3:.1f} ms per step, host-cast slabs per step "
          f"{st['host_cast_slabs'] / 8:.1f}, h2d {st['h2d_bytes'] / 8 / 1e9:.2f} GB per step")
    print("conversion seconds per slab:", [round(s, 4) for _, _, s in packs])

# ---- decision log of ONE call in the stream (the 4th): who took which slab when
log = []
t00 = [0.0]
oc, on, op = pipeline.TwoEndedPlan.claim_back, pipeline.TwoEndedPlan.next, orig
def claim_back(self):
    r = oc(self); log.append((time.perf_counter() - t00[0], "host claims", r, self.front, self.back, self.t_pack, self.t_step)); return r
def nxt(self, paced=True):
    r = on(self, paced); log.append((time.perf_counter() - t00[0], "copy takes", r, self.front, self.back, None, None)); return r
def pub(self, slab, k, seconds):
    log.append((time.perf_counter() - t00[0], "host done", slab, self.front, self.back, seconds, None)); return op(self, slab, k, seconds)
pipeline.TwoEndedPlan.claim_back, pipeline.TwoEndedPlan.next, pipeline.TwoEndedPlan.publish = claim_back, nxt, pub
pipeline.HOST_RESERVE = 2
t00[0] = time.perf_counter()
stream(4)
torch.cuda.synchronize()
for row in log:
    print("%8.1f ms  %-12s %-10s front=%s back=%s t_pack=%s t_step=%s" % (row[0] * 1e3, row[1], row[2], row[3], row[4],
          None if row[5] is None else round(row[5] * 1e3, 1), None if row[6] is None else round(row[6] * 1e3, 1)))
