"""Timeline of one forward_from_host call: per-slab copy and compute intervals (CUDA events)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import pipeline, engine as E
from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder
dev = torch.device("cuda")
torch.manual_seed(0)
model = FusionWithEmotionDecoder().eval().to(dev)
B, Ta, Tt, slab = 4096, 500, 64, 512
ha = torch.empty((B, Ta, 768)).pin_memory(); ht = torch.empty((B, Tt, 768)).pin_memory()
ha.normal_(); ht.normal_()
# hand-rolled version of the fp32 path with events
bufs = [(torch.empty((slab, Ta, 768), device=dev), torch.empty((slab, Tt, 768), device=dev)) for _ in range(2)]
copy = torch.cuda.Stream()
main = torch.cuda.current_stream()
def run(do_compute, do_copy):
    ev = []
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    t00 = torch.cuda.Event(enable_timing=True); t00.record()
    torch.cuda.synchronize(); w0 = time.perf_counter()
    for i, s in enumerate(range(0, B, slab)):
        a, t = bufs[i % 2]
        c0, c1, k0, k1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        with torch.cuda.stream(copy):
            copy.wait_event(consumed[i % 2])
            c0.record(copy)
            if do_copy:
                a.copy_(ha[s:s + slab], non_blocking=True); t.copy_(ht[s:s + slab], non_blocking=True)
            c1.record(copy)
        main.wait_event(c1)
        k0.record(main)
        if do_compute:
            xa = E.to_seq(a, "a").x.view(slab, Ta, -1); xt = E.to_seq(t, "t").x.view(slab, Tt, -1)
            consumed[i % 2].record(main)
            model(xa, xt)
        else:
            consumed[i % 2].record(main)
        k1.record(main)
        ev.append((c0, c1, k0, k1))
    cpu_enqueue = time.perf_counter() - w0
    torch.cuda.synchronize(); wall = time.perf_counter() - w0
    print(f"compute={do_compute} copy={do_copy}: wall {wall*1e3:.1f} ms, CPU enqueue {cpu_enqueue*1e3:.1f} ms")
    for i, (c0, c1, k0, k1) in enumerate(ev):
        print(f"  slab {i}: copy {t00.elapsed_time(c0):6.1f} -> {t00.elapsed_time(c1):6.1f} ({c0.elapsed_time(c1):5.1f})   compute {t00.elapsed_time(k0):6.1f} -> {t00.elapsed_time(k1):6.1f} ({k0.elapsed_time(k1):5.1f})")
run(True, True); run(True, True); run(True, False); run(False, True)
