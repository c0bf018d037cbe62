"""BASELINE.json config 4: large-batch inference sweep (B = 256 ... 16384, T_a up to 1000) on one GPU.
Prints one JSON line per point: utterances/s (device-resident inputs, CUDA events, 3 warm-up + 3 timed
forwards) and the fraction of the sustained tensor roofline of the whole path (SURVEY 8(d) FLOPs)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from bench import flops_per_utt, load_peaks
from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

def main():
    dev = torch.device("cuda")
    peaks = load_peaks()
    torch.manual_seed(1234)
    model = FusionWithEmotionDecoder().eval().to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    points = [(B, Ta) for Ta in (300, 500, 1000) for B in (256, 1024, 4096, 16384)]
    for B, Ta in points:
        Tt = 64
        if B * (Ta + Tt) * 768 * 4 > 60e9:
            continue
        h_a = torch.randn(B, Ta, 768, generator=g, device=dev); h_t = torch.randn(B, Tt, 768, generator=g, device=dev)
        for _ in range(3): model(h_a, h_t)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): lo, be, z = model(h_a, h_t)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        ups = B / ms * 1e3
        tf = ups * flops_per_utt(Ta, Tt) / 1e12
        print(json.dumps({"B": B, "T_a": Ta, "T_t": Tt, "ms_per_forward": round(ms, 3), "utt_per_s": round(ups, 1),
                          "path_tflops": round(tf, 1), "frac_of_sustained_tensor_peak": round(tf / peaks["tf_sust"], 3),
                          "finite": bool(torch.isfinite(lo).all())}), flush=True)
        del h_a, h_t, lo, be, z
        torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
