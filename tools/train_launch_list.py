"""One eager training step (B = 512) bracketed by cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/train_launch_list.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo.train import Trainer
from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
model = FusionWithEmotionDecoder(dropout=0.0).to(dev)
tr = Trainer(model)
g = torch.Generator(device=dev).manual_seed(1)
h_a = torch.randn(B, 500, 768, device=dev, generator=g); h_t = torch.randn(B, 64, 768, device=dev, generator=g)
y = torch.eye(4, device=dev)[torch.randint(0, 4, (B,), device=dev, generator=g)]
for _ in range(2): tr.step(h_a, h_t, None, None, y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(h_a, h_t, None, None, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
