"""torchrun --nproc-per-node 2 tools/train_overlap_check.py: three data-parallel steps with the overlapped exchange
(eager and CUDA-graph replay) must leave the parameters of the plain one-all-reduce step (bit-identical at 2 ranks: a
two-term mean does not depend on how the arena is cut); prints the step times of both."""
import os, sys, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo.train import Trainer
from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder
rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator(device=dev).manual_seed(100 + rank)
h_a = torch.randn(B, 500, 768, device=dev, generator=g); h_t = torch.randn(B, 64, 768, device=dev, generator=g)
y = torch.eye(4, device=dev)[torch.randint(0, 4, (B,), device=dev, generator=g)]
res = {}
for name, kw in (("plain", dict(overlap=False)), ("overlap", dict(overlap=True)), ("overlap_graph", dict(overlap=True, graph=True)),
                 ("plain_graph", dict(overlap=False, graph=True))):
    torch.manual_seed(0)
    model = FusionWithEmotionDecoder(dropout=0.0).to(dev)
    tr = Trainer(model, **kw)
    for _ in range(4): info = tr.step(h_a, h_t, None, None, y)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): info = tr.step(h_a, h_t, None, None, y)
    e1.record(); torch.cuda.synchronize()
    res[name] = (tr.params.clone(), float(info["loss"]), e0.elapsed_time(e1) / 5)
    del tr, model
    torch.cuda.empty_cache()
ref = res["plain"][0]
for name, (p, loss, ms) in res.items():
    diff = float((p - ref).abs().max())
    if rank == 0:
        print(f"{name:14s} loss {loss:.6f}  {ms:7.2f} ms/step  max |param - plain| = {diff:.3e}", flush=True)
    assert diff == 0.0 or name.endswith("graph") and diff < 1e-6, (name, diff)
dist.destroy_process_group()
