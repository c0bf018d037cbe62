#!/usr/bin/env python3
"""Training-step benchmark (BASELINE config 5): FusionWithEmotionDecoder BCE step (forward with tapes, backward,
clip, AdamW) through hriemo.train.Trainer on synthetic features, one GPU or one rank per GPU under torchrun.

    python tools/bench_train.py --batch 256 --steps 5 --warmup 3 [--T_a 500 --T_t 64]

Prints one JSON line: utterances/s (whole job, max over ranks), ms per step, and the share of the step spent in each
tensor-core kernel family (CUDA events around every tagged launch: forward / dgrad GEMMs, wgrad, attention forward,
attention backward) with their achieved TFLOP/s on algorithmic FLOPs.  Inputs are resident in HBM (fp32 features)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256, help="utterances per GPU per step")
    ap.add_argument("--T_a", type=int, default=500)
    ap.add_argument("--T_t", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ragged", action="store_true", help="trailing-PAD masks, valid lengths uniform in [T/2, T]")
    ap.add_argument("--dropout", type=float, default=0.0, help="dropout rate of the model (train() mode; > 0 runs eagerly)")
    ap.add_argument("--graph", action="store_true", help="Trainer(graph=True): forward + backward replayed from a CUDA graph (no per-kernel breakdown)")
    ap.add_argument("--attn-bwd-impl", type=int, default=0, help="ops.attention_backward impl (0 ldmatrix, 2 first form)")
    args = ap.parse_args()
    import torch.distributed as dist
    from hriemo import ops
    if args.attn_bwd_impl:
        import functools
        ops.attention_backward = functools.partial(ops.attention_backward, impl=args.attn_bwd_impl)
    from hriemo.train import Trainer
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    model = FusionWithEmotionDecoder(dropout=args.dropout).to(dev).train()
    trainer = Trainer(model, graph=args.graph)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    B, d, n_e = args.batch, 768, 4
    h_a = torch.randn(B, args.T_a, d, device=dev, generator=g)
    h_t = torch.randn(B, args.T_t, d, device=dev, generator=g)
    labels = torch.eye(n_e, device=dev)[torch.randint(0, n_e, (B,), device=dev, generator=g)]
    m_a = m_t = None
    if args.ragged:
        la = torch.randint(args.T_a // 2, args.T_a + 1, (B,), device=dev, generator=g)
        lt = torch.randint(args.T_t // 2, args.T_t + 1, (B,), device=dev, generator=g)
        m_a = torch.arange(args.T_a, device=dev)[None, :] >= la[:, None]
        m_t = torch.arange(args.T_t, device=dev)[None, :] >= lt[:, None]
    for _ in range(args.warmup):
        trainer.step(h_a, h_t, m_a, m_t, labels)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ops.PROFILE = None if args.graph else []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        info = trainer.step(h_a, h_t, m_a, m_t, labels)
    t1.record()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE or [], None
    ms = torch.tensor([t0.elapsed_time(t1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    kinds = {}
    for kind, work, a, b in prof:
        k = kinds.setdefault(kind, dict(ms=0.0, flops=0.0, launches=0))
        k["ms"] += a.elapsed_time(b)
        k["flops"] += work
        k["launches"] += 1
    breakdown = {k: dict(ms_per_step=v["ms"] / args.steps, share_of_step=v["ms"] / args.steps / ms,
                         tflops=v["flops"] / max(v["ms"], 1e-9) / 1e9, launches_per_step=v["launches"] // args.steps)
                 for k, v in kinds.items()}
    if rank == 0:
        print(json.dumps(dict(metric="training-step utterances/sec", value=B * world / ms * 1e3, unit="utterances/s",
                              n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms, dtype="bf16",
                              data="synthetic", loss=info["loss"].item(), grad_norm=info["grad_norm"].item(),
                              config=dict(workload=f"FusionWithEmotionDecoder BCE training step, B={B}/GPU, T_a={args.T_a}, "
                                                   f"T_t={args.T_t}, d=768, H=8, N_e=4, 2+2 layers, AdamW, clip 5.0, dropout {args.dropout}",
                                          parameters=trainer.numel, cuda_graph=bool(args.graph), ragged=bool(args.ragged), exchange="one all-reduce (AVG) of the fp32 gradient arena"),
                              breakdown=breakdown, peak_mem_gb=torch.cuda.max_memory_allocated() / 2**30)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
