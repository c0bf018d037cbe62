#!/usr/bin/env python3
"""Benchmark of the HRI-EMO fusion-and-decode forward path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (reference algorithm)
    python bench.py --impl torch_gpu --gpus N --steps K ...  # stock PyTorch on the same GPUs (the reference's GPU path)

A "step" is one FusionWithEmotionDecoder forward over one batch of synthetic
WavLM/BERT-shaped fp32 features with random-init weights (seed 1234).  `value` is
whole-job utterances/s with the inputs already resident in HBM; `e2e` is the same
metric through hriemo.pipeline.forward_from_host with pinned HOST buffers (H2D of the
features and D2H of logits/beta/z inside the timed region).  With N > 1 (torchrun) every
rank owns its own shard of utterances (weak scaling) and the only collective is one
all_gather of logits+beta per step, inside the timed region.  One JSON line on rank 0.

Extra keys of the line (all measured after the contract's own timed region, none inside it):
  parity              (N = 1) the oracle's outputs of the cpu_baseline leg are KEPT and compared with the GPU forward on
                      the same >= 1024 utterances (half of them with ragged masks): logits max-abs, threshold decisions,
                      argmax, beta decisions, each with its denominator and what was excluded
  gpu_eager_baseline  (N = 1) stock PyTorch eager on the same GPU, same batch: fp32 (TF32 off), TF32, bf16 autocast --
                      baseline/stock_torch.py, the reference's module graph (it ships no kernel of its own)
  ragged              the same workload with ragged True = PAD masks (lengths uniform in [T/2, T]): padded and bucketed
  train_step          BASELINE config 5 at EVERY N: FusionWithEmotionDecoder BCE training step (hriemo.train.Trainer,
                      B = 512 per GPU, CUDA-graph replay, one NCCL all-reduce of the gradient arena inside the timed
                      region), with the all-reduce's own time
--no-train / --no-cpu / --no-torch / --no-ragged skip them; a failure in one is reported inside its key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "hri-emo_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # north_star target point
    "ns": dict(T_a=500, T_t=64, B=4096, desc="FusionWithEmotionDecoder fwd, B=4096/GPU, T_a=500, T_t=64, d=768, H=8, N_e=4, 2+2 layers"),
    # BASELINE.json configs[1]
    "cfg2": dict(T_a=300, T_t=50, B=4096, desc="FusionWithEmotionDecoder fwd, B=4096/GPU, T_a=300, T_t=50, d=768, H=8, N_e=4, 2+2 layers"),
    # BASELINE.json config 4's longest sequences
    "long": dict(T_a=1000, T_t=64, B=2048, desc="FusionWithEmotionDecoder fwd, B=2048/GPU, T_a=1000, T_t=64, d=768, H=8, N_e=4, 2+2 layers"),
}
# BASELINE.json config 3: MOSEI wrapper (d_audio 74, d_text 300 -> d_model 256, 4 heads, 6 emotion queries)
WORKLOADS["cfg3"] = dict(T_a=300, T_t=128, B=8192, mosei=True,
                         desc="MoseiFusionWithEmotionDecoder fwd, B=8192/GPU, T_a=300, T_t=128, d_audio=74, d_text=300, d=256, H=4, N_e=6, 2+2 layers")
METRIC = "seq-level utterances/sec"
UNIT = "utterances/s"


def arm_config(wl, world):
    """The `config` object, identical in every arm (ours, reference, torch_gpu) for the same workload and N."""
    return {"workload": wl["desc"], "global_batch": world * wl["B"], "parallelism": f"batch-sharded x{world}",
            "inputs": "fp32 features, no masks (the masked variant of the same batch is the `ragged` key of our arm)",
            "l2": "per-step inputs (7.1 GB at ns) >> 126 MB L2"}


def lo_shape_bytes(model, B):
    """bytes read back per e2e step: logits [B,N_e] + beta [B,1] + z [B,N_e,d], fp32"""
    dec = model.backbone.emotion_decoder if hasattr(model, "backbone") else model.emotion_decoder
    return B * (dec.num_emotions + 1 + dec.num_emotions * dec.d_model) * 4


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


def flops_per_utt(T_a, T_t, d=768, n_e=4, L_f=2, L_d=2, h_beta=256, ffn_dec=2048, d_a=0, d_t=0):
    """SURVEY sec. 8(d) closed form (matmul FLOPs, 2*MAC); d_a / d_t > 0 adds the MOSEI input projections."""
    L = T_t
    f = L_f * (32 * d * d * (T_a + T_t) + 4 * d * (T_a + T_t) ** 2) + 2 * (T_a * d_a + T_t * d_t) * d
    f += 10 * d * h_beta
    f += L_d * (12 * n_e * d * d + 4 * L * d * d + 4 * n_e * n_e * d + 4 * n_e * L * d + 4 * n_e * d * ffn_dec)
    return f + 2 * n_e * d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None   # wall-clock bounds of the timed region (mark_start / mark_end)

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")] + [time.time()])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        # the sampler is started before the warm-up (nvidia-smi takes a few hundred ms to come up);
        # keep the rows read inside the timed region, or failing that the ones closest after its start
        rows = [r for r in self.rows if len(r) >= 8]
        if self.t0 is not None:
            inside = [r for r in rows if self.t0 <= r[-1] <= (self.t1 or 1e30) + 0.25]
            rows = inside or [r for r in rows if r[-1] >= self.t0 - 0.25][:3] or rows[-2:]
        self.rows = rows
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.rows[0][1])), "samples": len(sm),
                "power_w_max": max(float(r[2]) for r in self.rows if len(r) >= 7), "reasons": reasons}


# ------------------------------------------------------------------------- CPU arm
def cpu_forward_timer(T_a, T_t, sample_B, steps, warmup, mosei=False):
    """The reference algorithm on the host cores: the oracle port (torch CPU fp32, all threads).
    The reference itself is a Python package that cannot travel to the GPU box."""
    import hriemo_oracle as O
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder
    from models.mosei_fusion_with_emotion_decoder import MoseiFusionWithEmotionDecoder

    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(1234)
    m = MoseiFusionWithEmotionDecoder(74, 300, d_model=256, num_emotions=6, n_heads=4) if mosei else FusionWithEmotionDecoder()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    d_a, d_t = (74, 300) if mosei else (768, 768)
    g = torch.Generator().manual_seed(1234)
    h_a, h_t = torch.randn(sample_B, T_a, d_a, generator=g), torch.randn(sample_B, T_t, d_t, generator=g)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            if mosei:
                O.mosei_fusion_with_emotion_decoder(sd, h_a, h_t, None, None, n_heads=4)
            else:
                O.fusion_with_emotion_decoder(sd, h_a, h_t, None, None, n_heads=8)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


# ------------------------------------------------------------------------- parity + CPU baseline (one leg)
def parity_and_cpu_leg(model, dev, T_a, T_t, d_a, d_t, n_heads, mosei, n_total, chunk):
    """The oracle port on the host cores over `n_total` utterances of the workload (chunks of `chunk`; the first half
    without masks, the second half with ragged True = PAD masks), its outputs KEPT and compared with the GPU forward
    on the same inputs and weights.  -> (parity dict, cpu_baseline dict).

    cpu_baseline.value comes from the unmasked chunks after the first (the warm-up); parity follows SURVEY App. D-3/D-4:
    multi-label decision = logit > 0 (train_fusion_seq_level_decoder.py:319), label index = argmax (:314),
    dominance = beta > 0.5; decisions whose reference value lies within `tol` of the threshold are counted and
    reported both ways (with and without them)."""
    import hriemo_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(4321)
    n_chunks = max(2, n_total // chunk)
    half = n_chunks // 2
    fwd = O.mosei_fusion_with_emotion_decoder if mosei else O.fusion_with_emotion_decoder
    from hriemo import precise
    ref_lo, ref_be, got_lo, got_be, times = [], [], [], [], []
    x3_lo, x3_be = [], []
    with torch.no_grad():
        for c in range(n_chunks):
            ragged = c >= half
            h_a, h_t = torch.randn(chunk, T_a, d_a, generator=g), torch.randn(chunk, T_t, d_t, generator=g)
            m_a = O.ragged_masks(chunk, T_a, g) if ragged else None
            m_t = O.ragged_masks(chunk, T_t, g) if ragged else None
            t0 = time.perf_counter()
            lo, be, _ = fwd(sd, h_a, h_t, m_a, m_t, n_heads=n_heads)
            dt = time.perf_counter() - t0
            if not ragged and (c >= 1 or half == 1):
                times.append(dt)
            glo, gbe, _ = model(h_a.to(dev), h_t.to(dev), None if m_a is None else m_a.to(dev), None if m_t is None else m_t.to(dev))
            ref_lo.append(lo.double()); ref_be.append(be.double())
            got_lo.append(glo.double().cpu()); got_be.append(gbe.double().cpu())
            if precise.get_mode() == "bf16":   # the tf32-class mode on the same inputs (north_star: logits <= 1e-4)
                with precise.mode("tf32x3"):
                    plo, pbe, _ = model(h_a.to(dev), h_t.to(dev), None if m_a is None else m_a.to(dev), None if m_t is None else m_t.to(dev))
                x3_lo.append(plo.double().cpu()); x3_be.append(pbe.double().cpu())
    R, G = torch.cat(ref_lo), torch.cat(got_lo)
    Rb, Gb = torch.cat(ref_be).view(-1), torch.cat(got_be).view(-1)
    n = R.shape[0]
    is_x3 = precise.get_mode() == "tf32x3"
    tol = 1e-4 if is_x3 else 1e-2               # the logits bar itself: a reference logit closer to 0 than this is undecidable at that precision
    near = R.abs() < tol
    same = (R > 0) == (G > 0)
    top2 = R.topk(2, dim=1).values
    amb = (top2[:, 0] - top2[:, 1]) < tol
    am_same = R.argmax(1) == G.argmax(1)
    b_near = (Rb - 0.5).abs() < 1e-4            # the beta bar
    b_same = (Rb > 0.5) == (Gb > 0.5)
    parity = {
        "n": n, "n_ragged": (n_chunks - half) * chunk, "decisions": int(R.numel()),
        "reference": "oracle port (fp32, CPU) of the reference forward, same weights (state_dict of the GPU model), same inputs",
        "logits_max_abs": float((R - G).abs().max()), "logits_bar": tol, "precision": precise.get_mode(),
        "beta_max_abs": float((Rb - Gb).abs().max()), "beta_bar": 1e-4,
        "thr_agree": float(same[~near].double().mean()) if (~near).any() else None, "thr_excluded": int(near.sum()),
        "thr_excluded_rule": f"|reference logit| < {tol}", "thr_agree_all": float(same.double().mean()),
        "thr_disagreements": int((~same).sum()), "thr_disagreements_outside_tol": int((~same & ~near).sum()),
        "argmax_agree": float(am_same[~amb].double().mean()) if (~amb).any() else None, "argmax_excluded": int(amb.sum()),
        "argmax_excluded_rule": f"reference top-2 gap < {tol}", "argmax_agree_all": float(am_same.double().mean()),
        "beta_gt_half_agree": float(b_same[~b_near].double().mean()) if (~b_near).any() else None,
        "beta_gt_half_excluded": int(b_near.sum()), "beta_gt_half_agree_all": float(b_same.double().mean()),
        "beta_batch_argmax_equal": bool(Rb.argmax() == Gb.argmax()),
        "beta_batch_argmax_note": "beta spans ~1e-3 over a batch at random init (SURVEY App. D-3): the batch argmax is "
                                  "reported, the per-sample dominance decision beta > 0.5 is the one held to 100 %",
        "beta_range_reference": [float(Rb.min()), float(Rb.max())],
    }
    if x3_lo:
        P3, Pb3 = torch.cat(x3_lo), torch.cat(x3_be).view(-1)
        near3 = R.abs() < 1e-4
        parity["tf32x3"] = {
            "mode": "hriemo.precise.mode('tf32x3'): fp32 activations, Linear layers as split-bf16 (hi + lo) tcgen05 GEMMs, fp32 attention / LayerNorm / gate",
            "n": n, "logits_max_abs": float((R - P3).abs().max()), "logits_bar": 1e-4,
            "reference_note": "the reference here is the fp32 oracle port (itself ~5e-7 from float64)",
            "beta_max_abs": float((Rb - Pb3).abs().max()),
            "thr_agree_all": float(((R > 0) == (P3 > 0)).double().mean()), "thr_within_1e-4_of_zero": int(near3.sum()),
            "argmax_agree_all": float((R.argmax(1) == P3.argmax(1)).double().mean()),
            "beta_gt_half_agree_all": float(((Rb > 0.5) == (Pb3 > 0.5)).double().mean())}
    n_timed = chunk * len(times)
    cpu = {"value": n_timed / sum(times), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
           "sample": f"{n_timed} utterances in {len(times)} timed chunks of {chunk} (+1 warm-up chunk) of the same workload "
                     f"(T_a={T_a}, T_t={T_t}), fp32, all host threads, oracle port of the reference forward; its outputs "
                     f"are the parity reference"}
    return parity, cpu


# ------------------------------------------------------------------------- stock PyTorch on the same GPU
def stock_torch_leg(model, dev, h_a, h_t, n_heads, steps, slab, modes=None, world=1, dist=None):
    """Stock PyTorch eager (the reference's GPU path: cuBLASLt + SDPA + ATen LayerNorm / elementwise through
    nn.MultiheadAttention, baseline/stock_torch.py) on the same device, weights and batch, in slabs of `slab`
    utterances (the fp32 intermediates of a 4096-utterance batch do not fit; larger slabs are no faster in eager
    mode), CUDA-event timed after one warm-up pass, max over ranks.  -> {mode: {...}}."""
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    from baseline.stock_torch import MODES, StockFusion, run_mode

    stock = StockFusion({k: v.detach() for k, v in model.state_dict().items()}, n_heads).eval().to(dev)
    B = h_a.shape[0]
    out = {}
    for mode in (modes or MODES):
        n_steps = max(1, steps if mode != "fp32" else min(steps, 1))
        def one_pass():
            for s in range(0, B, slab):
                run_mode(stock, mode, h_a[s:s + slab], h_t[s:s + slab])
        try:
            run_mode(stock, mode, h_a[:slab], h_t[:slab])   # warm-up: cuBLAS / SDPA heuristics, autocast caches
            if mode != "fp32":
                one_pass()
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n_steps):
                one_pass()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ms = float(ms.item())
            out[mode] = {"value": world * B * n_steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / n_steps, "steps": n_steps,
                         "slab": slab}
        except Exception as e:  # noqa: BLE001
            out[mode] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    del stock
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------- ragged variant of the workload
def ragged_leg(model, dev, h_a, h_t, steps, world=1, dist=None):
    """The same batch with the masks every real batch of the reference carries (collate zero-pads to the batch
    maximum, True = PAD; train_fusion_seq_level_decoder.py:191-232): valid lengths uniform in [T/2, T].
    (a) the padded forward with masks (trailing all-PAD key tiles skipped), (b) pipeline.forward_bucketed
    (utterances sorted by length, slabs trimmed to their own maxima; same outputs)."""
    from hriemo import pipeline

    B, T_a, T_t = h_a.shape[0], h_a.shape[1], h_t.shape[1]
    g = torch.Generator(device=dev).manual_seed(77)
    la = torch.randint(T_a // 2, T_a + 1, (B,), device=dev, generator=g)
    lt = torch.randint(T_t // 2, T_t + 1, (B,), device=dev, generator=g)
    m_a = torch.arange(T_a, device=dev)[None, :] >= la[:, None]
    m_t = torch.arange(T_t, device=dev)[None, :] >= lt[:, None]

    def timed(fn):
        fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return out, float(ms.item()) / steps

    ref, ms_pad = timed(lambda: model(h_a, h_t, m_a, m_t)[:3])
    out, ms_bkt = timed(lambda: pipeline.forward_bucketed(model, h_a, h_t, m_a, m_t))
    rows_valid = int(la.sum() + lt.sum())
    return {"masks": "trailing PAD, valid lengths uniform in [T/2, T]", "steps": steps,
            "padded": {"value": world * B / ms_pad * 1e3, "unit": UNIT, "ms_per_step": ms_pad},
            "bucketed": {"value": world * B / ms_bkt * 1e3, "unit": UNIT, "ms_per_step": ms_bkt,
                         "max_abs_diff_vs_padded": [float((a - b).abs().max()) for a, b in zip(out, ref)]},
            "valid_row_fraction": rows_valid / (B * (T_a + T_t))}


# ------------------------------------------------------------------------- training step (config 5), every N
def train_leg(dev, world, rank, dist, batch=512, steps=5, warmup=4, T_a=500, T_t=64):
    """BASELINE config 5: FusionWithEmotionDecoder BCE training step through hriemo.train.Trainer (forward with tapes,
    backward, ONE all-reduce of the fp32 gradient arena over NCCL, global-norm clip, AdamW), B = `batch` per GPU,
    forward + backward replayed from a CUDA graph, synthetic features resident in HBM.  CUDA events around the
    timed steps (max over ranks) and around every all-reduce."""
    from hriemo.train import Trainer
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    torch.manual_seed(0)
    model = FusionWithEmotionDecoder(dropout=0.0).to(dev)
    trainer = Trainer(model, graph=True)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    h_a = torch.randn(batch, T_a, 768, device=dev, generator=g)
    h_t = torch.randn(batch, T_t, 768, device=dev, generator=g)
    labels = torch.eye(4, device=dev)[torch.randint(0, 4, (batch,), device=dev, generator=g)]
    for _ in range(warmup):
        trainer.step(h_a, h_t, None, None, labels)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)
    trainer.allreduce_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        info = trainer.step(h_a, h_t, None, None, labels)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    ar = torch.tensor([sum(a.elapsed_time(b) for a, b in trainer.allreduce_events) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ar, op=dist.ReduceOp.MAX)
    ms, ar = float(ms.item()), float(ar.item())
    res = {"metric": "training-step utterances/sec", "value": batch * world / ms * 1e3, "unit": UNIT, "n_gpus": world,
           "ms_per_step": ms, "steps": steps, "warmup": warmup, "dtype": "bf16", "loss": float(info["loss"].item()),
           "allreduce_ms": ar, "allreduce_bytes": trainer.numel * 4, "allreduce_share_of_step": ar / ms,
           "allreduce_overlapped": bool(trainer._exchanging()),
           "config": {"workload": f"FusionWithEmotionDecoder BCE training step, B={batch}/GPU, T_a={T_a}, T_t={T_t}, d=768, "
                                  "H=8, N_e=4, 2+2 layers, AdamW, clip 5.0, dropout 0", "cuda_graph": True,
                      "exchange": ("NCCL all-reduce (AVG) of the fp32 gradient arena in two parts inside the timed region: decoder, gate and "
                                   "encoder layers >= 1 (65 % of the bytes) on NCCL's stream under the backward of encoder layer 0, layer 0's "
                                   "share after it; allreduce_ms is what the step still waits for (events around the second part and the join)")
                                  if trainer._exchanging() else
                                  "one NCCL all-reduce (AVG) of the fp32 gradient arena per step, after the backward, inside the timed region"},
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
    del trainer, model
    torch.cuda.empty_cache()
    return res


def stock_cpu_timer(T_a, T_t, sample_B, steps, warmup, mosei=False):
    """The reference's own module graph on the host cores: baseline/stock_torch.py rebuilds it from torch.nn modules
    (nn.MultiheadAttention incl. its eval()/no_grad fast path, nn.LayerNorm, nn.Linear -- the very ATen CPU kernels the
    reference dispatches; checked against the reference's golden outputs in tests/test_stock_torch_cpu.py), fp32,
    eval() + no_grad, all threads.  The oracle port (explicit matmul / softmax) is ~20 % slower than this."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from baseline.stock_torch import StockFusion, run_mode
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder
    from models.mosei_fusion_with_emotion_decoder import MoseiFusionWithEmotionDecoder

    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(1234)
    m = MoseiFusionWithEmotionDecoder(74, 300, d_model=256, num_emotions=6, n_heads=4) if mosei else FusionWithEmotionDecoder()
    stock = StockFusion(m.state_dict(), 4 if mosei else 8).eval()
    d_a, d_t = (74, 300) if mosei else (768, 768)
    g = torch.Generator().manual_seed(1234)
    h_a, h_t = torch.randn(sample_B, T_a, d_a, generator=g), torch.randn(sample_B, T_t, d_t, generator=g)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        run_mode(stock, "fp32", h_a, h_t)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_B = args.cpu_sample or 96
    mosei = bool(wl.get("mosei"))
    times, cores = stock_cpu_timer(wl["T_a"], wl["T_t"], sample_B, args.steps, min(args.warmup, 1), mosei)
    total = sum(times)
    value = sample_B * len(times) / total
    o_times, _ = cpu_forward_timer(wl["T_a"], wl["T_t"], sample_B, 1, 0, mosei)
    sample = (f"{sample_B} utterances per step of the same workload (T_a={wl['T_a']}, T_t={wl['T_t']}), fp32, eval() + no_grad, "
              "the reference's module graph rebuilt from torch.nn modules (baseline/stock_torch.py: same ATen CPU kernels as the "
              "reference, incl. the nn.MultiheadAttention fast path; the reference package itself cannot travel to the GPU box)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": arm_config(wl, max(1, args.gpus)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "oracle_port_value": sample_B * len(o_times) / sum(o_times)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


class Deadline:
    """Supplementary legs (training step over NCCL, ...) run after the contract's numbers are final.  If one hangs
    (a rank stuck in a collective), the line assembled so far is printed with the leg marked as failed and the
    process exits 0: a supplementary record must never cost the run its contract line."""

    def __init__(self, seconds, emit, key):
        self.seconds, self.emit, self.key, self.done = seconds, emit, key, threading.Event()

    def __enter__(self):
        def watch():
            if not self.done.wait(self.seconds):
                self.emit({self.key: {"error": f"no result after {self.seconds} s; leg abandoned"}})
                os._exit(0)
        threading.Thread(target=watch, daemon=True).start()
        return self

    def __exit__(self, *a):
        self.done.set()


def build_model_and_inputs(wl, dev, rank):
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    torch.manual_seed(1234)
    mosei = bool(wl.get("mosei"))
    if mosei:
        from models.mosei_fusion_with_emotion_decoder import MoseiFusionWithEmotionDecoder
        model = MoseiFusionWithEmotionDecoder(74, 300, d_model=256, num_emotions=6, n_heads=4).eval().to(dev)
        d_a, d_t, n_heads = 74, 300, 4
    else:
        model = FusionWithEmotionDecoder().eval().to(dev)
        d_a, d_t, n_heads = 768, 768, 8
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    h_a = torch.randn(wl["B"], wl["T_a"], d_a, generator=g, device=dev)
    h_t = torch.randn(wl["B"], wl["T_t"], d_t, generator=g, device=dev)
    return model, h_a, h_t, d_a, d_t, n_heads, mosei


def pinned_inputs(B, T_a, T_t, d_a, d_t, world, rank, dtype=torch.float32):
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    bytes_per_utt = (T_a * d_a + T_t * d_t) * 4
    Be = B
    while Be > 256 and Be * bytes_per_utt * world * 1.5 > avail * 0.5:
        Be //= 2
    ha_h = torch.empty((Be, T_a, d_a), dtype=torch.float32).pin_memory()
    ht_h = torch.empty((Be, T_t, d_t), dtype=torch.float32).pin_memory()
    ha_h.normal_(generator=torch.Generator().manual_seed(99 + rank))
    ht_h.normal_(generator=torch.Generator().manual_seed(199 + rank))
    return Be, ha_h, ht_h, bytes_per_utt


def run_torch_gpu_arm(args, wl):
    """`--impl torch_gpu`: the competitor that matters -- stock PyTorch eager on the same B200s, the reference's own GPU
    path (baseline/stock_torch.py).  Same metric, workload, batch, seeds and timing rules as our arm; `value` is the
    bf16-autocast mode (what scripts/infer/README.md:55 recommends), the other precisions are listed beside it."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl torch_gpu: no CUDA device")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    model, h_a, h_t, d_a, d_t, n_heads, mosei = build_model_and_inputs(wl, dev, rank)
    B, T_a, T_t = wl["B"], wl["T_a"], wl["T_t"]
    slab = args.torch_slab
    with ClockSampler(local) as clk:
        clk.mark_start()
        modes = stock_torch_leg(model, dev, h_a, h_t, n_heads, args.steps, slab, world=world, dist=dist)
        clk.mark_end()
    # end to end the way the reference's loops do it (mosei_eval_infer.py:190-240): pinned batch .to(device), forward
    # under autocast, results .cpu()
    from baseline.stock_torch import StockFusion, run_mode
    e2e = None
    if not args.no_e2e:
        del h_a, h_t
        torch.cuda.empty_cache()
        Be, ha_h, ht_h, bytes_per_utt = pinned_inputs(B, T_a, T_t, d_a, d_t, world, rank)
        stock = StockFusion({k: v.detach() for k, v in model.state_dict().items()}, n_heads).eval().to(dev)

        def e2e_pass():
            outs = []
            for s in range(0, Be, slab):
                a, t = ha_h[s:s + slab].to(dev, non_blocking=True), ht_h[s:s + slab].to(dev, non_blocking=True)
                lo, be, z = run_mode(stock, "bf16_autocast", a, t)
                outs.append((lo.float().cpu(), be.float().cpu(), z.float().cpu()))
            return outs

        e2e_pass()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n_e2e = 2
        for _ in range(n_e2e):
            e2e_pass()
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * Be * n_e2e / float(dt.item()), "unit": UNIT, "batch_per_gpu": Be, "steps": n_e2e,
               "h2d_bytes_per_step": Be * bytes_per_utt, "d2h_bytes_per_step": int(lo_shape_bytes(model, Be)),
               "api": "pinned fp32 batch .to(device, non_blocking) per slab -> stock modules under torch.autocast(bfloat16) -> .cpu()"}
    if rank == 0:
        best = modes.get("bf16_autocast", {})
        line = {"impl": "torch_gpu", "metric": METRIC, "value": best.get("value"), "unit": UNIT, "n_gpus": world,
                "steps": best.get("steps"), "warmup": 1, "ms_per_step": best.get("ms_per_step"), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (torch.autocast)", "data": "synthetic",
                "config": arm_config(wl, world),
                "engine": {"slab": slab, "what": "stock PyTorch eager: nn.MultiheadAttention (cuBLASLt + SDPA / native MHA fast path), "
                                                 "nn.LayerNorm, nn.Linear -- the reference's module graph (baseline/stock_torch.py)"},
                "modes": modes, "e2e": e2e, "gpu_launches": 0, "clocks": clk.summary(),
                "torch": torch.__version__}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--workload", default="ns", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU (default: workload's)")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="utterances of the CPU leg (default: 1024 for the in-line parity + cpu_baseline leg = ~30 s of CPU "
                         "work on 16 cores; 96 per step for --impl reference)")
    ap.add_argument("--torch-slab", type=int, default=512, help="utterances per stock-PyTorch forward (torch_gpu arm / gpu_eager_baseline)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32x3"],
                    help="tf32x3: the tf32-class mode (hriemo/precise.py; logits <= 1e-4) as the measured forward; "
                         "e2e / ragged / training legs are bf16-path features and are skipped")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU leg (cpu_baseline and parity)")
    ap.add_argument("--no-torch", action="store_true", help="skip gpu_eager_baseline (stock PyTorch on the same GPU)")
    ap.add_argument("--no-ragged", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the supplementary training-step record (config 5)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["desc"] = wl["desc"].replace(f"B={wl['B']}/GPU", f"B={args.batch}/GPU")
        wl["B"] = args.batch
    if args.impl == "reference":
        return run_reference_arm(args, wl)
    if args.impl == "torch_gpu":
        return run_torch_gpu_arm(args, wl)

    from hriemo import lib, ops, pipeline, precise

    if args.precision == "tf32x3":
        precise.set_mode("tf32x3")
        args.no_e2e = args.no_ragged = args.no_train = True
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the forward path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    B, T_a, T_t = wl["B"], wl["T_a"], wl["T_t"]
    model, h_a, h_t, d_a, d_t, n_heads, mosei = build_model_and_inputs(wl, dev, rank)

    def step():
        logits, beta, _ = model(h_a, h_t)
        if world > 1:
            logits, beta = pipeline.gather_outputs(logits, beta)
        return logits, beta

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    with ClockSampler(local) as clk:
        for _ in range(max(args.warmup, 3)):
            step()
        sync_all()
        ops.PROFILE = []
        ops.PROFILE_BYTES = []
        ops.PROFILE_ATTN = []
        n0 = lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk.mark_start()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        sync_all()
        clk.mark_end()
    launches = lib.launch_count() - n0
    prof, ops.PROFILE = ops.PROFILE, None
    prof_bytes, ops.PROFILE_BYTES = ops.PROFILE_BYTES, None
    prof_attn, ops.PROFILE_ATTN = ops.PROFILE_ATTN, None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * args.steps / (ms_total * 1e-3)

    # per-kernel roofline from the events recorded around each launch in the timed region
    def agg(kind):
        rows = [(w, a.elapsed_time(b)) for (k, w, a, b) in prof if k == kind]
        return sum(r[0] for r in rows), sum(r[1] for r in rows), len(rows)

    # GEMM launches are tagged by the reference call group they replace: "attn_proj" (MHA in/out
    # projections), "ffn", "gemm" (everything else: decoder, MOSEI input projections)
    parts = [agg(k) for k in ("gemm", "attn_proj", "ffn", "precise")]
    g_fl, g_ms, g_n = (sum(p[i] for p in parts) for i in range(3))
    a_fl, a_ms, a_n = agg("attention")
    p_fl, p_ms, p_n = agg("attn_proj")
    gemm_tf = g_fl / (g_ms * 1e-3) / 1e12 if g_ms else 0.0
    attn_tf = a_fl / (a_ms * 1e-3) / 1e12 if a_ms else 0.0
    # SURVEY 8(d) "attention-kernel % of roofline": in-proj + QK^T + PV + out-proj FLOPs over their time
    mha_tf = (a_fl + p_fl) / ((a_ms + p_ms) * 1e-3) / 1e12 if (a_ms + p_ms) else 0.0
    # DRAM traffic per launch of the dominant kernel comes from the committed ncu capture of this same
    # command (profiles/r0N_traffic.json, written by tools/traffic_from_ncu.py); None if absent
    traffic, traffic_note = None, None
    for tname in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath) and args.workload == "ns" and B == WORKLOADS["ns"]["B"]:
            tj = json.load(open(tpath))
            traffic, traffic_note = tj["gemm"]["dram_bytes_per_launch"], tj["gemm"]["note"] + f" [{tname}]"
            break
    g_bytes = sum(w for (k, w, a, b) in prof_bytes) / max(g_n, 1) if prof_bytes else None
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05 GEMM, all projections + FFN)",
                "achieved": gemm_tf, "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": gemm_tf / peaks["tf_sust"],
                "traffic": traffic, "traffic_note": traffic_note, "algorithmic_bytes_per_launch": g_bytes,
                "flops_per_launch": g_fl / max(g_n, 1), "avg_launch_ms": g_ms / max(g_n, 1),
                "launches": g_n, "share_of_step": g_ms / ms_total, "peak_source": peaks["src"] + ", sustained"}
    attention_roofline = {"bound": "tensor", "kernel": "attention_fwd kernels (tcgen05 QK^T/PV + online softmax)",
                          "achieved": attn_tf, "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": attn_tf / peaks["tf_sust"],
                          "launches": a_n, "share_of_step": a_ms / ms_total,
                          "mha_incl_projections": {"achieved": mha_tf, "frac": mha_tf / peaks["tf_sust"], "unit": "TFLOP/s",
                                                   "launches": a_n + p_n, "share_of_step": (a_ms + p_ms) / ms_total,
                                                   "definition": "SURVEY 8(d): (in-proj + QK^T + PV + out-proj FLOPs) / their kernel time"}}
    # each attention shape against ITS roofline: the time the launch would take at the sustained tensor peak and at the
    # measured HBM bandwidth, whichever is larger (the cross shapes of a layer move 3.5 GB for 0.2 TFLOP: HBM work)
    by_shape, ideal_ms = {}, 0.0
    for (tq, tk, fl, by, a, b) in prof_attn or []:
        r = by_shape.setdefault((tq, tk), dict(Tq=tq, Tk=tk, launches=0, ms=0.0, flops=0.0, bytes=0.0))
        r["launches"] += 1
        r["ms"] += a.elapsed_time(b)
        r["flops"] += fl
        r["bytes"] += by
    shapes = []
    for r in by_shape.values():
        t_tensor, t_hbm = r["flops"] / (peaks["tf_sust"] * 1e12) * 1e3, r["bytes"] / (peaks["hbm"] * 1e9) * 1e3
        ideal_ms += max(t_tensor, t_hbm)
        shapes.append({"Tq": r["Tq"], "Tk": r["Tk"], "launches": r["launches"], "avg_launch_ms": r["ms"] / r["launches"],
                       "tflops": r["flops"] / (r["ms"] * 1e-3) / 1e12, "algorithmic_gbs": r["bytes"] / (r["ms"] * 1e-3) / 1e9,
                       "bound": "tensor" if t_tensor >= t_hbm else "hbm",
                       "frac_of_its_roofline": max(t_tensor, t_hbm) / r["ms"]})
    if shapes:
        attention_roofline["by_shape"] = sorted(shapes, key=lambda x: -x["avg_launch_ms"] * x["launches"])
        attention_roofline["frac_of_shape_rooflines"] = ideal_ms / a_ms
        attention_roofline["by_shape_note"] = ("algorithmic bytes = Q + K + V + O in bf16; frac_of_its_roofline = max(FLOPs / sustained "
                                               "tensor peak, bytes / measured HBM bandwidth) / measured time")
    fpu = (flops_per_utt(T_a, T_t, d=256, n_e=6, h_beta=128, d_a=74, d_t=300) if mosei else flops_per_utt(T_a, T_t))
    path_tf = fpu * value / world / 1e12

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "tf32x3 (fp32 activations; bf16 hi + lo split operands on the tcgen05 GEMM, fp32 accumulate)",
            "data": "synthetic",
            "config": arm_config(wl, world),
            "roofline": roofline, "attention_roofline": attention_roofline,
            "path": {"flops_per_utt": fpu, "achieved_tflops_per_gpu": path_tf, "frac_of_tensor_peak": path_tf / peaks["tf_sust"]},
            "cpu_baseline": None, "e2e": None, "gpu_launches": launches, "clocks": clk.summary()}
    printed = threading.Event()

    def emit(extra=None):
        if rank == 0 and not printed.is_set():
            printed.set()
            out = dict(line)
            out.update(extra or {})
            print(json.dumps(out), flush=True)

    # ---- the same batch with ragged masks (padded and bucketed), still device-resident
    if not args.no_ragged and T_a > 1:
        try:
            line["ragged"] = ragged_leg(model, dev, h_a, h_t, 3, world, dist)
        except Exception as e:  # noqa: BLE001
            line["ragged"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- the tf32-class mode on the head of the same batch (its full line: --precision tf32x3)
    if rank == 0 and world == 1 and args.precision == "bf16" and not args.no_torch:
        try:
            nb = min(B, 1024)
            with precise.mode("tf32x3"):
                model(h_a[:nb], h_t[:nb])
                torch.cuda.synchronize(dev)
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                for _ in range(2):
                    model(h_a[:nb], h_t[:nb])
                p1.record()
                torch.cuda.synchronize(dev)
            line["tf32x3"] = {"value": 2 * nb / (p0.elapsed_time(p1) * 1e-3), "unit": UNIT, "batch": nb, "steps": 2,
                              "note": "hriemo.precise.mode('tf32x3'); parity of the mode: parity.tf32x3"}
        except Exception as e:  # noqa: BLE001
            line["tf32x3"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    # ---- stock PyTorch on the same GPU, same batch (N = 1: the competitor of the kernels; --impl torch_gpu runs it at any N)
    if rank == 0 and world == 1 and not args.no_torch:
        try:
            modes = stock_torch_leg(model, dev, h_a, h_t, n_heads, 3, args.torch_slab)
            best = modes.get("bf16_autocast", {}).get("value")
            line["gpu_eager_baseline"] = {"modes": modes, "unit": UNIT, "torch": torch.__version__,
                                          "engine": "stock PyTorch eager, the reference's module graph (baseline/stock_torch.py)",
                                          "ours_over_torch_bf16_autocast": value / best if best else None,
                                          "ours_over_torch_tf32": (value / modes["tf32"]["value"]) if modes.get("tf32", {}).get("value") else None}
        except Exception as e:  # noqa: BLE001
            line["gpu_eager_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- end to end from pinned host memory
    if not args.no_e2e:
        del h_a, h_t
        torch.cuda.empty_cache()
        Be, ha_h, ht_h, bytes_per_utt = pinned_inputs(B, T_a, T_t, d_a, d_t, world, rank)
        # the host-side bf16 pre-cast runs on the CPU cores: share them between the ranks of this box
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))

        def e2e_stream(n, a, t):
            # a stream of batches, two in flight: step i+1 is issued before step i's results are awaited, so its
            # first slab is copied while step i's last slabs compute; every step's H2D and D2H are in the region
            pend = None
            for _ in range(n):
                nxt = pipeline.forward_from_host(model, a, t, device=dev, slab=512, out_device="cpu", wait=False)
                if pend is not None:
                    pend.wait()
                pend = nxt
            return pend.wait()

        def e2e_sequential(n, a, t):
            for _ in range(n):
                pipeline.forward_from_host(model, a, t, device=dev, slab=512, out_device="cpu")   # results dropped at once

        def timed(fn, n, a, t):
            sync_all()
            pipeline.reset_stats()
            t0 = time.perf_counter()
            fn(n, a, t)
            torch.cuda.synchronize(dev)
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            st = dict(pipeline.STATS)
            return world * Be * n / float(dt.item()), st

        e2e_sequential(1, ha_h, ht_h)
        n_e2e, n_stream = 3, 8   # the stream starts cold after a synchronise: ~4 ms per step of ramp at 5 steps, ~2 at 8
        seq, _ = timed(e2e_sequential, n_e2e, ha_h, ht_h)
        e2e_stream(2, ha_h, ht_h)
        stream, st = timed(e2e_stream, n_stream, ha_h, ht_h)
        e2e = {"value": stream, "unit": UNIT, "batch_per_gpu": Be, "steps": n_stream,
               "sequential_value": seq, "sequential_steps": n_e2e,
               "h2d_bytes_per_step": st["h2d_bytes"] // max(st["calls"], 1),
               "host_cast_slabs_per_step": st["host_cast_slabs"] / max(st["calls"], 1), "slabs_per_step": st["slabs"] / max(st["calls"], 1),
               "host_threads_per_rank": torch.get_num_threads(),
               "host_bytes_read_per_step": Be * bytes_per_utt, "d2h_bytes_per_step": int(lo_shape_bytes(model, Be)),
               "api": "hriemo.pipeline.forward_from_host(model, pinned fp32 h_a, pinned fp32 h_t, wait=False) -> pinned host logits/beta/z; "
                      "a stream of batches with two in flight (step i+1 issued before step i's results are awaited; the pipeline "
                      "starts cold inside the timed region); the copy engine sends fp32 slabs from the front of the batch while the "
                      "host cores convert slabs to bf16 from the back (as many as they manage: host_cast_slabs_per_step; "
                      "h2d_bytes_per_step is counted from the tensors copied); sequential_value = one call at a time, each awaited"}
        # features ALREADY held as bf16 on the host (packed shards, or a loader that keeps bf16): half the bytes, no host
        # work; same outputs bit for bit (the first device op of the fp32 path is this same cast)
        if d_a % 8 == 0 and d_t % 8 == 0:
            try:
                ha16 = torch.empty(ha_h.shape, dtype=torch.bfloat16).pin_memory()
                ht16 = torch.empty(ht_h.shape, dtype=torch.bfloat16).pin_memory()
                ha16.copy_(ha_h)
                ht16.copy_(ht_h)
                e2e_stream(2, ha16, ht16)
                v16, st16 = timed(e2e_stream, n_stream, ha16, ht16)
                e2e["bf16_host"] = {"value": v16, "unit": UNIT, "steps": n_stream,
                                    "h2d_bytes_per_step": st16["h2d_bytes"] // max(st16["calls"], 1),
                                    "note": "same call with features already held as bf16 in pinned host memory"}
                del ha16, ht16
            except Exception as e:  # noqa: BLE001
                e2e["bf16_host"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        line["e2e"] = e2e
        del ha_h, ht_h
    else:
        del h_a, h_t
    torch.cuda.empty_cache()

    # ---- CPU leg: the oracle port on the host cores; its outputs are the parity reference for the GPU forward
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = args.cpu_sample or 1024
        try:
            parity, cpu = parity_and_cpu_leg(model, dev, T_a, T_t, d_a, d_t, n_heads, mosei, n_cpu, min(128, max(8, n_cpu // 8)))
            line["cpu_baseline"], line["parity"] = cpu, parity
        except Exception as e:  # noqa: BLE001
            line["parity"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- BASELINE config 5 at this N (every rank takes part: NCCL all-reduce of the gradients)
    if args.workload == "ns" and not args.batch and not args.no_train:
        with Deadline(240, emit, "train_step"):
            try:
                line["train_step"] = train_leg(dev, world, rank, dist)
            except Exception as e:  # noqa: BLE001
                line["train_step"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    emit()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
