#!/usr/bin/env python3
"""Benchmark of the HRI-EMO fusion-and-decode forward path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (reference algorithm)

A "step" is one FusionWithEmotionDecoder forward over one batch of synthetic
WavLM/BERT-shaped fp32 features with random-init weights (seed 1234).  `value` is
whole-job utterances/s with the inputs already resident in HBM; `e2e` is the same
metric through hriemo.pipeline.forward_from_host with pinned HOST buffers (H2D of the
features and D2H of logits/beta/z inside the timed region).  With N > 1 (torchrun) every
rank owns its own shard of utterances (weak scaling) and the only collective is one
all_gather of logits+beta per step, inside the timed region.  One JSON line on rank 0.

At N = 1 on the default workload the line carries one extra, clearly separate key, "train_step": the BASELINE config 5
record (FusionWithEmotionDecoder BCE training step, hriemo.train.Trainer, B = 512, CUDA-graph replay) measured by
tools/bench_train.py in a child process after the contract's own timed regions are over (--no-train skips it; a
failure there is reported inside the key and never fails the bench).
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


def train_step_leg(batch=512, timeout_s=180):
    """Supplementary record for BASELINE config 5 (not part of the contract keys): one B200, FusionWithEmotionDecoder BCE
    training step through hriemo.train.Trainer with CUDA-graph replay, measured by tools/bench_train.py in a CHILD process
    (its own CUDA context and timed region; a failure there cannot touch this run).  -> dict for the "train_step" key."""
    import subprocess
    root = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, os.path.join(root, "tools", "bench_train.py"), "--batch", str(batch), "--graph",
           "--steps", "5", "--warmup", "4"]
    try:
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env, cwd=root)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": (r.stderr.strip().splitlines() or ["no output"])[-1][:300]}
        d = json.loads(lines[-1])
        return {k: d[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "dtype", "config", "loss",
                                  "peak_mem_gb") if k in d}
    except Exception as e:  # noqa: BLE001  (supplementary leg: never fails the bench)
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_B = args.cpu_sample or 96
    times, cores = cpu_forward_timer(wl["T_a"], wl["T_t"], sample_B, args.steps, min(args.warmup, 1), bool(wl.get("mosei")))
    total = sum(times)
    value = sample_B * len(times) / total
    sample = f"{sample_B} utterances per step of the same workload (T_a={wl['T_a']}, T_t={wl['T_t']}), fp32, oracle port of the reference forward"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ns", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU (default: workload's)")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="utterances per CPU step (default: 256 for the in-line cpu_baseline = ~15 s of CPU work "
                         "over 1 warm-up + 2 timed passes; 96 per step for --impl reference)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the supplementary training-step record (config 5)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["B"] = args.batch
        wl["desc"] = wl["desc"].replace("B=4096/GPU", f"B={args.batch}/GPU")
    if args.impl == "reference":
        return run_reference_arm(args, wl)

    from hriemo import lib, ops, pipeline
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the forward path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    B, T_a, T_t = wl["B"], wl["T_a"], wl["T_t"]

    torch.manual_seed(1234)
    mosei = bool(wl.get("mosei"))
    if mosei:
        from models.mosei_fusion_with_emotion_decoder import MoseiFusionWithEmotionDecoder
        model = MoseiFusionWithEmotionDecoder(74, 300, d_model=256, num_emotions=6, n_heads=4).eval().to(dev)
        d_a, d_t = 74, 300
    else:
        model = FusionWithEmotionDecoder().eval().to(dev)
        d_a = d_t = 768
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    h_a = torch.randn(B, T_a, d_a, generator=g, device=dev)
    h_t = torch.randn(B, T_t, d_t, generator=g, device=dev)

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
    parts = [agg(k) for k in ("gemm", "attn_proj", "ffn")]
    g_fl, g_ms, g_n = (sum(p[i] for p in parts) for i in range(3))
    a_fl, a_ms, a_n = agg("attention")
    p_fl, p_ms, p_n = agg("attn_proj")
    gemm_tf = g_fl / (g_ms * 1e-3) / 1e12 if g_ms else 0.0
    attn_tf = a_fl / (a_ms * 1e-3) / 1e12 if a_ms else 0.0
    # SURVEY 8(d) "attention-kernel % of roofline": in-proj + QK^T + PV + out-proj FLOPs over their time
    mha_tf = (a_fl + p_fl) / ((a_ms + p_ms) * 1e-3) / 1e12 if (a_ms + p_ms) else 0.0
    # DRAM traffic per launch of the dominant kernel comes from the committed ncu capture of this same
    # command (profiles/r01_traffic.json, written by tools/traffic_from_ncu.py); None if absent
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and args.workload == "ns" and B == WORKLOADS["ns"]["B"]:
        tj = json.load(open(tpath))
        traffic, traffic_note = tj["gemm"]["dram_bytes_per_launch"], tj["gemm"]["note"]
    g_bytes = sum(w for (k, w, a, b) in prof_bytes) / max(g_n, 1) if prof_bytes else None
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05 GEMM, all projections + FFN)",
                "achieved": gemm_tf, "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": gemm_tf / peaks["tf_sust"],
                "traffic": traffic, "traffic_note": traffic_note, "algorithmic_bytes_per_launch": g_bytes,
                "flops_per_launch": g_fl / max(g_n, 1), "avg_launch_ms": g_ms / max(g_n, 1),
                "launches": g_n, "share_of_step": g_ms / ms_total, "peak_source": peaks["src"] + ", sustained"}
    attention_roofline = {"bound": "tensor", "kernel": "attention_fwd3_kernel (tcgen05 QK^T/PV + online softmax)",
                          "achieved": attn_tf, "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": attn_tf / peaks["tf_sust"],
                          "launches": a_n, "share_of_step": a_ms / ms_total,
                          "mha_incl_projections": {"achieved": mha_tf, "frac": mha_tf / peaks["tf_sust"], "unit": "TFLOP/s",
                                                   "launches": a_n + p_n, "share_of_step": (a_ms + p_ms) / ms_total,
                                                   "definition": "SURVEY 8(d): (in-proj + QK^T + PV + out-proj FLOPs) / their kernel time"}}
    fpu = (flops_per_utt(T_a, T_t, d=256, n_e=6, h_beta=128, d_a=74, d_t=300) if mosei else flops_per_utt(T_a, T_t))
    path_tf = fpu * value / world / 1e12

    # ---- end to end from pinned host memory
    e2e = None
    if not args.no_e2e:
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

        # the host-side bf16 pre-cast of every second slab runs on the CPU cores: share them between ranks
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))

        def e2e_step():
            return pipeline.forward_from_host(model, ha_h, ht_h, device=dev, slab=512, out_device="cpu")

        def e2e_stream(n):
            # a stream of batches, two in flight: step i+1 is issued before step i's results are awaited, so its
            # first slab is copied while step i's last slabs compute; every step's H2D and D2H are in the region
            pend = None
            for _ in range(n):
                nxt = pipeline.forward_from_host(model, ha_h, ht_h, device=dev, slab=512, out_device="cpu", wait=False)
                if pend is not None:
                    pend.wait()
                pend = nxt
            return pend.wait()

        def timed(fn, n):
            sync_all()
            t0 = time.perf_counter()
            fn(n)
            torch.cuda.synchronize(dev)
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return world * Be * n / float(dt.item())

        e2e_step()
        n_e2e = 3
        def e2e_sequential(n):
            for _ in range(n):
                e2e_step()      # results dropped at once: their pinned buffers are reused by the next call

        seq = timed(e2e_sequential, n_e2e)
        e2e_stream(2)
        n_stream = 5
        stream = timed(e2e_stream, n_stream)
        e2e = {"value": stream, "unit": UNIT, "batch_per_gpu": Be, "steps": n_stream,
               "sequential_value": seq, "sequential_steps": n_e2e,
               "h2d_bytes_per_step": (pipeline.h2d_bytes(Be, bytes_per_utt, host_cast_every=pipeline.default_host_cast_every())
                                      if d_a % 8 == 0 and d_t % 8 == 0 else Be * bytes_per_utt),
               "host_bytes_read_per_step": Be * bytes_per_utt, "d2h_bytes_per_step": int(lo_shape_bytes(model, Be)),
               "api": "hriemo.pipeline.forward_from_host(model, pinned h_a, pinned h_t, wait=False) -> pinned host logits/beta/z; "
                      "a stream of batches with two in flight (step i+1 issued before step i's results are awaited; the pipeline "
                      "starts cold inside the timed region); every 2nd slab pre-cast to bf16 on the host cores when the rank has >= 8 host threads; "
                      "sequential_value = one call at a time, each awaited before the next"}
        del ha_h, ht_h

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = args.cpu_sample or 256
        times, cores = cpu_forward_timer(T_a, T_t, n_cpu, 2, 1, mosei)
        cpu_baseline = {"value": n_cpu * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{n_cpu} utterances x {len(times)} timed passes (+1 warm-up) of the same workload "
                                  f"(T_a={T_a}, T_t={T_t}), fp32, all host threads, oracle port of the reference forward"}

    train_step = None
    if rank == 0 and world == 1 and args.workload == "ns" and not args.batch and not args.no_train:
        torch.cuda.empty_cache()
        train_step = train_step_leg()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": wl["desc"], "global_batch": world * B, "parallelism": f"batch-sharded x{world}",
                           "inputs": "fp32 features resident in HBM, no masks", "l2": "per-step inputs (7.1 GB at ns) >> 126 MB L2"},
                "roofline": roofline, "attention_roofline": attention_roofline,
                "path": {"flops_per_utt": fpu, "achieved_tflops_per_gpu": path_tf, "frac_of_tensor_peak": path_tf / peaks["tf_sust"]},
                "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clk.summary()}
        if train_step is not None:
            line["train_step"] = train_step
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
