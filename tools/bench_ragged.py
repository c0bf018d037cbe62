"""Ragged-batch throughput (the padded-tensor form of the reference's collate):
the north-star model on B utterances whose valid lengths are uniform in [T/2, T], (a) padded forward with
masks, (b) length-bucketed forward (pipeline.forward_bucketed), (c) end to end from pinned host memory,
padded and bucketed.  CUDA events / wall clock around K steps after W warm-ups; one JSON line.
    python tools/bench_ragged.py [--batch 4096] [--steps 5] [--warmup 3]"""
import json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import pipeline, lib
from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


B, K, W = arg("--batch", 4096), arg("--steps", 5), arg("--warmup", 3)
T_a, T_t = 500, 64
dev = torch.device("cuda")
torch.manual_seed(1234)
model = FusionWithEmotionDecoder().eval().to(dev)
g = torch.Generator().manual_seed(77)
la = torch.randint(T_a // 2, T_a + 1, (B,), generator=g)
lt = torch.randint(T_t // 2, T_t + 1, (B,), generator=g)
ha = torch.empty((B, T_a, 768)).pin_memory(); ht = torch.empty((B, T_t, 768)).pin_memory()
ha.normal_(generator=g); ht.normal_(generator=g)
ma = (torch.arange(T_a)[None] >= la[:, None]).pin_memory(); mt = (torch.arange(T_t)[None] >= lt[:, None]).pin_memory()
ha[ma] = 0.0; ht[mt] = 0.0     # what the collate produces: zero padding
d_ha, d_ht, d_ma, d_mt = ha.to(dev), ht.to(dev), ma.to(dev), mt.to(dev)


def timed(fn, device_events=True):
    for _ in range(W):
        out = fn()
    torch.cuda.synchronize()
    n0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms = e0.elapsed_time(e1) if device_events else wall
    return out, ms / K, (lib.launch_count() - n0) // K


ref, ms_pad, l_pad = timed(lambda: model(d_ha, d_ht, d_ma, d_mt))
out, ms_bkt, l_bkt = timed(lambda: pipeline.forward_bucketed(model, d_ha, d_ht, d_ma, d_mt))
err = [float((a - b).abs().max()) for a, b in zip(out, ref)]
sweep = {}
for rows in (128 * 500, 256 * 500, 512 * 500, 1024 * 500, 2048 * 500):
    _, ms_r, l_r = timed(lambda: pipeline.forward_bucketed(model, d_ha, d_ht, d_ma, d_mt, rows_per_slab=rows, max_utts=8192))
    sweep[str(rows)] = {"ms": ms_r, "utt_s": B / ms_r * 1e3, "launches": l_r}
_, ms_e2e_pad, _ = timed(lambda: pipeline.forward_from_host(model, ha, ht, ma, mt, device=dev), device_events=False)
out_h, ms_e2e_bkt, _ = timed(lambda: pipeline.forward_from_host(model, ha, ht, ma, mt, device=dev, bucket=True), device_events=False)
err_h = [float((a - b.cpu()).abs().max()) for a, b in zip(out_h, ref)]
# (d) the same utterances from a packed bf16 shard in the page cache
import tempfile
from hriemo import shards
tmpd = tempfile.mkdtemp()
spath = os.path.join(tmpd, "ragged.hriemo")
t0 = time.perf_counter()
shards.write_shard(spath, [(ha[i], ma[i], ht[i], mt[i]) for i in range(B)])
t_write = time.perf_counter() - t0
sh = shards.Shard(spath)
out_s, ms_shard, _ = timed(lambda: pipeline.forward_from_shard(model, sh, device=dev), device_events=False)
back = sh.original_order()
err_s = [float((a - b.cpu()[back]).abs().max()) for a, b in zip(out_s, ref)]
shard_bytes = os.path.getsize(spath)
sh.close(); os.remove(spath)
lens_a, lens_t = pipeline.valid_lengths(ma, B, T_a), pipeline.valid_lengths(mt, B, T_t)
_, buckets = pipeline.bucket_plan(lens_a, lens_t, T_a, T_t)
st = pipeline.bucket_stats(lens_a, lens_t, T_a, T_t, buckets)
bytes_pad = pipeline.h2d_bytes(B, (T_a + T_t) * 768 * 4)
_, hb = pipeline.bucket_plan(lens_a, lens_t, T_a, T_t, rows_per_slab=512 * T_a, max_utts=2048, ramp=True)
bytes_bkt = sum((b.end - b.start) * (b.T_a + b.T_t) * 768 * 2 for b in hb)
print(json.dumps({
    "workload": f"FusionWithEmotionDecoder fwd, B={B}, T_a=500, T_t=64, valid lengths uniform in [T/2, T] (tail PAD masks)",
    "steps": K, "warmup": W, "rows": st,
    "device_resident": {"padded_utt_s": B / ms_pad * 1e3, "bucketed_utt_s": B / ms_bkt * 1e3, "speedup": ms_pad / ms_bkt,
                        "padded_ms": ms_pad, "bucketed_ms": ms_bkt, "launches_per_step": [l_pad, l_bkt],
                        "max_abs_diff_logits_beta_z": err, "rows_per_slab_sweep": sweep},
    "e2e_from_host": {"padded_utt_s": B / ms_e2e_pad * 1e3, "bucketed_utt_s": B / ms_e2e_bkt * 1e3,
                      "speedup": ms_e2e_pad / ms_e2e_bkt, "padded_ms": ms_e2e_pad, "bucketed_ms": ms_e2e_bkt,
                      "h2d_bytes_per_step": [bytes_pad, bytes_bkt], "max_abs_diff_logits_beta_z": err_h},
    "e2e_from_shard": {"utt_s": B / ms_shard * 1e3, "ms": ms_shard, "shard_bytes": shard_bytes, "dtype": "bf16",
                       "fp32_pickles_equivalent_bytes": int(B * (T_a + T_t) * 768 * 4), "write_s": t_write,
                       "max_abs_diff_logits_beta_z": err_s, "note": "shard in the page cache; results in shard order"},
}))
