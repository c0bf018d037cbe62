"""Gate / decoder kernels of one 2048-utterance slab against the HBM roofline (algorithmic bytes / CUDA-event time).
Run twice for an A / B: plain (streaming blend, warp-private decoder attention) and with
HRIEMO_GATE_BLEND_V1=1 HRIEMO_DECODER_ATTN_V1=1 (the former kernels).  Diagnostic tool; bench.py is the contract benchmark."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import ops

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6536.7


def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # L2 flush between timed launches
    evs = []
    for _ in range(iters):
        big.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main():
    dev = "cuda"
    res = []
    form = "v1" if os.environ.get("HRIEMO_GATE_BLEND_V1") or os.environ.get("HRIEMO_DECODER_ATTN_V1") or os.environ.get("HRIEMO_LN_BWD_V1") or os.environ.get("HRIEMO_DECODER_ATTN_BWD_V1") else "v2"
    # gate blend of the north star: pending LayerNorms with statistics on both streams, bf16 out
    for (B, T_a, L, d) in [(2048, 500, 64, 768), (4096, 300, 128, 256)]:
        g = torch.Generator(device=dev).manual_seed(1)
        a = torch.randn(B * T_a, d, device=dev, generator=g).bfloat16(); t = torch.randn(B * L, d, device=dev, generator=g).bfloat16()
        vecs = [torch.rand(d, device=dev, generator=g) + 0.5 for _ in range(8)]
        st = lambda x: torch.stack([x.float().mean(1), torch.rsqrt(x.float().var(1, unbiased=False) + 1e-5)], 1).contiguous()
        sa, stt = st(a), st(t)
        w = torch.sigmoid(torch.randn(B, d, device=dev, generator=g))
        fn = lambda: ops.gate_blend(a, T_a, t, (vecs[0], vecs[1]), (vecs[2], vecs[3]), w, B, L, pre_ln_a=(vecs[4], vecs[5], sa), pre_ln_t=(vecs[6], vecs[7], stt))
        ms = timeit(fn)
        gb = 3 * B * L * d * 2 / 1e9
        res.append(dict(kernel="gate_blend", form=form, B=B, T_a=T_a, L=L, d=d, us=ms * 1e3, gbs=gb / ms * 1e3, frac_hbm=gb / ms * 1e3 / PEAK))
        print(res[-1], flush=True)
        if B == 2048:
            pad = torch.arange(T_a, device=dev)[None, :] >= torch.randint(T_a // 2, T_a + 1, (B, 1), device=dev)
            for name, m in (("ln_masked_mean", None), ("ln_masked_mean_ragged", pad)):
                ms = timeit(lambda: ops.ln_masked_mean(a, vecs[0], vecs[1], m, B, T_a, pre_ln=(vecs[4], vecs[5], sa)))
                frac = 1.0 if m is None else float((~pad).float().mean())
                gb = B * T_a * d * 2 * frac / 1e9
                res.append(dict(kernel=name, B=B, T=T_a, d=d, us=ms * 1e3, gbs=gb / ms * 1e3, frac_hbm=gb / ms * 1e3 / PEAK))
                print(res[-1], flush=True)
    # LayerNorm backward of one 512-utterance training slab (rows = 512 * 500): x, dy read, dx written
    for (rows, d) in [(256000, 768), (32768, 768), (614400, 256)]:
        x = torch.randn(rows, d, device=dev).bfloat16(); dy = torch.randn(rows, d, device=dev).bfloat16(); gam = torch.rand(d, device=dev) + 0.5
        ms = timeit(lambda: ops.layernorm_backward(x, dy, gam))
        gb = 3 * rows * d * 2 / 1e9
        res.append(dict(kernel="layernorm_backward", form="v1" if os.environ.get("HRIEMO_LN_BWD_V1") else "v2", rows=rows, d=d, us=ms * 1e3, gbs=gb / ms * 1e3, frac_hbm=gb / ms * 1e3 / PEAK))
        print(res[-1], flush=True)
    # decoder attention: cross (4 queries x 64 keys), self (4 x 4), MOSEI (6 x 128, dh 64)
    for (B, H, Nq, Tk, dh) in [(2048, 8, 4, 64, 96), (2048, 8, 4, 4, 96), (4096, 4, 6, 128, 64), (2048, 8, 4, 50, 96)]:
        d = H * dh
        q = torch.randn(B * Nq, d, device=dev).bfloat16(); kv = torch.randn(B * Tk, 2 * d, device=dev).bfloat16()
        pad = torch.arange(Tk, device=dev)[None, :] >= torch.randint(max(Tk // 2, 1), Tk + 1, (B, 1), device=dev)
        for name, m in (("decoder_attention", None), ("decoder_attention_masked", pad)):
            ms = timeit(lambda: ops.small_attention(q, kv[:, :d], kv[:, d:], m, B, H, Nq, Tk, dh))
            gb = (B * Tk * 2 * d * 2 + 2 * B * Nq * d * 2) / 1e9
            res.append(dict(kernel=name, form=form, B=B, H=H, Nq=Nq, Tk=Tk, dh=dh, us=ms * 1e3, gbs=gb / ms * 1e3, frac_hbm=gb / ms * 1e3 / PEAK))
            print(res[-1], flush=True)
    # decoder attention backward of one 512-utterance training step: cross (4 x 64) and self (4 x 4)
    for (B, H, Nq, Tk, dh) in [(512, 8, 4, 64, 96), (512, 8, 4, 4, 96)]:
        d = H * dh
        q = torch.randn(B * Nq, d, device=dev).bfloat16(); kv = torch.randn(B * Tk, 2 * d, device=dev).bfloat16(); do = torch.randn(B * Nq, d, device=dev).bfloat16()
        ms = timeit(lambda: ops.small_attention_backward(q, kv[:, :d], kv[:, d:], do, None, B, H, Nq, Tk, dh))
        gb = (2 * B * Tk * 2 * d * 2 + 3 * B * Nq * d * 2) / 1e9
        res.append(dict(kernel="decoder_attention_backward", form="v1" if os.environ.get("HRIEMO_DECODER_ATTN_BWD_V1") else "v2", B=B, H=H, Nq=Nq, Tk=Tk, dh=dh, us=ms * 1e3, gbs=gb / ms * 1e3, frac_hbm=gb / ms * 1e3 / PEAK))
        print(res[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"bench_decoder_gate_{form}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
