"""Micro-benchmarks of the individual kernels against the measured peaks (and cuBLAS / SDPA
as same-box yardsticks).  Diagnostic tool; bench.py is the contract benchmark."""
import json, math, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import lib as L, ops

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]

def main():
    dev = "cuda"
    res = []
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    gemm_shapes = [] if only not in (None, "gemm") else [(2048000 // 4, 2304, 768, L.EPI_BIAS), (2048000 // 4, 3072, 768, L.EPI_BIAS_RELU), (2048000 // 4, 768, 3072, L.EPI_BIAS),
                           (2048000 // 4, 768, 768, L.EPI_BIAS), (262144 // 4, 2304, 768, L.EPI_BIAS), (16384, 768, 768, L.EPI_BIAS)]
    for (M, N, K, epi) in gemm_shapes:
        a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16(); b = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        med, best = timeit(lambda: ops.gemm(a, w, b, epi, out=out, cta_pair=1))
        med2, best2 = timeit(lambda: ops.gemm(a, w, b, epi, out=out, cta_pair=2))
        medc, bestc = timeit(lambda: torch.nn.functional.linear(a, w, b.bfloat16()))
        fl = 2.0 * M * N * K
        res.append(dict(kernel="gemm", M=M, N=N, K=K, ms_1cta=med, tflops_1cta=fl / med / 1e9, ms_pair=med2, tflops_pair=fl / med2 / 1e9,
                        frac_burst_pair=fl / med2 / 1e9 / PEAKS["bf16_tflops"], cublas_ms=medc, cublas_tflops=fl / medc / 1e9))
        print(res[-1], flush=True)
    attn_shapes = [] if only not in (None, "attention") else [(512, 8, 500, 500, 96), (512, 8, 500, 64, 96), (512, 8, 64, 500, 96), (512, 8, 300, 300, 96), (512, 4, 300, 128, 64), (64, 8, 1000, 1000, 96)]
    for (B, H, Tq, Tk, dh) in attn_shapes:
        d = H * dh
        q = torch.randn(B * Tq, d, device=dev).bfloat16(); k = torch.randn(B * Tk, d, device=dev).bfloat16()
        v = torch.randn(B, Tk, d, device=dev).bfloat16()
        med, best = timeit(lambda: ops.attention(q, k, v.view(B * Tk, d), None, B, H, Tq, Tk, dh))
        qh = q.view(B, Tq, H, dh).transpose(1, 2); kh = k.view(B, Tk, H, dh).transpose(1, 2); vh = v.view(B, Tk, H, dh).transpose(1, 2)
        meds, _ = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh))
        fl = 4.0 * B * H * Tq * Tk * dh
        res.append(dict(kernel="attention", B=B, H=H, Tq=Tq, Tk=Tk, dh=dh, ms=med, tflops=fl / med / 1e9, frac_burst=fl / med / 1e9 / PEAKS["bf16_tflops"], sdpa_ms=meds))
        print(res[-1], flush=True)
    if only in (None, "attention", "pairs"):
        # short query sequences: two heads per work item against one (bit-identical results)
        for (B, H, Tq, Tk, dh) in [(2048, 8, 64, 500, 96), (2048, 8, 64, 64, 96), (2048, 4, 128, 128, 64), (2048, 4, 128, 300, 64)]:
            d = H * dh
            q = torch.randn(B * Tq, d, device=dev).bfloat16(); k = torch.randn(B * Tk, d, device=dev).bfloat16()
            v = torch.randn(B * Tk, d, device=dev).bfloat16()
            m1, _ = timeit(lambda: ops.attention(q, k, v, None, B, H, Tq, Tk, dh, pair_heads=False))
            m2, _ = timeit(lambda: ops.attention(q, k, v, None, B, H, Tq, Tk, dh, pair_heads=True))
            gb = (B * Tq * d * 2 + B * Tk * d * 2) * 2 / 1e9
            res.append(dict(kernel="attention_head_pairs", B=B, H=H, Tq=Tq, Tk=Tk, dh=dh, ms_one_head=m1, ms_pairs=m2, speedup=m1 / m2,
                            gbs_pairs=gb / m2 * 1e3, frac_hbm=gb / m2 * 1e3 / PEAKS["hbm_gbs"]))
            print(res[-1], flush=True)
    if only in (None, "wgrad"):
        # weight gradients of the encoder's Linear layers for a 512-utterance training slab (M = 512 * 500 rows)
        for (M, N, K) in [(256000, 2304, 768), (256000, 3072, 768), (256000, 768, 3072), (256000, 768, 768), (32768, 768, 768)]:
            dy = torch.randn(M, N, device=dev).bfloat16(); x = torch.randn(M, K, device=dev).bfloat16()
            med, best = timeit(lambda: ops.linear_wgrad(dy, x))
            med_nb, _ = timeit(lambda: ops.linear_wgrad(dy, x, want_bias=False))
            medc, _ = timeit(lambda: torch.matmul(dy.t(), x))
            fl = 2.0 * M * N * K
            res.append(dict(kernel="linear_wgrad", M=M, N=N, K=K, ms=med, ms_without_bias=med_nb, tflops=fl / med / 1e9, frac_sustained=fl / med / 1e9 / PEAKS.get("bf16_tflops_sustained", PEAKS["bf16_tflops"]),
                            cublas_ms=medc, cublas_tflops=fl / medc / 1e9))
            print(res[-1], flush=True)
    if only in (None, "small"):
        # decoder attention of one 2048-utterance slab: cross (4 queries x 64 keys) and self (4 x 4)
        for (B, H, Nq, Tk, dh) in [(2048, 8, 4, 64, 96), (2048, 8, 4, 4, 96), (4096, 4, 6, 128, 64)]:
            d = H * dh
            q = torch.randn(B * Nq, d, device=dev).bfloat16(); kv = torch.randn(B * Tk, 2 * d, device=dev).bfloat16()
            m1, _ = timeit(lambda: ops.small_attention(q, kv[:, :d], kv[:, d:], None, B, H, Nq, Tk, dh))
            m2, _ = timeit(lambda: ops.small_attention(q, kv[:, :d], kv[:, d:], None, B, H, Nq, Tk, dh, want_probs=True))
            gb = (B * Tk * 2 * d * 2 + 2 * B * Nq * d * 2) / 1e9
            res.append(dict(kernel="small_attention", B=B, H=H, Nq=Nq, Tk=Tk, dh=dh, ms=m1, ms_with_probs=m2,
                            gbs=gb / m1 * 1e3, frac_hbm=gb / m1 * 1e3 / PEAKS["hbm_gbs"]))
            print(res[-1], flush=True)
    if only in (None, "attention", "ragged"):
        # ragged batch (key lengths uniform in [T/2, T]): skipping trailing all-PAD key tiles
        for (B, H, Tq, Tk, dh) in [(512, 8, 500, 500, 96), (512, 8, 64, 500, 96), (512, 8, 300, 300, 96)]:
            d = H * dh
            q = torch.randn(B * Tq, d, device=dev).bfloat16(); k = torch.randn(B * Tk, d, device=dev).bfloat16()
            v = torch.randn(B * Tk, d, device=dev).bfloat16()
            lens = torch.randint(Tk // 2, Tk + 1, (B, 1), device=dev)
            pad = torch.arange(Tk, device=dev)[None, :] >= lens
            m0, _ = timeit(lambda: ops.attention(q, k, v, pad, B, H, Tq, Tk, dh, skip_padded_tiles=False))
            m1, _ = timeit(lambda: ops.attention(q, k, v, pad, B, H, Tq, Tk, dh, skip_padded_tiles=True))
            res.append(dict(kernel="attention_ragged", B=B, H=H, Tq=Tq, Tk=Tk, dh=dh, ms_all_tiles=m0, ms_skip=m1, speedup=m0 / m1))
            print(res[-1], flush=True)
    if only not in (None, "elem"):
        return
    rows, d = 2048000 // 4, 768
    x = torch.randn(rows, d, device=dev).bfloat16(); g = torch.ones(d, device=dev); bb = torch.zeros(d, device=dev)
    med, best = timeit(lambda: ops.layernorm(x, g, bb))
    res.append(dict(kernel="layernorm", rows=rows, d=d, ms=med, gbs=rows * d * 4 / med / 1e6, frac=rows * d * 4 / med / 1e6 / PEAKS["hbm_gbs"]))
    print(res[-1], flush=True)
    xf = torch.randn(rows, d, device=dev)
    med, best = timeit(lambda: ops.cast_bf16(xf))
    res.append(dict(kernel="cast", rows=rows, d=d, ms=med, gbs=rows * d * 6 / med / 1e6, frac=rows * d * 6 / med / 1e6 / PEAKS["hbm_gbs"]))
    print(res[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_kernels.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
