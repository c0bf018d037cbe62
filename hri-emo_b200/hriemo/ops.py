"""torch.Tensor-level wrappers over the C ABI (one function per kernel).

Every function enqueues on the current CUDA stream and returns freshly
allocated tensors owned by the caller.  2-D operands may be row-strided views
(e.g. a column slice of a wider buffer): the row pitch is passed as the leading
dimension, the last dimension must be contiguous.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import lib as _l

bf16 = torch.bfloat16
f32 = torch.float32
LN_EPS = 1e-5


# When set to a list (bench.py), every tensor-core launch is bracketed by CUDA events on the
# launching stream and (kind, algorithmic FLOPs, start, end) is appended.
PROFILE = None
# When set to a list, GEMM launches append (tag, algorithmic bytes = (M*K + N*K + M*N [+ M*N residual]) * 2, None, None).
PROFILE_BYTES = None
# When set to a list, attention launches append (Tq, Tk, algorithmic FLOPs, algorithmic bytes = Q + K + V + O in bf16, start,
# end): the cross shapes of a layer are HBM work, the 500 x 500 self-attention tensor-pipe work (bench.py reports each
# against its own roofline).
PROFILE_ATTN = None


def _prof_begin(kind: str, work: float):
    if PROFILE is None:
        return None
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    return (kind, work, a, b)


def _prof_end(tok) -> None:
    if tok is not None:
        tok[3].record()
        PROFILE.append(tok)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_tensor_device(fn):
    """Run the wrapper with the CUDA device of its first tensor argument current (the C ABI launches
    on the current device and stream; a model on cuda:1 must not launch on cuda:0)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = next((a.device for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _chk2d(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise _l.HriemoError(f"{name}: expected a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise _l.HriemoError(f"{name}: expected {dtype}, got {t.dtype}")
    if t.dim() != 2 or t.stride(1) != 1:
        raise _l.HriemoError(f"{name}: expected a 2-D tensor with a contiguous last dim, got {tuple(t.shape)} / {t.stride()}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _mask_u8(mask: Optional[torch.Tensor], B: int, T: int, name: str) -> Optional[torch.Tensor]:
    if mask is None:
        return None
    if tuple(mask.shape) != (B, T):
        raise _l.HriemoError(f"{name}: mask shape {tuple(mask.shape)} != {(B, T)}")
    m = mask.contiguous()
    return m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ------------------------------------------------------------------ cast
@_on_tensor_device
def cast_bf16(x: torch.Tensor, ld_out: Optional[int] = None) -> torch.Tensor:
    """fp32 [rows, cols] -> bf16 [rows, ld_out] (zero-padded columns)."""
    _chk2d(x, f32, "cast_bf16")
    rows, cols = x.shape
    ld_out = round_up(cols, 8) if ld_out is None else ld_out
    out = torch.empty((rows, ld_out), dtype=bf16, device=x.device)
    _l.check(_l.load().hriemo_cast_f32_to_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), ld_out, rows, cols,
                                                _stream()), "cast_f32_to_bf16")
    return out


# ------------------------------------------------------------------ GEMM
@_on_tensor_device
def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], epilogue: int = _l.EPI_BIAS,
         resid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         cta_pair: int = 0, a_ln=None, resid_ln=None, want_stats: bool = False, tag: str = "gemm"):
    """out = epilogue(a[M,K] @ w[N,K]^T + bias) on tcgen05 tensor cores.
    cta_pair: 0 = library picks the tile form, 1 = one-CTA tiles, 2 = CTA-pair (cta_group::2) tiles.
    Fused LayerNorm (include/hriemo.h): a_ln = (stats [M,2], colsum [N]) -- `a` is pre-LayerNorm and
    `w` / `bias` are the folded operands of fold_ln_weight; resid_ln = (stats [M,2], gamma, beta) --
    `resid` is pre-LayerNorm; want_stats -> returns (out, stats [M,2] = (mean, rstd) of out's rows)."""
    _chk2d(a, bf16, "gemm A")
    _chk2d(w, bf16, "gemm W")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise _l.HriemoError(f"gemm: K mismatch {a.shape} vs {w.shape}")
    f32_out = epilogue in (_l.EPI_BIAS_RESID_F32, _l.EPI_BIAS_F32)
    if out is None:
        out = torch.empty((M, N), dtype=f32 if f32_out else bf16, device=a.device)
    _chk2d(out, f32 if f32_out else bf16, "gemm out")
    args = _l.GemmArgs()
    args.A, args.lda, args.W, args.ldw = a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0)
    args.bias = _ptr(bias)
    args.M, args.N, args.K, args.epilogue = M, N, K, epilogue
    args.cta_pair = cta_pair
    args.out, args.ldo = out.data_ptr(), out.stride(0)
    if resid is not None:
        _chk2d(resid, f32 if f32_out else bf16, "gemm resid")
        args.resid, args.ldr = resid.data_ptr(), resid.stride(0)
    keep = []
    if a_ln is not None:
        st, cs = a_ln
        _chk_f32(st, (M, 2), "gemm a_ln stats")
        _chk_f32(cs, (N,), "gemm a_ln colsum")
        args.a_stats, args.a_colsum = st.data_ptr(), cs.data_ptr()
    if resid_ln is not None:
        st, g, b = resid_ln
        _chk_f32(st, (M, 2), "gemm resid_ln stats")
        _chk_f32(g, (N,), "gemm resid_ln gamma")
        _chk_f32(b, (N,), "gemm resid_ln beta")
        args.resid_stats, args.resid_gamma, args.resid_beta = st.data_ptr(), g.data_ptr(), b.data_ptr()
    partials = None
    if want_stats:
        partials = torch.empty(((N + 63) // 64, M, 2), dtype=f32, device=a.device)
        args.stats_out = partials.data_ptr()
    tok = _prof_begin(tag, 2.0 * M * N * K)
    if PROFILE_BYTES is not None:
        osz = 4 if f32_out else 2
        PROFILE_BYTES.append((tag, 2.0 * (M * K + N * K) + osz * M * N * (2 if resid is not None else 1), None, None))
    _l.check(_l.load().hriemo_gemm_bf16(C.byref(args), _stream()), "gemm_bf16")
    _prof_end(tok)
    if not want_stats:
        return out
    stats = torch.empty((M, 2), dtype=f32, device=a.device)
    _l.check(_l.load().hriemo_ln_stats_finalize(partials.data_ptr(), partials.shape[0], M, N, LN_EPS,
                                                 stats.data_ptr(), _stream()), "ln_stats_finalize")
    return out, stats


def _chk_f32(t: torch.Tensor, shape, name: str) -> None:
    if not t.is_cuda or t.dtype != f32 or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise _l.HriemoError(f"{name}: expected a contiguous CUDA fp32 tensor of shape {tuple(shape)}, "
                             f"got {t.dtype} {tuple(t.shape)} on {t.device}")


@_on_tensor_device
def fold_ln_weight(w: torch.Tensor, bias: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor,
                   k_pad: Optional[int] = None):
    """One-time preparation of a Linear that consumes LN(x): returns (w*gamma as bf16 [N, K_pad],
    colsum [N] fp32, bias + w.beta [N] fp32)."""
    _chk2d(w, f32, "fold_ln_weight W")
    N, K = w.shape
    ld = round_up(K, 8) if k_pad is None else k_pad
    wf = torch.zeros((N, ld), dtype=bf16, device=w.device)
    cs = torch.empty((N,), dtype=f32, device=w.device)
    bf = torch.empty((N,), dtype=f32, device=w.device)
    _l.check(_l.load().hriemo_fold_ln_weight(w.data_ptr(), w.stride(0), gamma.data_ptr(), beta.data_ptr(), _ptr(bias),
                                              wf.data_ptr(), ld, cs.data_ptr(), bf.data_ptr(), N, K, _stream()),
             "fold_ln_weight")
    return wf, cs, bf


# ------------------------------------------------------------------ attention
@_on_tensor_device
def kv_steps(key_pad: torch.Tensor) -> torch.Tensor:
    """[B, Tk] PAD mask -> int32 [B]: 64-key tiles up to the last valid key (>= 1) of each utterance."""
    B, Tk = key_pad.shape
    m = _mask_u8(key_pad, B, Tk, "kv_steps")
    out = torch.empty((B,), dtype=torch.int32, device=key_pad.device)
    _l.check(_l.load().hriemo_attention_kv_steps(m.data_ptr(), B, Tk, out.data_ptr(), _stream()), "attention_kv_steps")
    return out


@_on_tensor_device
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, key_pad: Optional[torch.Tensor],
              B: int, H: int, Tq: int, Tk: int, dh: int, skip_padded_tiles: bool = True,
              pair_heads: bool = True, want_lse: bool = False, out: Optional[torch.Tensor] = None, drop=None):
    """q: [B*Tq, >=H*dh] view, k / v: [B*Tk, >=H*dh] views (column slices of a packed projection are
    fine).  Returns [B*Tq, H*dh] bf16 (written into `out` when given: a [B*Tq, >= H*dh] view with 16-byte aligned rows).  pair_heads=False switches the two-heads-per-work-item form of
    short query sequences off (include/hriemo.h: no_head_pairs; same result bit for bit).
    want_lse: also return lse [B, H, Tq] fp32 = ln sum_k exp(scale * q.k) over the unmasked keys."""
    _chk2d(q, bf16, "attention q")
    _chk2d(k, bf16, "attention k")
    _chk2d(v, bf16, "attention v")
    if q.shape[0] != B * Tq or k.shape[0] != B * Tk or v.shape[0] != B * Tk:
        raise _l.HriemoError(f"attention: row counts {q.shape[0]}, {k.shape[0]}, {v.shape[0]} do not match "
                             f"B*Tq={B * Tq}, B*Tk={B * Tk}")
    if out is None:
        out = torch.empty((B * Tq, H * dh), dtype=bf16, device=q.device)
    else:
        _chk2d(out, bf16, "attention out")
        if out.shape[0] != B * Tq or out.shape[1] < H * dh:
            raise _l.HriemoError(f"attention: out shape {tuple(out.shape)} does not hold [{B * Tq}, {H * dh}]")
    m = _mask_u8(key_pad, B, Tk, "attention")
    args = _l.AttnArgs()
    args.q, args.ldq, args.k, args.ldk = q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0)
    args.v, args.ldv = v.data_ptr(), v.stride(0)
    args.key_pad = _ptr(m)
    steps = None
    if m is not None and skip_padded_tiles and Tk > 320:   # measured: pays from 6 key tiles on
        steps = kv_steps(m)   # trailing all-PAD key tiles are not processed (bit-identical result)
        args.kv_steps = steps.data_ptr()
    args.out, args.ldo = out.data_ptr(), out.stride(0)
    args.B, args.H, args.Tq, args.Tk, args.dh = B, H, Tq, Tk, dh
    args.scale = 1.0 / math.sqrt(dh)
    args.no_head_pairs = 0 if pair_heads else 1
    lse = torch.empty((B, H, Tq), dtype=f32, device=q.device) if want_lse else None
    args.lse = _ptr(lse)
    if drop is not None:   # training: dropout on the probabilities, (p8, scale, key) from hriemo.dropout.Drop.site
        args.drop_p8, args.drop_scale, args.drop_key = drop
    tok = _prof_begin("attention", 4.0 * B * H * Tq * Tk * dh)
    _l.check(_l.load().hriemo_attention_bf16(C.byref(args), _stream()), "attention_bf16")
    _prof_end(tok)
    if tok is not None and PROFILE_ATTN is not None:
        PROFILE_ATTN.append((Tq, Tk, tok[1], 2.0 * B * H * dh * (2 * Tq + 2 * Tk), tok[2], tok[3]))
    return (out, lse) if want_lse else out


@_on_tensor_device
def attention_probs(q: torch.Tensor, k: torch.Tensor, key_pad: Optional[torch.Tensor], B: int, H: int,
                    Tq: int, Tk: int, dh: int) -> torch.Tensor:
    """Head-averaged softmax probabilities [B, Tq, Tk] fp32 (return_attention path)."""
    _chk2d(q, bf16, "attention_probs q")
    _chk2d(k, bf16, "attention_probs k")
    probs = torch.empty((B, Tq, Tk), dtype=f32, device=q.device)
    m = _mask_u8(key_pad, B, Tk, "attention_probs")
    _l.check(_l.load().hriemo_attention_probs(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), _ptr(m),
                                               probs.data_ptr(), B, H, Tq, Tk, dh, 1.0 / math.sqrt(dh),
                                               _stream()), "attention_probs")
    return probs


@_on_tensor_device
def small_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, key_pad: Optional[torch.Tensor],
                    B: int, H: int, Nq: int, Tk: int, dh: int, want_probs: bool = False, drop=None):
    """Decoder attention: q [B*Nq, *], k/v [B*Tk, *] row-major views.  Returns (out bf16, probs|None).
    drop = (p8, scale, key): dropout on the probabilities (training; no attention map then)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk2d(t, bf16, f"small_attention {n}")
    out = torch.empty((B * Nq, H * dh), dtype=bf16, device=q.device)
    if drop is not None:
        if want_probs:
            raise _l.HriemoError("small_attention: attention maps are not returned together with dropout")
        m = _mask_u8(key_pad, B, Tk, "small_attention")
        _l.check(_l.load().hriemo_small_attention_dropout(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
                                                           v.stride(0), _ptr(m), out.data_ptr(), out.stride(0), B, H, Nq, Tk,
                                                           dh, 1.0 / math.sqrt(dh), drop[0], drop[2], drop[1], _stream()),
                 "small_attention_dropout")
        return out, None
    probs = torch.empty((B, Nq, Tk), dtype=f32, device=q.device) if want_probs else None
    m = _mask_u8(key_pad, B, Tk, "small_attention")
    _l.check(_l.load().hriemo_small_attention(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
                                               v.stride(0), _ptr(m), out.data_ptr(), out.stride(0), _ptr(probs),
                                               B, H, Nq, Tk, dh, 1.0 / math.sqrt(dh), _stream()),
             "small_attention")
    return out, probs


# ------------------------------------------------------------------ LayerNorm
@_on_tensor_device
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, want_bf16: bool = True,
              want_f32: bool = False, eps: float = LN_EPS):
    """x: [rows, d] bf16 or f32 (residual already added).  Returns (y_bf16|None, y_f32|None)."""
    if x.dtype not in (bf16, f32):
        raise _l.HriemoError(f"layernorm: unsupported dtype {x.dtype}")
    _chk2d(x, x.dtype, "layernorm x")
    rows, d = x.shape
    yb = torch.empty((rows, d), dtype=bf16, device=x.device) if want_bf16 else None
    yf = torch.empty((rows, d), dtype=f32, device=x.device) if want_f32 else None
    _l.check(_l.load().hriemo_layernorm(x.data_ptr(), int(x.dtype == f32), x.stride(0), gamma.data_ptr(),
                                         beta.data_ptr(), eps, _ptr(yb), _ptr(yf), d, rows, d, _stream()),
             "layernorm")
    return yb, yf


# ------------------------------------------------------------------ gate
@_on_tensor_device
def ln_masked_mean(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor],
                   pad: Optional[torch.Tensor], B: int, T: int, apply_ln: bool = True,
                   eps: float = LN_EPS, pre_ln=None) -> torch.Tensor:
    """x: [B*T, d] bf16 -> pooled [B, d] fp32 = masked_mean_t(LN(x)); with pre_ln = (gamma', beta'[, stats'])
    x is pre-LayerNorm and the rows are LN(LN'(x)) (stats' [B*T, 2] spares the kernel LN's statistics)."""
    _chk2d(x, bf16, "ln_masked_mean x")
    d = x.shape[1]
    pooled = torch.empty((B, d), dtype=f32, device=x.device)
    m = _mask_u8(pad, B, T, "ln_masked_mean")
    _l.check(_l.load().hriemo_ln_masked_mean(x.data_ptr(), x.stride(0), _ptr(gamma), _ptr(beta), eps,
                                              int(apply_ln), _ptr(m), pooled.data_ptr(), d, B, T, d,
                                              _ptr(pre_ln[0]) if pre_ln else None,
                                              _ptr(pre_ln[1]) if pre_ln else None,
                                              _ptr(pre_ln[2]) if pre_ln and len(pre_ln) > 2 else None, _stream()),
             "ln_masked_mean")
    return pooled


@_on_tensor_device
def gate_input(a_pool: torch.Tensor, t_pool: torch.Tensor) -> torch.Tensor:
    B, d = a_pool.shape
    g = torch.empty((B, 4 * d), dtype=f32, device=a_pool.device)
    _l.check(_l.load().hriemo_gate_input(a_pool.data_ptr(), t_pool.data_ptr(), g.data_ptr(), B, d, _stream()),
             "gate_input")
    return g


@_on_tensor_device
def sgemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act: int = _l.ACT_NONE) -> torch.Tensor:
    """fp32 CUDA-core GEMM: act(a[M,K] @ w[N,K]^T + bias)."""
    _chk2d(a, f32, "sgemm A")
    _chk2d(w, f32, "sgemm W")
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=f32, device=a.device)
    _l.check(_l.load().hriemo_sgemm_f32(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias),
                                         out.data_ptr(), N, M, N, K, act, _stream()), "sgemm_f32")
    return out


@_on_tensor_device
def gate_blend(a: torch.Tensor, T_a: int, t: torch.Tensor, ln_a, ln_t, w: torch.Tensor, B: int, L: int,
               apply_ln: bool = True, w_is_scalar: bool = False, want_bf16: bool = True,
               want_f32: bool = False, eps: float = LN_EPS, pre_ln_a=None, pre_ln_t=None):
    """a: [B*T_a, d] bf16, t: [B*L, d] bf16, w: [B, d] (or [B, 1] scalar gate) fp32.
    Returns (h_bf16|None, h_f32|None, beta [B,1] fp32)."""
    _chk2d(a, bf16, "gate_blend a")
    _chk2d(t, bf16, "gate_blend t")
    d = a.shape[1]
    hb = torch.empty((B * L, d), dtype=bf16, device=a.device) if want_bf16 else None
    hf = torch.empty((B * L, d), dtype=f32, device=a.device) if want_f32 else None
    beta = torch.empty((B, 1), dtype=f32, device=a.device)
    ga, ba = (ln_a if apply_ln else (None, None))
    gt, bt = (ln_t if apply_ln else (None, None))
    _l.check(_l.load().hriemo_gate_blend(a.data_ptr(), a.stride(0), T_a, t.data_ptr(), t.stride(0), _ptr(ga),
                                          _ptr(ba), _ptr(gt), _ptr(bt), eps, int(apply_ln), w.data_ptr(),
                                          int(w_is_scalar), _ptr(hb), _ptr(hf), d, beta.data_ptr(), B, L, d,
                                          _ptr(pre_ln_a[0]) if pre_ln_a else None,
                                          _ptr(pre_ln_a[1]) if pre_ln_a else None,
                                          _ptr(pre_ln_t[0]) if pre_ln_t else None,
                                          _ptr(pre_ln_t[1]) if pre_ln_t else None,
                                          _ptr(pre_ln_a[2]) if pre_ln_a and len(pre_ln_a) > 2 else None,
                                          _ptr(pre_ln_t[2]) if pre_ln_t and len(pre_ln_t) > 2 else None,
                                          _stream()), "gate_blend")
    return hb, hf, beta


@_on_tensor_device
def mean_over_time(x: torch.Tensor, B: int, L: int) -> torch.Tensor:
    _chk2d(x, f32, "mean_over_time")
    d = x.shape[1]
    out = torch.empty((B, d), dtype=f32, device=x.device)
    _l.check(_l.load().hriemo_mean_over_time(x.data_ptr(), out.data_ptr(), B, L, d, _stream()), "mean_over_time")
    return out


@_on_tensor_device
def emotion_outputs(logits: torch.Tensor, thresholds: Optional[torch.Tensor] = None):
    """Post-path outputs of the reference's inference script (mosei_eval_infer.py:237-270):
    (probs = sigmoid(logits) fp32 [B, C], decisions = probs >= thresholds[c] as bool [B, C])."""
    _chk2d(logits, f32, "emotion_outputs logits")
    if not logits.is_contiguous():
        logits = logits.contiguous()
    B, Cn = logits.shape
    if thresholds is not None:
        _chk_f32(thresholds, (Cn,), "emotion_outputs thresholds")
    probs = torch.empty_like(logits)
    dec = torch.empty((B, Cn), dtype=torch.uint8, device=logits.device)
    _l.check(_l.load().hriemo_emotion_outputs(logits.data_ptr(), _ptr(thresholds), probs.data_ptr(), dec.data_ptr(),
                                               B, Cn, _stream()), "emotion_outputs")
    return probs, dec.view(torch.bool)


# ------------------------------------------------------------------ length-bucketed staging
@_on_tensor_device
def mask_lengths(pad: torch.Tensor) -> torch.Tensor:
    """[B, T] PAD mask -> int32 [B]: index of the last valid position + 1 (0 = everything is PAD)."""
    B, T = pad.shape
    m = _mask_u8(pad, B, T, "mask_lengths")
    out = torch.empty((B,), dtype=torch.int32, device=pad.device)
    _l.check(_l.load().hriemo_mask_lengths(m.data_ptr(), B, T, out.data_ptr(), _stream()), "mask_lengths")
    return out


def _chk_utt(utt: torch.Tensor, dev, name: str) -> None:
    if utt.device != dev or utt.dtype != torch.int32 or utt.dim() != 1 or not utt.is_contiguous():
        raise _l.HriemoError(f"{name}: utterance index must be a contiguous int32 vector on {dev}")


@_on_tensor_device
def gather_utterances(x: torch.Tensor, utt: torch.Tensor, T_out: int, ld_out: Optional[int] = None) -> torch.Tensor:
    """x [B, T, cols] (fp32 or bf16, CUDA) -> bf16 [n, T_out, ld_out]: utterances utt[i], time axis trimmed
    (or zero-extended) to T_out, columns zero-padded to ld_out."""
    if not x.is_cuda or x.dim() != 3 or x.dtype not in (f32, bf16):
        raise _l.HriemoError("gather_utterances: expected a CUDA [B, T, cols] fp32 / bf16 tensor")
    x = x.contiguous()
    _chk_utt(utt, x.device, "gather_utterances")
    B, T, cols = x.shape
    ld_out = round_up(cols, 8) if ld_out is None else ld_out
    n = utt.shape[0]
    out = torch.empty((n, T_out, ld_out), dtype=bf16, device=x.device)
    _l.check(_l.load().hriemo_gather_utterances_bf16(x.data_ptr(), 1 if x.dtype == f32 else 0, cols, T, utt.data_ptr(),
                                                      out.data_ptr(), ld_out, n, T_out, cols, _stream()),
             "gather_utterances_bf16")
    return out


@_on_tensor_device
def gather_masks(pad: torch.Tensor, utt: torch.Tensor, T_out: int) -> torch.Tensor:
    """[B, T] PAD mask -> bool [n, T_out] for utterances utt[i] (positions past T are PAD)."""
    B, T = pad.shape
    m = _mask_u8(pad, B, T, "gather_masks")
    _chk_utt(utt, pad.device, "gather_masks")
    n = utt.shape[0]
    out = torch.empty((n, T_out), dtype=torch.uint8, device=pad.device)
    _l.check(_l.load().hriemo_gather_masks(m.data_ptr(), T, utt.data_ptr(), out.data_ptr(), n, T_out, _stream()),
             "gather_masks")
    return out.view(torch.bool)


@_on_tensor_device
def scatter_rows(src: torch.Tensor, utt: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[utt[i]] = src[i] for fp32 tensors whose leading dim indexes utterances."""
    if src.dtype != f32 or dst.dtype != f32 or not src.is_cuda or not dst.is_contiguous() or dst.device != src.device:
        raise _l.HriemoError("scatter_rows: expected contiguous fp32 CUDA tensors on one device")
    src = src.contiguous()
    _chk_utt(utt, src.device, "scatter_rows")
    n = src.shape[0]
    cols = src.numel() // max(n, 1)
    if n != utt.shape[0] or (dst.numel() // max(dst.shape[0], 1)) != cols:
        raise _l.HriemoError(f"scatter_rows: shapes {tuple(src.shape)} -> {tuple(dst.shape)} with {utt.shape[0]} indices")
    _l.check(_l.load().hriemo_scatter_rows_f32(src.data_ptr(), utt.data_ptr(), dst.data_ptr(), n, cols, _stream()),
             "scatter_rows_f32")


def host_pack_bf16(src: torch.Tensor, dst: torch.Tensor, T_out: int, utt: Optional[torch.Tensor] = None,
                   lens: Optional[torch.Tensor] = None, n: Optional[int] = None, threads: int = 0) -> torch.Tensor:
    """HOST tensors only: dst[:n, :T_out, :] = bf16(src[utt[i], :T_out, :]) (rows past lens[i] zeroed), by C++
    threads (the GIL is released for the duration of the call).  src fp32 [B, T, cols] contiguous, dst a
    contiguous bf16 buffer with room for n * T_out * cols elements.  Returns the [n, T_out, cols] view."""
    import os
    if src.is_cuda or dst.is_cuda or src.dtype != f32 or dst.dtype != bf16 or src.dim() != 3:
        raise _l.HriemoError("host_pack_bf16: expected host tensors, fp32 [B, T, cols] -> bf16")
    if not src.is_contiguous() or not dst.is_contiguous():
        raise _l.HriemoError("host_pack_bf16: tensors must be contiguous")
    B, T, cols = src.shape
    n = (utt.shape[0] if utt is not None else B) if n is None else n
    for v, name in ((utt, "utt"), (lens, "lens")):
        if v is not None and (v.is_cuda or v.dtype != torch.int32 or not v.is_contiguous() or v.shape[0] < n):
            raise _l.HriemoError(f"host_pack_bf16: {name} must be a contiguous host int32 vector with >= n entries")
    if utt is None and n > B:
        raise _l.HriemoError("host_pack_bf16: n exceeds the batch")
    if dst.numel() < n * T_out * cols:
        raise _l.HriemoError("host_pack_bf16: destination too small")
    threads = threads or max(1, min(32, (os.cpu_count() or 1)))
    _l.check(_l.load().hriemo_host_pack_bf16(src.data_ptr(), cols, T, cols, _ptr(utt), _ptr(lens), dst.data_ptr(), cols,
                                              T_out, n, B, threads), "host_pack_bf16")
    return dst.view(-1)[: n * T_out * cols].view(n, T_out, cols)



# ------------------------------------------------------------------ loss and optimizer (training step)
@_on_tensor_device
def bce_beta_loss(logits: torch.Tensor, labels: torch.Tensor, beta: torch.Tensor, beta_weight: float = 0.01,
                  want_grads: bool = True):
    """BCEWithLogitsLoss(mean)(logits, labels) - beta_weight * mean(beta * (1 - beta)), the loss of
    train_fusion_seq_level_decoder.py:319-327.  Returns (loss [1], d_logits | None, d_beta | None), all fp32."""
    _chk2d(logits, f32, "bce_beta_loss logits")
    B, Cn = logits.shape
    logits = logits.contiguous()
    labels = labels.to(f32).contiguous()
    beta = beta.to(f32).contiguous().view(-1)
    if tuple(labels.shape) != (B, Cn) or beta.shape[0] != B or not labels.is_cuda or not beta.is_cuda:
        raise _l.HriemoError(f"bce_beta_loss: labels {tuple(labels.shape)} / beta {tuple(beta.shape)} do not match logits {(B, Cn)}")
    loss = torch.empty((1,), dtype=f32, device=logits.device)
    dl = torch.empty_like(logits) if want_grads else None
    db = torch.empty((B, 1), dtype=f32, device=logits.device) if want_grads else None
    _l.check(_l.load().hriemo_bce_beta_loss(logits.data_ptr(), labels.data_ptr(), beta.data_ptr(), beta_weight, B, Cn,
                                             loss.data_ptr(), _ptr(dl), _ptr(db), _stream()), "bce_beta_loss")
    return loss, dl, db


def _chk_flat(t: torch.Tensor, n: int, name: str) -> None:
    if not t.is_cuda or t.dtype != f32 or t.dim() != 1 or not t.is_contiguous() or t.shape[0] != n:
        raise _l.HriemoError(f"{name}: expected a contiguous flat CUDA fp32 arena of {n} elements")


@_on_tensor_device
def grad_norm_clip(grads: torch.Tensor, max_norm: float) -> torch.Tensor:
    """clip_grad_norm_ over a flat gradient arena without a host sync: returns a device tensor
    [total L2 norm, clip coefficient]; pass out[1:] to adamw_step as grad_scale."""
    _chk_flat(grads, grads.shape[0], "grad_norm_clip")
    ws = torch.empty((int(_l.load().hriemo_grad_norm_workspace_bytes()) // 8,), dtype=torch.float64, device=grads.device)
    out = torch.empty((2,), dtype=f32, device=grads.device)
    _l.check(_l.load().hriemo_grad_norm_clip(grads.data_ptr(), grads.shape[0], max_norm, ws.data_ptr(), out.data_ptr(),
                                              _stream()), "grad_norm_clip")
    return out


@_on_tensor_device
def adamw_step(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
               lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
               grad_scale: Optional[torch.Tensor] = None, params_bf16: Optional[torch.Tensor] = None) -> None:
    """In-place torch.optim.AdamW step over flat fp32 arenas (params, exp_avg, exp_avg_sq are updated)."""
    n = params.shape[0]
    for t, name in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _chk_flat(t, n, f"adamw_step {name}")
    if grad_scale is not None and (not grad_scale.is_cuda or grad_scale.dtype != f32 or grad_scale.numel() < 1):
        raise _l.HriemoError("adamw_step: grad_scale must be a CUDA fp32 tensor")
    if params_bf16 is not None and (params_bf16.dtype != bf16 or params_bf16.numel() != n or not params_bf16.is_contiguous()):
        raise _l.HriemoError("adamw_step: params_bf16 must be a contiguous bf16 tensor of the arena's size")
    _l.check(_l.load().hriemo_adamw_step(params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), n,
                                          step, lr, betas[0], betas[1], eps, weight_decay, _ptr(grad_scale),
                                          _ptr(params_bf16), _stream()), "adamw_step")


# ------------------------------------------------------------------ backward of a Linear layer
@_on_tensor_device
def transpose_bf16(w: torch.Tensor) -> torch.Tensor:
    """bf16 [rows, cols] -> [cols, rows] (the operand of dX = dY . W for the forward GEMM kernel)."""
    _chk2d(w, bf16, "transpose_bf16")
    rows, cols = w.shape
    out = torch.empty((cols, rows), dtype=bf16, device=w.device)
    _l.check(_l.load().hriemo_transpose_bf16(w.data_ptr(), w.stride(0), out.data_ptr(), rows, rows, cols, _stream()),
             "transpose_bf16")
    return out


@_on_tensor_device
def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: Optional[torch.Tensor] = None, db: Optional[torch.Tensor] = None,
                 want_bias: bool = True, accumulate: bool = False):
    """Gradients of y = x W^T + b w.r.t. W and b: dW [N, K] = dy^T x, db [N] = column sums of dy (fp32), on tcgen05
    tensor cores with both bf16 operands read as they are.  dw / db: existing fp32 tensors to write (or, with
    accumulate=True, add) into; otherwise fresh ones are returned."""
    _chk2d(dy, bf16, "linear_wgrad dy")
    _chk2d(x, bf16, "linear_wgrad x")
    M, N = dy.shape
    K = x.shape[1]
    if x.shape[0] != M:
        raise _l.HriemoError(f"linear_wgrad: row counts differ: {tuple(dy.shape)} vs {tuple(x.shape)}")
    if accumulate and dw is None:
        raise _l.HriemoError("linear_wgrad: accumulate=True needs dw")
    dw = torch.empty((N, K), dtype=f32, device=dy.device) if dw is None else dw
    _chk_f32(dw, (N, K), "linear_wgrad dw")
    if want_bias:
        if accumulate and db is None:
            raise _l.HriemoError("linear_wgrad: accumulate=True needs db when want_bias")
        db = torch.empty((N,), dtype=f32, device=dy.device) if db is None else db
        _chk_f32(db, (N,), "linear_wgrad db")
    else:
        db = None
    nbytes = int(_l.load().hriemo_linear_wgrad_workspace_bytes(M, N, K))
    if nbytes <= 0:
        raise _l.HriemoError(f"linear_wgrad: unsupported shape M={M} N={N} K={K} (N and K must be multiples of 128)")
    ws = torch.empty((nbytes // 4,), dtype=f32, device=dy.device)
    tok = _prof_begin("wgrad", 2.0 * M * N * K)
    _l.check(_l.load().hriemo_linear_wgrad_bf16(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), M, N, K,
                                                 dw.data_ptr(), _ptr(db), 1 if accumulate else 0, ws.data_ptr(), _stream()),
             "linear_wgrad_bf16")
    _prof_end(tok)
    return dw, db


def linear_backward(dy: torch.Tensor, x: torch.Tensor, w_t: torch.Tensor, want_bias: bool = True, relu_input: bool = False):
    """(dx bf16 [M, K], dW fp32 [N, K], db fp32 [N] | None) of y = x W^T + b given dy [M, N] (bf16), the layer's
    input x [M, K] (bf16) and the TRANSPOSED weight w_t = transpose_bf16(W) [K, N].  relu_input: x = relu(z) and the
    gradient wanted is dz = dx * (x > 0) -- the mask is applied in the GEMM's epilogue (EPI_BIAS_MASK)."""
    if relu_input:
        dx = gemm(dy, _as_gemm_weight(w_t, x.shape[1]), None, _l.EPI_BIAS_MASK, resid=x, tag="dgrad")
    else:
        dx = gemm(dy, _as_gemm_weight(w_t, x.shape[1]), None, _l.EPI_BIAS, tag="dgrad")
    dw, db = linear_wgrad(dy, x, want_bias=want_bias)
    return dx, dw, db


def _as_gemm_weight(w_t: torch.Tensor, k_in: int) -> torch.Tensor:
    """The forward GEMM computes A . B^T with B stored [out, contraction]; for dX = dY . W the output dimension is
    the layer's input width K and the contraction its output width N, i.e. B = W^T stored as [K, N]."""
    if w_t.shape[0] != k_in:
        raise _l.HriemoError(f"linear_backward: w_t must be the transposed weight [K={k_in}, N], got {tuple(w_t.shape)}")
    return w_t


@_on_tensor_device
def layernorm_backward(x: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, dgamma: Optional[torch.Tensor] = None,
                       dbeta: Optional[torch.Tensor] = None, accumulate: bool = False, eps: float = LN_EPS):
    """Backward of LayerNorm: x [rows, d] bf16 (the layer's input), dy [rows, d] bf16 -> (dx bf16, dgamma fp32, dbeta fp32)."""
    _chk2d(x, bf16, "layernorm_backward x")
    _chk2d(dy, bf16, "layernorm_backward dy")
    rows, d = x.shape
    if tuple(dy.shape) != (rows, d):
        raise _l.HriemoError("layernorm_backward: x and dy shapes differ")
    if accumulate and (dgamma is None or dbeta is None):
        raise _l.HriemoError("layernorm_backward: accumulate=True needs dgamma and dbeta")
    dgamma = torch.empty((d,), dtype=f32, device=x.device) if dgamma is None else dgamma
    dbeta = torch.empty((d,), dtype=f32, device=x.device) if dbeta is None else dbeta
    _chk_f32(dgamma, (d,), "layernorm_backward dgamma")
    _chk_f32(dbeta, (d,), "layernorm_backward dbeta")
    _chk_f32(gamma, (d,), "layernorm_backward gamma")
    dx = torch.empty((rows, d), dtype=bf16, device=x.device)
    ws = torch.empty((int(_l.load().hriemo_layernorm_backward_workspace_bytes(rows, d)) // 4,), dtype=f32, device=x.device)
    _l.check(_l.load().hriemo_layernorm_backward(x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), gamma.data_ptr(), eps,
                                                  dx.data_ptr(), d, dgamma.data_ptr(), dbeta.data_ptr(), 1 if accumulate else 0,
                                                  ws.data_ptr(), rows, d, _stream()), "layernorm_backward")
    return dx, dgamma, dbeta


@_on_tensor_device
def relu_backward(dy: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """dy where h > 0 else 0, h being the ReLU's output (bf16 [rows, cols])."""
    _chk2d(dy, bf16, "relu_backward dy")
    _chk2d(h, bf16, "relu_backward h")
    rows, cols = dy.shape
    if tuple(h.shape) != (rows, cols):
        raise _l.HriemoError("relu_backward: shapes differ")
    dx = torch.empty((rows, cols), dtype=bf16, device=dy.device)
    _l.check(_l.load().hriemo_relu_backward_bf16(dy.data_ptr(), dy.stride(0), h.data_ptr(), h.stride(0), dx.data_ptr(), cols,
                                                  rows, cols, _stream()), "relu_backward")
    return dx


@_on_tensor_device
def small_attention_backward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, d_out: torch.Tensor,
                             key_pad: Optional[torch.Tensor], B: int, H: int, Nq: int, Tk: int, dh: int, out=None, drop=None):
    """Backward of small_attention: (dq [B*Nq, H*dh], dk [B*Tk, H*dh], dv [B*Tk, H*dh]) in bf16.
    drop = (p8, scale, key): the forward's dropout on the probabilities (the same mask is recomputed).
    out = (dq, dk, dv): existing row-major bf16 views to write (e.g. the columns of one [rows, 3d] / [rows, 2d]
    buffer, so that the projection's backward sees one operand)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (d_out, "d_out")):
        _chk2d(t, bf16, f"small_attention_backward {n}")
    d = H * dh
    if out is None:
        dq = torch.empty((B * Nq, d), dtype=bf16, device=q.device)
        dk = torch.empty((B * Tk, d), dtype=bf16, device=q.device)
        dv = torch.empty((B * Tk, d), dtype=bf16, device=q.device)
    else:
        dq, dk, dv = out
        for t, n, rows in ((dq, "dq", B * Nq), (dk, "dk", B * Tk), (dv, "dv", B * Tk)):
            _chk2d(t, bf16, f"small_attention_backward {n}")
            if tuple(t.shape) != (rows, d):
                raise _l.HriemoError(f"small_attention_backward: {n} must be [{rows}, {d}], got {tuple(t.shape)}")
    m = _mask_u8(key_pad, B, Tk, "small_attention_backward")
    if drop is not None:
        _l.check(_l.load().hriemo_small_attention_backward_dropout(
            q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), d_out.data_ptr(), d_out.stride(0),
            _ptr(m), dq.data_ptr(), dq.stride(0), dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0), B, H, Nq, Tk, dh,
            1.0 / math.sqrt(dh), drop[0], drop[2], drop[1], _stream()),
            "small_attention_backward_dropout")
        return dq, dk, dv
    _l.check(_l.load().hriemo_small_attention_backward(
        q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), d_out.data_ptr(), d_out.stride(0),
        _ptr(m), dq.data_ptr(), dq.stride(0), dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0), B, H, Nq, Tk, dh,
        1.0 / math.sqrt(dh), _stream()),
        "small_attention_backward")
    return dq, dk, dv


@_on_tensor_device
def attention_backward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, d_out: torch.Tensor,
                       lse: torch.Tensor, key_pad: Optional[torch.Tensor], B: int, H: int, Tq: int, Tk: int, dh: int,
                       grads=None, impl: int = 0, drop=None):
    """Backward of `attention` (the encoder's attention): q / k / v as in the forward (column slices are fine), out
    [B*Tq, H*dh] and lse [B, H, Tq] from attention(..., want_lse=True), d_out [B*Tq, H*dh].
    -> (dq [B*Tq, H*dh], dk [B*Tk, H*dh], dv [B*Tk, H*dh]) bf16; grads = (dq, dk, dv): existing views to write into.
    impl: 0 (default) / 3 = the tcgen05 form (csrc/attention_bwd_tc.cu), 1 = fp32-FMA loops over mma.sync-sized tiles
    (slow; validation), 2 = the first mma.sync form, 4 = the ldmatrix mma.sync form (kept for A/B measurements)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out"), (d_out, "d_out")):
        _chk2d(t, bf16, f"attention_backward {n}")
    d = H * dh
    if q.shape[0] != B * Tq or k.shape[0] != B * Tk or v.shape[0] != B * Tk or tuple(out.shape) != (B * Tq, d) \
            or tuple(d_out.shape) != (B * Tq, d):
        raise _l.HriemoError("attention_backward: operand shapes do not match (B, H, Tq, Tk, dh)")
    _chk_f32(lse, (B, H, Tq), "attention_backward lse")
    if grads is None:
        dq = torch.empty((B * Tq, d), dtype=bf16, device=q.device)
        dk = torch.empty((B * Tk, d), dtype=bf16, device=q.device)
        dv = torch.empty((B * Tk, d), dtype=bf16, device=q.device)
    else:
        dq, dk, dv = grads
        for t, n, rows in ((dq, "dq", B * Tq), (dk, "dk", B * Tk), (dv, "dv", B * Tk)):
            _chk2d(t, bf16, f"attention_backward {n}")
            if tuple(t.shape) != (rows, d):
                raise _l.HriemoError(f"attention_backward: {n} must be [{rows}, {d}], got {tuple(t.shape)}")
    m = _mask_u8(key_pad, B, Tk, "attention_backward")
    dsum = torch.empty((B, H, Tq), dtype=f32, device=q.device)
    a = _l.AttnBwdArgs()
    a.q, a.ldq, a.k, a.ldk, a.v, a.ldv = q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0)
    a.key_pad = _ptr(m)
    a.out, a.ldo, a.d_out, a.lddo = out.data_ptr(), out.stride(0), d_out.data_ptr(), d_out.stride(0)
    a.lse, a.dsum = lse.data_ptr(), dsum.data_ptr()
    a.dq, a.lddq, a.dk, a.lddk, a.dv, a.lddv = dq.data_ptr(), dq.stride(0), dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0)
    a.B, a.H, a.Tq, a.Tk, a.dh = B, H, Tq, Tk, dh
    a.scale = 1.0 / math.sqrt(dh)
    a.impl = int(impl)
    if drop is not None:   # the forward's dropout on the probabilities: (p8, scale, key)
        a.drop_p8, a.drop_scale, a.drop_key = drop
    steps = None
    if m is not None and Tk > 64:
        steps = kv_steps(m)   # trailing all-PAD key tiles are skipped (ragged batches; exact: a masked key has P = 0)
        a.kv_steps = steps.data_ptr()
    tok = _prof_begin("attention_bwd", 14.0 * B * H * Tq * Tk * dh)   # 7 tile GEMMs (S and dP are formed in both passes)
    _l.check(_l.load().hriemo_attention_backward_bf16(C.byref(a), _stream()), "attention_backward_bf16")
    _prof_end(tok)
    return dq, dk, dv


# ------------------------------------------------------------------ backward of the fp32 gate / head
@_on_tensor_device
def linear_backward_f32(dy: torch.Tensor, x: Optional[torch.Tensor], w: Optional[torch.Tensor], want_dx: bool = True,
                        want_dw: bool = True, want_bias: bool = True, dw: Optional[torch.Tensor] = None,
                        db: Optional[torch.Tensor] = None, accumulate: bool = False):
    """Backward of sgemm's y = x W^T + b in fp32: (dx [M,K] | None, dW [N,K] | None, db [N] | None)."""
    _chk2d(dy, f32, "linear_backward_f32 dy")
    M, N = dy.shape
    if want_dx:
        _chk2d(w, f32, "linear_backward_f32 W")
        K = w.shape[1]
    if want_dw:
        _chk2d(x, f32, "linear_backward_f32 x")
        K = x.shape[1]
    if not (want_dx or want_dw):
        K = 1
    if (want_dx and tuple(w.shape) != (N, K)) or (want_dw and tuple(x.shape) != (M, K)):
        raise _l.HriemoError("linear_backward_f32: operand shapes do not match dy")
    if accumulate and ((want_dw and dw is None) or (want_bias and db is None)):
        raise _l.HriemoError("linear_backward_f32: accumulate=True needs dw / db")
    dx = torch.empty((M, K), dtype=f32, device=dy.device) if want_dx else None
    if want_dw:
        dw = torch.empty((N, K), dtype=f32, device=dy.device) if dw is None else dw
        _chk_f32(dw, (N, K), "linear_backward_f32 dw")
    else:
        dw = None
    if want_bias:
        db = torch.empty((N,), dtype=f32, device=dy.device) if db is None else db
        _chk_f32(db, (N,), "linear_backward_f32 db")
    else:
        db = None
    _l.check(_l.load().hriemo_linear_backward_f32(dy.data_ptr(), dy.stride(0), _ptr(x) if want_dw else None,
                                                   x.stride(0) if want_dw else 0, _ptr(w) if want_dx else None,
                                                   w.stride(0) if want_dx else 0, M, N, K, _ptr(dx), K, _ptr(dw), _ptr(db),
                                                   1 if accumulate else 0, _stream()), "linear_backward_f32")
    return dx, dw, db


@_on_tensor_device
def act_backward_f32(dy: torch.Tensor, y: torch.Tensor, act: int) -> torch.Tensor:
    """dy * act'(y) from the activation's output y (ACT_RELU / ACT_SIGMOID), contiguous fp32 tensors of one shape."""
    if dy.dtype != f32 or y.dtype != f32 or dy.shape != y.shape or not dy.is_contiguous() or not y.is_contiguous():
        raise _l.HriemoError("act_backward_f32: expected two contiguous fp32 tensors of one shape")
    dx = torch.empty_like(dy)
    _l.check(_l.load().hriemo_act_backward_f32(dy.data_ptr(), y.data_ptr(), dx.data_ptr(), dy.numel(), act, _stream()),
             "act_backward_f32")
    return dx


@_on_tensor_device
def sum_rows(x: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """Column sums of a [rows, cols] bf16 / fp32 matrix in fp32 (fixed summation order)."""
    if x.dtype not in (bf16, f32):
        raise _l.HriemoError(f"sum_rows: unsupported dtype {x.dtype}")
    _chk2d(x, x.dtype, "sum_rows x")
    rows, cols = x.shape
    if accumulate and out is None:
        raise _l.HriemoError("sum_rows: accumulate=True needs out")
    out = torch.empty((cols,), dtype=f32, device=x.device) if out is None else out
    _chk_f32(out, (cols,), "sum_rows out")
    _l.check(_l.load().hriemo_sum_rows(x.data_ptr(), int(x.dtype == f32), x.stride(0), out.data_ptr(), rows, cols,
                                        1 if accumulate else 0, _stream()), "sum_rows")
    return out


@_on_tensor_device
def gate_input_backward(dg: torch.Tensor, a_pool: torch.Tensor, t_pool: torch.Tensor):
    B, d = a_pool.shape
    _chk_f32(dg, (B, 4 * d), "gate_input_backward dg")
    _chk_f32(a_pool, (B, d), "gate_input_backward a_pool")
    _chk_f32(t_pool, (B, d), "gate_input_backward t_pool")
    da, dt = torch.empty_like(a_pool), torch.empty_like(t_pool)
    _l.check(_l.load().hriemo_gate_input_backward(dg.data_ptr(), a_pool.data_ptr(), t_pool.data_ptr(), da.data_ptr(),
                                                   dt.data_ptr(), B, d, _stream()), "gate_input_backward")
    return da, dt


@_on_tensor_device
def mask_inv_counts(like: torch.Tensor, pad: Optional[torch.Tensor], B: int, T: int) -> torch.Tensor:
    """1 / max(1, #non-PAD positions) per utterance (1 / T without a mask), fp32 [B] on the device of `like`."""
    m = _mask_u8(pad, B, T, "mask_inv_counts")
    out = torch.empty((B,), dtype=f32, device=like.device)
    _l.check(_l.load().hriemo_mask_inv_counts(_ptr(m), B, T, out.data_ptr(), _stream()), "mask_inv_counts")
    return out


@_on_tensor_device
def gate_blend_backward_w(dh: torch.Tensor, na: torch.Tensor, T_a: int, nt: torch.Tensor, dbeta: Optional[torch.Tensor],
                          B: int, L: int) -> torch.Tensor:
    """dw [B, d] fp32 of h = w*na[:, :L] + (1-w)*nt, beta = mean_d(w), from dh [B*L, d] and dbeta [B, 1]."""
    for t, n in ((dh, "dh"), (na, "na"), (nt, "nt")):
        _chk2d(t, bf16, f"gate_blend_backward_w {n}")
    d = dh.shape[1]
    if dh.shape[0] != B * L or nt.shape[0] != B * L or na.shape[0] != B * T_a:
        raise _l.HriemoError("gate_blend_backward_w: row counts do not match (B, L, T_a)")
    if dbeta is not None:
        _chk_f32(dbeta.view(-1), (B,), "gate_blend_backward_w dbeta")
    dw = torch.empty((B, d), dtype=f32, device=dh.device)
    _l.check(_l.load().hriemo_gate_blend_backward_w(dh.data_ptr(), dh.stride(0), na.data_ptr(), na.stride(0), T_a,
                                                     nt.data_ptr(), nt.stride(0), _ptr(dbeta), dw.data_ptr(), B, L, d,
                                                     _stream()), "gate_blend_backward_w")
    return dw


@_on_tensor_device
def gate_stream_grad(dh: torch.Tensor, L: int, w: torch.Tensor, one_minus: bool, dpool: torch.Tensor,
                     pad: Optional[torch.Tensor], inv_counts: torch.Tensor, B: int, T: int) -> torch.Tensor:
    """Gradient w.r.t. a LayerNorm-ed gate stream [B*T, d] (bf16): blend share on the first L rows + pooled-mean share."""
    _chk2d(dh, bf16, "gate_stream_grad dh")
    d = dh.shape[1]
    _chk_f32(w, (B, d), "gate_stream_grad w")
    _chk_f32(dpool, (B, d), "gate_stream_grad dpool")
    _chk_f32(inv_counts, (B,), "gate_stream_grad inv_counts")
    m = _mask_u8(pad, B, T, "gate_stream_grad")
    dn = torch.empty((B * T, d), dtype=bf16, device=dh.device)
    _l.check(_l.load().hriemo_gate_stream_grad(dh.data_ptr(), dh.stride(0), L, w.data_ptr(), int(one_minus),
                                                dpool.data_ptr(), _ptr(m), inv_counts.data_ptr(), dn.data_ptr(), d, B, T, d,
                                                _stream()), "gate_stream_grad")
    return dn


# ------------------------------------------------------------------ "tf32-class" precision mode (hriemo/precise.py)
@_on_tensor_device
def split3(x: torch.Tensor, weight: bool = False, relu: bool = False) -> torch.Tensor:
    """fp32 [rows, K] -> bf16 [rows, 3 * roundup(K, 8)]: activations as [hi | lo | hi], weights as [hi | hi | lo]
    (hi = bf16(x), lo = bf16(x - hi)); one bf16 GEMM over the tripled K then carries hi.hi + lo.hi + hi.lo."""
    _chk2d(x, f32, "split3 x")
    rows, K = x.shape
    Kp = round_up(K, 8)
    out = torch.empty((rows, 3 * Kp), dtype=bf16, device=x.device)
    _l.check(_l.load().hriemo_split3(x.data_ptr(), x.stride(0), out.data_ptr(), 3 * Kp, rows, K, int(weight), int(relu),
                                      _stream()), "split3")
    return out


@_on_tensor_device
def attention_f32(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, key_pad: Optional[torch.Tensor], B: int, H: int,
                  Tq: int, Tk: int, dh: int, want_probs: bool = False):
    """fp32 attention on the CUDA cores; q [B*Tq, >= H*dh], k / v [B*Tk, >= H*dh] views.  Returns (out f32
    [B*Tq, H*dh], head-averaged probabilities f32 [B, Tq, Tk] | None)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk2d(t, f32, f"attention_f32 {n}")
    if q.shape[0] != B * Tq or k.shape[0] != B * Tk or v.shape[0] != B * Tk:
        raise _l.HriemoError("attention_f32: row counts do not match B*Tq / B*Tk")
    out = torch.empty((B * Tq, H * dh), dtype=f32, device=q.device)
    probs = torch.zeros((B, Tq, Tk), dtype=f32, device=q.device) if want_probs else None
    m = _mask_u8(key_pad, B, Tk, "attention_f32")
    _l.check(_l.load().hriemo_attention_f32(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(),
                                             v.stride(0), _ptr(m), out.data_ptr(), out.stride(0), _ptr(probs), B, H, Tq,
                                             Tk, dh, 1.0 / math.sqrt(dh), _stream()), "attention_f32")
    return out, probs


@_on_tensor_device
def masked_mean_f32(x: torch.Tensor, pad: Optional[torch.Tensor], B: int, T: int) -> torch.Tensor:
    _chk2d(x, f32, "masked_mean_f32 x")
    if not x.is_contiguous() or x.shape[0] != B * T:
        raise _l.HriemoError("masked_mean_f32: expected a contiguous [B*T, d] tensor")
    d = x.shape[1]
    pooled = torch.empty((B, d), dtype=f32, device=x.device)
    m = _mask_u8(pad, B, T, "masked_mean_f32")
    _l.check(_l.load().hriemo_masked_mean_f32(x.data_ptr(), _ptr(m), pooled.data_ptr(), B, T, d, _stream()),
             "masked_mean_f32")
    return pooled


@_on_tensor_device
def gate_blend_f32(a: torch.Tensor, T_a: int, t: torch.Tensor, w: torch.Tensor, B: int, L: int):
    """h = w * a[:, :L] + (1 - w) * t on contiguous f32 [B*T_a, d] / [B*L, d] tensors, w [B, d]; returns (h, beta [B, 1])."""
    for x, n in ((a, "a"), (t, "t"), (w, "w")):
        _chk2d(x, f32, f"gate_blend_f32 {n}")
        if not x.is_contiguous():
            raise _l.HriemoError(f"gate_blend_f32 {n}: expected a contiguous tensor")
    d = a.shape[1]
    h = torch.empty((B * L, d), dtype=f32, device=a.device)
    beta = torch.empty((B, 1), dtype=f32, device=a.device)
    _l.check(_l.load().hriemo_gate_blend_f32(a.data_ptr(), T_a, t.data_ptr(), w.data_ptr(), h.data_ptr(), beta.data_ptr(),
                                              B, L, d, _stream()), "gate_blend_f32")
    return h, beta


# ------------------------------------------------------------------ dropout of the training step (hriemo/dropout.py)
@_on_tensor_device
def dropout(x: torch.Tensor, drop, resid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = (keep ? x * scale : 0) [+ resid]; x / resid / out all bf16 or all f32 [rows, cols]; drop = (p8, scale, key).
    Called on a sub-layer's output with the residual in the forward and on the gradient (no residual) in the backward."""
    if x.dtype not in (bf16, f32):
        raise _l.HriemoError(f"dropout: unsupported dtype {x.dtype}")
    _chk2d(x, x.dtype, "dropout x")
    if resid is not None:
        _chk2d(resid, x.dtype, "dropout resid")
        if resid.shape != x.shape:
            raise _l.HriemoError(f"dropout: residual shape {tuple(resid.shape)} != {tuple(x.shape)}")
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=x.dtype, device=x.device)
    p8, scale, key = drop
    _l.check(_l.load().hriemo_dropout(x.data_ptr(), int(x.dtype == f32), x.stride(0), _ptr(resid),
                                       resid.stride(0) if resid is not None else 0, out.data_ptr(), out.stride(0), rows, cols,
                                       p8, scale, key, _stream()), "dropout")
    return out


def dropout_mask(rows: int, cols: int, key: int, p8: int, device, rows_per_stream: int = 0) -> torch.Tensor:
    """The keep mask as bool [rows, cols] (tests; rows_per_stream > 0: the (utterance, head) streams of an attention)."""
    out = torch.empty((rows, cols), dtype=torch.uint8, device=device)
    with torch.cuda.device(out.device):
        _l.check(_l.load().hriemo_dropout_mask(out.data_ptr(), rows, cols, key, p8, rows_per_stream, _stream()), "dropout_mask")
    return out.view(torch.bool)
