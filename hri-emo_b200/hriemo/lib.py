"""ctypes binding of libhriemo_b200.so (C ABI declared in include/hriemo.h).

There is no fallback: if the shared library is missing or a call fails, an
exception is raised.  PyTorch is only used for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HRIEMO_LIB_PATH: an alternative build of the same library (tools/ A/B experiments only)
LIB_PATH = os.environ.get("HRIEMO_LIB_PATH") or os.path.join(_HERE, "libhriemo_b200.so")

EPI_BIAS, EPI_BIAS_RELU, EPI_BIAS_RESID, EPI_BIAS_RESID_F32, _EPI_RETIRED, EPI_BIAS_F32, EPI_BIAS_MASK = range(7)
ACT_NONE, ACT_RELU, ACT_SIGMOID = range(3)
ACT_RELU_IN = 4   # or-ed in: ReLU on the sgemm's A operand as it is read


class HriemoError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("W", C.c_void_p), ("ldw", C.c_int64),
        ("bias", C.c_void_p),
        ("M", C.c_int64), ("N", C.c_int32), ("K", C.c_int32),
        ("epilogue", C.c_int32), ("cta_pair", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("a_stats", C.c_void_p), ("a_colsum", C.c_void_p),
        ("resid_stats", C.c_void_p), ("resid_gamma", C.c_void_p), ("resid_beta", C.c_void_p),
        ("stats_out", C.c_void_p),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int64),
        ("k", C.c_void_p), ("ldk", C.c_int64),
        ("v", C.c_void_p), ("ldv", C.c_int64),
        ("key_pad", C.c_void_p),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Tq", C.c_int32), ("Tk", C.c_int32), ("dh", C.c_int32),
        ("scale", C.c_float),
        ("kv_steps", C.c_void_p),
        ("no_head_pairs", C.c_int32),
        ("lse", C.c_void_p),
        ("drop_p8", C.c_uint32), ("drop_key", C.c_uint32), ("drop_scale", C.c_float),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int64),
        ("k", C.c_void_p), ("ldk", C.c_int64),
        ("v", C.c_void_p), ("ldv", C.c_int64),
        ("key_pad", C.c_void_p),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("d_out", C.c_void_p), ("lddo", C.c_int64),
        ("lse", C.c_void_p), ("dsum", C.c_void_p),
        ("dq", C.c_void_p), ("lddq", C.c_int64),
        ("dk", C.c_void_p), ("lddk", C.c_int64),
        ("dv", C.c_void_p), ("lddv", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Tq", C.c_int32), ("Tk", C.c_int32), ("dh", C.c_int32),
        ("scale", C.c_float),
        ("impl", C.c_int32),
        ("kv_steps", C.c_void_p),
        ("drop_p8", C.c_uint32), ("drop_key", C.c_uint32), ("drop_scale", C.c_float),
    ]


class ShardInfo(C.Structure):
    _fields_ = [("n_utt", C.c_int64), ("rows_a", C.c_int64), ("rows_t", C.c_int64), ("meta_bytes", C.c_int64),
                ("d_a", C.c_int32), ("d_t", C.c_int32), ("dtype", C.c_int32),
                ("max_len_a", C.c_int32), ("max_len_t", C.c_int32)]


_P, _I64, _I32, _F = C.c_void_p, C.c_int64, C.c_int32, C.c_float

# name -> (restype, argtypes); must list every symbol include/hriemo.h declares
SIGNATURES = {
    "hriemo_version": (C.c_int, []),
    "hriemo_last_error": (C.c_char_p, []),
    "hriemo_launch_count": (C.c_int64, []),
    "hriemo_cast_f32_to_bf16": (C.c_int, [_P, _I64, _P, _I64, _I64, _I32, _P]),
    "hriemo_gemm_bf16": (C.c_int, [C.POINTER(GemmArgs), _P]),
    "hriemo_attention_bf16": (C.c_int, [C.POINTER(AttnArgs), _P]),
    "hriemo_attention_kv_steps": (C.c_int, [_P, _I32, _I32, _P, _P]),
    "hriemo_attention_probs": (C.c_int, [_P, _I64, _P, _I64, _P, _P, _I32, _I32, _I32, _I32, _I32, _F, _P]),
    "hriemo_small_attention": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _P, _I64, _P,
                                          _I32, _I32, _I32, _I32, _I32, _F, _P]),
    "hriemo_layernorm": (C.c_int, [_P, _I32, _I64, _P, _P, _F, _P, _P, _I64, _I64, _I32, _P]),
    "hriemo_ln_masked_mean": (C.c_int, [_P, _I64, _P, _P, _F, _I32, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P]),
    "hriemo_ln_stats_finalize": (C.c_int, [_P, _I32, _I64, _I32, _F, _P, _P]),
    "hriemo_fold_ln_weight": (C.c_int, [_P, _I64, _P, _P, _P, _P, _I64, _P, _P, _I32, _I32, _P]),
    "hriemo_gate_input": (C.c_int, [_P, _P, _P, _I32, _I32, _P]),
    "hriemo_sgemm_f32": (C.c_int, [_P, _I64, _P, _I64, _P, _P, _I64, _I64, _I32, _I32, _I32, _P]),
    "hriemo_gate_blend": (C.c_int, [_P, _I64, _I32, _P, _I64, _P, _P, _P, _P, _F, _I32, _P, _I32,
                                     _P, _P, _I64, _P, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "hriemo_mean_over_time": (C.c_int, [_P, _P, _I32, _I32, _I32, _P]),
    "hriemo_emotion_outputs": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P]),
    "hriemo_mask_lengths": (C.c_int, [_P, _I32, _I32, _P, _P]),
    "hriemo_gather_utterances_bf16": (C.c_int, [_P, _I32, _I64, _I32, _P, _P, _I64, _I32, _I32, _I32, _P]),
    "hriemo_gather_masks": (C.c_int, [_P, _I32, _P, _P, _I32, _I32, _P]),
    "hriemo_scatter_rows_f32": (C.c_int, [_P, _P, _P, _I64, _I64, _P]),
    "hriemo_shard_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "hriemo_shard_info": (C.c_int, [_P, C.POINTER(ShardInfo)]),
    "hriemo_shard_lengths": (C.c_int, [_P, _P, _P]),
    "hriemo_shard_meta": (C.c_int, [_P, _P, _I64]),
    "hriemo_shard_read": (C.c_int, [_P, _P, _I64, _I64, _I32, _I32, _P, _P, _P, _P, _I32]),
    "hriemo_shard_close": (C.c_int, [_P]),
    "hriemo_bce_beta_loss": (C.c_int, [_P, _P, _P, _F, _I64, _I32, _P, _P, _P, _P]),
    "hriemo_grad_norm_workspace_bytes": (C.c_int64, []),
    "hriemo_grad_norm_clip": (C.c_int, [_P, _I64, _F, _P, _P, _P]),
    "hriemo_adamw_step": (C.c_int, [_P, _P, _P, _P, _I64, _I32, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.c_double, _P, _P, _P]),
    "hriemo_linear_wgrad_workspace_bytes": (C.c_int64, [_I64, _I32, _I32]),
    "hriemo_linear_wgrad_bf16": (C.c_int, [_P, _I64, _P, _I64, _I64, _I32, _I32, _P, _P, _I32, _P, _P]),
    "hriemo_transpose_bf16": (C.c_int, [_P, _I64, _P, _I64, _I32, _I32, _P]),
    "hriemo_layernorm_backward_workspace_bytes": (C.c_int64, [_I64, _I32]),
    "hriemo_layernorm_backward": (C.c_int, [_P, _I64, _P, _I64, _P, _F, _P, _I64, _P, _P, _I32, _P, _I64, _I32, _P]),
    "hriemo_relu_backward_bf16": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _I64, _I32, _P]),
    "hriemo_small_attention_backward": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _P, _I64, _P, _I64, _P, _I64,
                                                  _I32, _I32, _I32, _I32, _I32, _F, _P]),
    "hriemo_linear_backward_f32": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _I64, _I32, _I32, _P, _I64, _P, _P, _I32, _P]),
    "hriemo_act_backward_f32": (C.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "hriemo_sum_rows": (C.c_int, [_P, _I32, _I64, _P, _I64, _I32, _I32, _P]),
    "hriemo_gate_input_backward": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _P]),
    "hriemo_mask_inv_counts": (C.c_int, [_P, _I32, _I32, _P, _P]),
    "hriemo_gate_blend_backward_w": (C.c_int, [_P, _I64, _P, _I64, _I32, _P, _I64, _P, _P, _I32, _I32, _I32, _P]),
    "hriemo_gate_stream_grad": (C.c_int, [_P, _I64, _I32, _P, _I32, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P]),
    "hriemo_attention_backward_bf16": (C.c_int, [C.POINTER(AttnBwdArgs), _P]),
    "hriemo_split3": (C.c_int, [_P, _I64, _P, _I64, _I64, _I32, _I32, _I32, _P]),
    "hriemo_attention_f32": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _P, _I64, _P, _I32, _I32, _I32, _I32, _I32, _F, _P]),
    "hriemo_masked_mean_f32": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P]),
    "hriemo_gate_blend_f32": (C.c_int, [_P, _I32, _P, _P, _P, _P, _I32, _I32, _I32, _P]),
    "hriemo_small_attention_dropout": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _F,
                                                  C.c_uint32, C.c_uint32, _F, _P]),
    "hriemo_small_attention_backward_dropout": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _P, _I64, _P, _I64, _P, _I64,
                                                           _I32, _I32, _I32, _I32, _I32, _F, C.c_uint32, C.c_uint32, _F, _P]),
    "hriemo_dropout": (C.c_int, [_P, _I32, _I64, _P, _I64, _P, _I64, _I64, _I32, C.c_uint32, _F, C.c_uint32, _P]),
    "hriemo_dropout_mask": (C.c_int, [_P, _I64, _I32, C.c_uint32, C.c_uint32, _I64, _P]),
    "hriemo_host_pack_bf16": (C.c_int, [_P, _I64, _I64, _I64, _P, _P, _P, _I64, _I64, _I64, _I64, _I32]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HriemoError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for the HRI-EMO B200 kernels."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("HRIEMO_LIB_PATH") and not hasattr(lib, name):
            continue  # an older A/B build (tools/) may predate an entry point; the in-tree library never may
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().hriemo_last_error()
        raise HriemoError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().hriemo_launch_count())
