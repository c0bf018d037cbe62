"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/hriemo.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from hriemo import build, lib as L

    build.build()
    return L.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hriemo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hriemo_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/hriemo.h but not exported"


def test_binding_covers_header():
    from hriemo import lib as L

    assert sorted(L.SIGNATURES) == _declared_symbols()


def test_version_and_error_text(lib):
    assert lib.hriemo_version() == 100
    assert isinstance(lib.hriemo_last_error(), bytes)
    assert lib.hriemo_launch_count() == 0


def test_struct_layout_matches_header(tmp_path):
    """ctypes Structures must have the layout gcc gives the C declarations."""
    import subprocess
    from hriemo import lib as L

    fields = {"hriemo_gemm_args": L.GemmArgs, "hriemo_attn_args": L.AttnArgs, "hriemo_attn_bwd_args": L.AttnBwdArgs,
              "hriemo_shard_info_t": L.ShardInfo}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "hriemo.h"', "int main(void){"]
    for cname, st in fields.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, st in fields.items():
        assert int(out[cname]) == ctypes.sizeof(st)
        for fname, _ in st._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(st, fname).offset, (cname, fname)


def test_argument_validation_without_gpu(lib):
    """Validation runs before any CUDA call, so bad arguments are reported on a CPU box."""
    from hriemo import lib as L

    a = L.GemmArgs()
    rc = lib.hriemo_gemm_bf16(ctypes.byref(a), None)
    assert rc == -1 and b"null operand" in lib.hriemo_last_error()
    a.A, a.W, a.out = 16, 16, 16
    a.M, a.N, a.K, a.lda, a.ldw, a.ldo = 128, 100, 64, 64, 64, 104
    rc = lib.hriemo_gemm_bf16(ctypes.byref(a), None)
    assert rc == -1 and b"multiple of 32" in lib.hriemo_last_error()
    t = L.AttnArgs()
    t.q, t.k, t.v, t.out = 16, 16, 16, 16
    t.B, t.H, t.Tq, t.Tk, t.dh, t.ldq, t.ldk, t.ldv, t.ldo = 1, 1, 8, 8, 48, 48, 48, 48, 48
    t.scale = 0.125
    rc = lib.hriemo_attention_bf16(ctypes.byref(t), None)
    assert rc == -1 and b"head dim" in lib.hriemo_last_error()


def test_hard_limits_are_reported_not_crashed(lib):
    """The documented limits of the path come back as error codes with a message: key sequences
    too long for the shared-memory key caps, rows wider than 2048, unknown GEMM epilogues, fp32 attention key rows."""
    from hriemo import lib as L

    t = L.AttnArgs()
    t.q, t.k, t.v, t.out = 16, 16, 16, 16
    t.B, t.H, t.Tq, t.Tk, t.dh, t.ldq, t.ldk, t.ldv, t.ldo = 1, 8, 64, 40000, 96, 768, 768, 768, 768
    t.scale = 0.1
    assert lib.hriemo_attention_bf16(ctypes.byref(t), None) == -1 and b"too long" in lib.hriemo_last_error()
    rc = lib.hriemo_layernorm(16, 0, 4096, 16, 16, 1e-5, 16, None, 4096, 8, 4096, None)
    assert rc == -1 and b"<= 2048" in lib.hriemo_last_error()
    a = L.GemmArgs()
    a.A, a.W, a.out = 16, 16, 16
    a.M, a.N, a.K, a.lda, a.ldw, a.ldo, a.epilogue = 128, 128, 64, 64, 64, 128, 4
    assert lib.hriemo_gemm_bf16(ctypes.byref(a), None) == -1 and b"unknown epilogue" in lib.hriemo_last_error()
    a.epilogue = L.EPI_BIAS_MASK   # the ReLU-mask epilogue needs the post-ReLU tensor
    assert lib.hriemo_gemm_bf16(ctypes.byref(a), None) == -1 and b"without resid" in lib.hriemo_last_error()
    rc = lib.hriemo_attention_f32(16, 768, 16, 768, 16, 768, None, 16, 768, None, 1, 8, 4, 100000, 96, 0.1, None)
    assert rc == -1 and b"too long" in lib.hriemo_last_error()
    rc = lib.hriemo_split3(16, 64, 16, 100, 4, 64, 0, 0, None)
    assert rc == -1 and b"3 * roundup" in lib.hriemo_last_error()


def test_ops_refuse_cpu_tensors():
    import torch
    from hriemo import lib as L, ops

    with pytest.raises(L.HriemoError, match="CUDA tensor"):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(32, 8, dtype=torch.bfloat16), None)
