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


def test_ops_refuse_cpu_tensors():
    import torch
    from hriemo import lib as L, ops

    with pytest.raises(L.HriemoError, match="CUDA tensor"):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(32, 8, dtype=torch.bfloat16), None)
