"""Builds a -DHRIEMO_ATTN_TRACE copy of the library and prints per-step pipeline timings of CTA 0."""
import ctypes, os, subprocess, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import build as B, lib as L
TRACE_SO = os.path.join(ROOT, "tools", "_build", "libhriemo_trace.so")

def build_trace():
    os.makedirs(os.path.dirname(TRACE_SO), exist_ok=True)
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    cmd = [B._nvcc(), *B.NVCC_FLAGS, "-DHRIEMO_ATTN_TRACE", *extra, *[os.path.join(B.CSRC, s) for s in B.SOURCES], "-o", TRACE_SO]
    subprocess.run(cmd, check=True)

def main():
    if "--build-only" in sys.argv:
        return build_trace()
    L.LIB_PATH = TRACE_SO
    lib = L.load()
    from hriemo import ops
    Bsz, H, Tq, Tk, dh = (int(x) for x in (sys.argv[1:6] if len(sys.argv) >= 6 else (64, 8, 500, 500, 96)))
    d = H * dh
    dev = "cuda"
    q = torch.randn(Bsz * Tq, d, device=dev).bfloat16(); k = torch.randn(Bsz * Tk, d, device=dev).bfloat16()
    vt = torch.randn(Bsz * Tk, d, device=dev).bfloat16()
    for _ in range(3): ops.attention(q, k, vt, None, Bsz, H, Tq, Tk, dh)
    trace = torch.zeros(4 * 64 * 8, dtype=torch.int64, device=dev)
    fn = lib.hriemo_debug_set_attn_trace; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p]
    assert fn(trace.data_ptr()) == 0
    ops.attention(q, k, vt, None, Bsz, H, Tq, Tk, dh)
    torch.cuda.synchronize()
    fn(None)
    t = trace.view(4, 64, 8).cpu()
    t0 = int(t[t > 0].min())
    names = {0: "MMA0", 1: "MMA1", 2: "WG0", 3: "WG1"}
    ev = {0: ["top", "vfull_ok", "pfull_ok", "PV_issued", "S_top", "kfull_ok", "S_issued"],
          2: ["pre_sfull", "sfull_ok", "S_loaded", "max_done", "exp_done", "st_done", "handed", "epi_done"]}
    for role in (0, 2, 3):
        print(f"== {names[role]} (cycles since the first event; [delta to the previous event of the step])")
        labels = ev[0] if role < 2 else ev[2]
        for step in range(4, 28):
            row = t[role, step]
            if int(row.max()) == 0: continue
            pairs = sorted((int(x) - t0, lab) for lab, x in zip(labels, row) if int(x) > 0)
            parts, last = [], None
            for v, lab in pairs:
                parts.append(f"{lab}={v}" + (f"[+{v - last}]" if last is not None else ""))
                last = v
            print(f"  step {step:2d}: " + " ".join(parts))

if __name__ == "__main__":
    main()
