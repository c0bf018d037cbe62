"""Discrete-event model of forward_from_host: one copy engine, one host-cast worker, one compute stream,
two fp32 and two bf16 device staging sets, two pinned bf16 host sets.  Used to pick the slab plan
(pipeline.slab_schedule); rates in utterances per ms at the north-star shape."""
import sys


def simulate(plan, C=36.0, P32=32.1, H=40.0, fixed=0.35):
    """plan = [(n, host_cast)]; returns makespan in ms.  fixed = per-slab launch/latency overhead (ms)."""
    P16 = 2 * P32
    copy_free = 0.0
    comp_free = 0.0
    worker_free = 0.0
    host_sent = []      # copy-end times of host-cast slabs (host buffer reuse)
    dev32_cons, dev16_cons = [], []
    for n, hc in plan:
        if hc:
            k = len(host_sent)
            start = max(worker_free, host_sent[k - 2] if k >= 2 else 0.0)
            ready = start + n / H
            worker_free = ready
            buf_free = dev16_cons[k - 2] if k >= 2 else 0.0
            c0 = max(copy_free, ready, buf_free)
            c1 = c0 + n / P16
            host_sent.append(c1)
        else:
            k = len(dev32_cons)
            buf_free = dev32_cons[k - 2] if k >= 2 else 0.0
            c0 = max(copy_free, buf_free)
            c1 = c0 + n / P32
        copy_free = c1
        k0 = max(comp_free, c1)
        k1 = k0 + n / C + fixed
        comp_free = k1
        if hc:
            dev16_cons.append(k1)
        else:
            dev32_cons.append(k0 + 0.05 * n / C)   # the cast kernel reads the staging set first
    return comp_free


def plans(B=4096):
    out = {}
    out["v10: 8x512 alt"] = [(512, i % 2 == 1) for i in range(8)]
    out["ramp 64,128,256 + 512 alt"] = [(64, 0), (128, 0), (256, 0)] + [(512, i % 2 == 1) for i in range(7)] + [(64, 1)]
    out["16x256 alt"] = [(256, i % 2 == 1) for i in range(16)]
    out["64f,128h,256h, then 256 alt"] = [(64, 0), (128, 1), (192, 1)] + [(256, i % 2 == 0) for i in range(14)] + [(128, 0)]
    out["64f,64f then 256: h h f ..."] = [(64, 0), (64, 0), (128, 1)] + [(256, i % 3 != 2) for i in range(15)]
    out["64f,128f, 256 (2 of 3 host)"] = [(64, 0), (128, 0)] + [(256, i % 3 != 2) for i in range(15)] + [(64, 0)]
    out["all host 256"] = [(64, 0), (128, 0)] + [(256, 1) for i in range(15)] + [(64, 0)]
    for k, p in out.items():
        assert sum(n for n, _ in p) == B, (k, sum(n for n, _ in p))
    return out


if __name__ == "__main__":
    for H in (16.0, 25.0, 40.0):
        for C in (34.3, 36.9):
            print(f"--- host cast {H} utt/ms, compute {C} utt/ms (ideal {4096 / C:.1f} ms)")
            for k, p in plans().items():
                print(f"  {k:34s} {simulate(p, C=C, H=H):7.1f} ms")
