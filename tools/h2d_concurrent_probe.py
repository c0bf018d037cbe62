#!/usr/bin/env python3
"""Pinned host -> device copy bandwidth of every GPU of a box, one rank at a time and all ranks at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_concurrent_probe.py

Prints one JSON line on rank 0: per-GPU GB/s alone, per-GPU GB/s with all N ranks copying concurrently, and the
aggregates.  This is the measurement behind DESIGN sec. 7's statement that the end-to-end number at N = 4 / 8 is bound
by PCIe uplinks shared between GPUs (the fp32 features are 7.1 GB per rank per step)."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_bytes = 1 << 30
    host = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    devb = torch.empty(n_bytes, dtype=torch.uint8, device=dev)

    def copy_gbs(reps=8):
        devb.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            devb.copy_(host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return reps * n_bytes / (e0.elapsed_time(e1) * 1e-3) / 1e9

    def barrier():
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    alone = torch.zeros(world, device=dev)
    for r in range(world):
        barrier()
        if r == rank:
            alone[r] = copy_gbs()
    barrier()
    t0 = time.perf_counter()
    together = torch.zeros(world, device=dev)
    together[rank] = copy_gbs(16)
    wall = time.perf_counter() - t0
    # pairs: ranks (2k, 2k+1) together, the others idle -> do two neighbours share an uplink?
    pairs = torch.zeros(world, device=dev)
    for k in range(0, world, 2):
        barrier()
        if rank in (k, k + 1):
            pairs[rank] = copy_gbs()
    barrier()
    if world > 1:
        for t in (alone, together, pairs):
            dist.all_reduce(t)
    if rank == 0:
        print(json.dumps({"what": "pinned H2D copy bandwidth, GB/s", "n_gpus": world, "bytes_per_copy": n_bytes,
                          "alone_per_gpu": [round(x, 1) for x in alone.tolist()],
                          "all_ranks_concurrent_per_gpu": [round(x, 1) for x in together.tolist()],
                          "all_ranks_concurrent_aggregate": round(float(together.sum()), 1),
                          "neighbour_pairs_concurrent_per_gpu": [round(x, 1) for x in pairs.tolist()],
                          "host_cpus": os.cpu_count()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
