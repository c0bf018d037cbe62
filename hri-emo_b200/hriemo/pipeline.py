"""Host staging and batch sharding around the forward path.

* ``forward_from_host``: the end-to-end call for features that live in (pinned) host
  memory, as the reference's DataLoader/collate delivers them
  (scripts/fusion/train_fusion_seq_level_decoder.py:191-232, :306-308).  The batch is cut
  into slabs of utterances; slab i+1 is copied host->device on a side stream while slab i
  computes, and only logits / beta / z come back.
* ``shard_bounds`` / ``gather_outputs``: batch sharding over ranks (utterances are
  independent); the only collective is one all_gather of the small outputs.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import engine as E


def shard_bounds(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of the batch: rank r gets [lo, hi); the first B % world ranks get one extra."""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_outputs(logits: torch.Tensor, beta: torch.Tensor, group=None):
    """all_gather of per-rank [b, N_e] logits and [b, 1] beta (equal b on every rank) as ONE
    collective on a packed [b, N_e + 1] buffer.  Returns ([B, N_e], [B, 1])."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    packed = torch.cat([logits, beta], dim=1).contiguous()
    out = torch.empty((world * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    n_e = logits.shape[1]
    return out[:, :n_e], out[:, n_e:]


@torch.no_grad()
def forward_from_host(model, h_a: torch.Tensor, h_t: torch.Tensor, mask_a: Optional[torch.Tensor] = None,
                      mask_t: Optional[torch.Tensor] = None, device="cuda", slab: int = 512,
                      out_device="cpu"):
    """model(h_a, h_t, mask_a, mask_t) for HOST tensors with copy/compute overlap.
    Returns (logits, beta, z) on `out_device`."""
    dev = torch.device(device)
    B = h_a.shape[0]
    main = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(dev)
    outs = []

    def stage(s):
        e = min(B, s + slab)
        with torch.cuda.stream(copy):
            items = [h_a[s:e].to(dev, non_blocking=True), h_t[s:e].to(dev, non_blocking=True),
                     None if mask_a is None else mask_a[s:e].to(dev, non_blocking=True),
                     None if mask_t is None else mask_t[s:e].to(dev, non_blocking=True)]
            ev = torch.cuda.Event()
            ev.record(copy)
        return items, ev

    nxt = stage(0)
    for s in range(0, B, slab):
        items, ev = nxt
        if s + slab < B:
            nxt = stage(s + slab)
        main.wait_event(ev)
        for x in items:
            if x is not None:
                x.record_stream(main)
        outs.append(model(*items)[:3])
    logits = torch.cat([o[0] for o in outs]).to(out_device, non_blocking=False)
    beta = torch.cat([o[1] for o in outs]).to(out_device)
    z = torch.cat([o[2] for o in outs]).to(out_device)
    return logits, beta, z
