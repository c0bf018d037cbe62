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


_STAGING = {}   # (device, slab, shapes) -> two sets of device staging buffers + events, reused across calls


def _staging(dev, slab, h_a, h_t, mask_a, mask_t):
    key = (str(dev), slab, tuple(h_a.shape[1:]), tuple(h_t.shape[1:]), h_a.dtype, h_t.dtype,
           mask_a is not None, mask_t is not None)
    st = _STAGING.get(key)
    if st is None:
        def mk():
            return dict(a=torch.empty((slab,) + tuple(h_a.shape[1:]), dtype=h_a.dtype, device=dev),
                        t=torch.empty((slab,) + tuple(h_t.shape[1:]), dtype=h_t.dtype, device=dev),
                        ma=None if mask_a is None else torch.empty((slab, mask_a.shape[1]), dtype=torch.bool, device=dev),
                        mt=None if mask_t is None else torch.empty((slab, mask_t.shape[1]), dtype=torch.bool, device=dev),
                        copied=torch.cuda.Event(), consumed=torch.cuda.Event())
        st = dict(bufs=[mk(), mk()], copy=torch.cuda.Stream(dev))
        _STAGING.clear()      # keep one configuration resident
        _STAGING[key] = st
    return st


@torch.no_grad()
def forward_from_host(model, h_a: torch.Tensor, h_t: torch.Tensor, mask_a: Optional[torch.Tensor] = None,
                      mask_t: Optional[torch.Tensor] = None, device="cuda", slab: int = 256,
                      out_device="cpu"):
    """model(h_a, h_t, mask_a, mask_t) for HOST tensors with copy/compute overlap.
    Two fixed device staging sets are cycled: the side stream copies slab i+1 (and, once the
    fp32->bf16 cast of slab i has consumed its staging set, slab i+2) while slab i computes; no
    device memory is allocated or freed per slab for the inputs.  Returns (logits, beta, z) on
    `out_device`."""
    dev = torch.device(device)
    B = h_a.shape[0]
    if h_a.dim() != 3 or h_t.dim() != 3:
        # utterance-level [B, d] inputs: nothing to pipeline
        out = model(h_a.to(dev), h_t.to(dev), None if mask_a is None else mask_a.to(dev),
                    None if mask_t is None else mask_t.to(dev))[:3]
        return tuple(o.to(out_device) for o in out)
    slab = max(1, min(slab, B))
    st = _staging(dev, slab, h_a, h_t, mask_a, mask_t)
    main = torch.cuda.current_stream(dev)
    copy = st["copy"]
    outs = []
    starts = list(range(0, B, slab))

    def stage(i):
        s = starts[i]
        n = min(B, s + slab) - s
        buf = st["bufs"][i % 2]
        with torch.cuda.stream(copy):
            copy.wait_event(buf["consumed"])     # the slab that used this set two steps ago has been cast
            buf["a"][:n].copy_(h_a[s:s + n], non_blocking=True)
            buf["t"][:n].copy_(h_t[s:s + n], non_blocking=True)
            if mask_a is not None:
                buf["ma"][:n].copy_(mask_a[s:s + n], non_blocking=True)
            if mask_t is not None:
                buf["mt"][:n].copy_(mask_t[s:s + n], non_blocking=True)
            buf["copied"].record(copy)
        return n

    ns = {0: stage(0)}
    if len(starts) > 1:
        ns[1] = stage(1)
    for i in range(len(starts)):
        buf, n = st["bufs"][i % 2], ns[i]
        main.wait_event(buf["copied"])
        ma = None if mask_a is None else buf["ma"][:n].clone()
        mt = None if mask_t is None else buf["mt"][:n].clone()
        early = h_a.dtype == torch.float32 and h_t.dtype == torch.float32 and h_a.shape[2] % 8 == 0 and h_t.shape[2] % 8 == 0
        if early:
            # the fp32 features are only read by the bf16 cast: after it the staging set is free again
            xa = E.to_seq(buf["a"][:n], "h_a").x.view(n, h_a.shape[1], -1)
            xt = E.to_seq(buf["t"][:n], "h_t").x.view(n, h_t.shape[1], -1)
            buf["consumed"].record(main)
            if i + 2 < len(starts):
                ns[i + 2] = stage(i + 2)
            outs.append(model(xa, xt, ma, mt)[:3])
        else:
            outs.append(model(buf["a"][:n], buf["t"][:n], ma, mt)[:3])
            buf["consumed"].record(main)
            if i + 2 < len(starts):
                ns[i + 2] = stage(i + 2)
    logits = torch.cat([o[0] for o in outs]).to(out_device)
    beta = torch.cat([o[1] for o in outs]).to(out_device)
    z = torch.cat([o[2] for o in outs]).to(out_device)
    return logits, beta, z
