"""Host staging, length bucketing and batch sharding around the forward path.

* ``forward_from_host``: the end-to-end call for features that live in (pinned) host
  memory, as the reference's DataLoader/collate delivers them
  (scripts/fusion/train_fusion_seq_level_decoder.py:191-232, :306-308).  The batch is cut
  into slabs of utterances; slab i+1 is copied host->device on a side stream while slab i
  computes, and only logits / beta / z come back.
* ``bucket=True`` / ``forward_bucketed``: the collate zero-pads every utterance to the batch
  maximum (True = PAD mask).  Padded rows never reach logits / beta / z -- keys are masked,
  pooled means are masked, the decoder reads the fused mask -- so utterances are sorted by valid
  length and each slab is trimmed to its own maximum (``bucket_plan``): the GEMMs and the attention
  no longer pay for the padding, and the padding never crosses PCIe.
* ``shard_bounds`` / ``shard_by_length`` / ``gather_outputs``: batch sharding over ranks
  (utterances are independent); the only collective is one all_gather of the
  small outputs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import engine as E
from . import lib as L
from . import ops


# --------------------------------------------------------------------------- #
# sharding
# --------------------------------------------------------------------------- #
def shard_bounds(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of the batch: rank r gets [lo, hi); the first B % world ranks get one extra."""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_length(len_a: torch.Tensor, len_t: Optional[torch.Tensor], world: int) -> List[torch.Tensor]:
    """Length-balanced split of a ragged batch ("sort/bucket by length before sharding"):
    utterances sorted by (len_a, len_t) are dealt in a serpentine (0..w-1, w-1..0, ...), so every rank
    gets the same number of utterances (+-1) and the same length distribution -- the per-rank work,
    which grows with (T_a + T_t)^2 in the attention and linearly elsewhere, is balanced to a fraction of
    a percent.  Returns one int64 index vector per rank (ascending length within a rank)."""
    key = len_a.to("cpu", torch.int64)
    if len_t is not None and len_t.numel():
        lt = len_t.to("cpu", torch.int64)
        key = key * (int(lt.max().item()) + 1) + lt
    order = torch.argsort(key, stable=True)
    pos = torch.arange(order.shape[0])
    lane = pos % world
    rank_of = torch.where((pos // world) % 2 == 0, lane, world - 1 - lane)
    return [order[rank_of == r].contiguous() for r in range(world)]


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


# --------------------------------------------------------------------------- #
# slab plans
# --------------------------------------------------------------------------- #
def slab_schedule(B: int, slab: int = 512, host_cast_every: int = 2, ramp: bool = False):
    """The slab plan forward_from_host follows for a dense batch: [(start, end, host_cast)], in order.

    Full slabs, every `host_cast_every`-th of which (0 = never) is converted to bf16 by the host cores
    and travels at half the bytes.  ramp=True opens the plan with small fp32 slabs (slab/8, slab/4,
    slab/2 utterances; only when at least four full slabs follow) so that the GPU starts after 2 ms of
    copying instead of 16 -- measured on B200 it does not pay: a 64-utterance slab computes at half the
    rate of a 512-utterance one (launch-bound decoder tail), which costs what the earlier start saves
    (tools/e2e_schedule_sim.py, profiles/r01_e2e_timeline_v11.txt), so it is off by default."""
    slab = max(1, min(slab, B))
    sizes = []
    left = B
    if ramp and slab >= 64 and B >= 5 * slab:
        for n in (slab // 8, slab // 4, slab // 2):
            sizes.append(n)
            left -= n
    n_ramp = len(sizes)
    while left > 0:
        n = min(slab, left)
        sizes.append(n)
        left -= n
    n_full = len(sizes) - n_ramp
    host_cast = host_cast_every > 0 and n_full > 2
    plan, s = [], 0
    for i, n in enumerate(sizes):
        j = i - n_ramp
        plan.append((s, s + n, bool(host_cast and j >= 0 and j % host_cast_every == host_cast_every - 1)))
        s += n
    return plan


def default_host_cast_every(threads: Optional[int] = None) -> int:
    """Every second slab is pre-cast by the host only when this process has the cores for it: 16 threads convert a
    512-utterance north-star slab in 11-15 ms (about the GPU's 14 ms per slab), 2-4 threads (8 or 4 ranks sharing a
    16-core host) in 45-130 ms -- slower than sending the slab as fp32."""
    threads = max(1, torch.get_num_threads()) if threads is None else threads
    return 2 if threads >= 8 else 0


def h2d_bytes(B: int, fp32_bytes_per_utt: int, slab: int = 512, host_cast_every: int = 2, ramp: bool = False) -> int:
    """Bytes forward_from_host copies host->device for B dense utterances of fp32 features whose feature
    dims are multiples of 8: host-pre-cast slabs travel as bf16 (half the bytes)."""
    return sum((e - s) * fp32_bytes_per_utt // (2 if half else 1)
               for s, e, half in slab_schedule(B, slab, host_cast_every, ramp))


def valid_lengths(mask: Optional[torch.Tensor], B: int, T: int) -> torch.Tensor:
    """int32 [B] (on the mask's device, CPU when mask is None): index of the last valid position + 1;
    T without a mask, 0 when every position is PAD.  Host masks are handled with torch index
    bookkeeping, device masks by the hriemo_mask_lengths kernel."""
    if mask is None:
        return torch.full((B,), T, dtype=torch.int32)
    if mask.is_cuda:
        return ops.mask_lengths(mask)
    valid = ~mask.to(torch.bool)
    pos = torch.arange(1, T + 1, dtype=torch.int32)
    return (valid.to(torch.int32) * pos).amax(dim=1).to(torch.int32)


@dataclass
class Bucket:
    start: int   # [start, end) in the sorted order
    end: int
    T_a: int     # time extents of this slab after trimming
    T_t: int


def bucket_plan(len_a: torch.Tensor, len_t: torch.Tensor, T_a: int, T_t: int, rows_per_slab: int = 512 * 500,
                max_utts: int = 2048, ramp: bool = False) -> Tuple[torch.Tensor, List[Bucket]]:
    """Sort utterances by valid length (audio first, then text; ascending, so the plan opens with the
    cheap slabs and the pipeline fills quickly) and cut the order into slabs of at most `rows_per_slab`
    audio rows / `max_utts` utterances, each trimmed to its own maxima.  ramp: the first two slabs get a
    quarter and a half of the row budget (host staging: nothing computes until the first slab has landed).

    The trimmed extents keep what the reference's forward needs: at least one row per stream (a fully
    padded utterance still runs, all-masked, and comes out NaN like the reference), and T_a >= T_t
    unless T_a == 1 (models/beta_gate_tacfn.py:116 slices h_a[:, :T_t]).  Returns (order, buckets);
    order is an int32 host vector, order[k] = original index of the k-th utterance in sorted order."""
    la = len_a.to("cpu", torch.int64).clamp(1, T_a)
    lt = len_t.to("cpu", torch.int64).clamp(1, T_t)
    B = la.shape[0]
    order = torch.argsort(la * (T_t + 1) + lt, stable=True)
    la_s, lt_s = la[order].tolist(), lt[order].tolist()
    buckets: List[Bucket] = []
    total_rows = sum(la_s)
    s = 0
    while s < B:
        e = s
        ta = tt = 1
        budget = rows_per_slab
        if ramp and len(buckets) < 2 and total_rows > 4 * rows_per_slab:
            budget = rows_per_slab // (4 >> len(buckets))
        while e < B and e - s < max_utts:
            ta2, tt2 = max(ta, la_s[e]), max(tt, lt_s[e])
            if e > s and (e - s + 1) * max(ta2, tt2) > budget:
                break
            ta, tt = ta2, tt2
            e += 1
        if T_a > 1:
            ta = min(T_a, max(ta, tt))   # the gate blends h_a[:, :T_t] with h_t
        buckets.append(Bucket(s, e, ta, tt))
        s = e
    return order.to(torch.int32), buckets


def bucket_stats(len_a: torch.Tensor, len_t: torch.Tensor, T_a: int, T_t: int, buckets: Sequence[Bucket]) -> dict:
    """Rows the padded batch holds, rows the bucketed plan processes, and rows that are actually valid."""
    B = int(len_a.shape[0])
    return {"padded_rows": B * (T_a + T_t),
            "bucketed_rows": sum((b.end - b.start) * (b.T_a + b.T_t) for b in buckets),
            "valid_rows": int(len_a.clamp(1, T_a).sum() + len_t.clamp(1, T_t).sum())}


# --------------------------------------------------------------------------- #
# device-resident bucketed forward
# --------------------------------------------------------------------------- #
@torch.no_grad()
def forward_bucketed(model, h_a: torch.Tensor, h_t: torch.Tensor, mask_a: Optional[torch.Tensor] = None,
                     mask_t: Optional[torch.Tensor] = None, rows_per_slab: int = 512 * 500, max_utts: int = 2048):
    """model(h_a, h_t, mask_a, mask_t)[:3] for DEVICE tensors, without the padding: utterances are sorted
    by valid length (one small D2H of the lengths -- this call synchronises the host once), gathered into
    slabs trimmed to their own maxima (hriemo_gather_utterances_bf16 / _masks), run through the model,
    and the results scattered back to the original order.  Same outputs as the padded forward up to the
    bf16 path's own rounding (a masked key contributes exactly 0; padded query rows reach no output)."""
    if h_a.dim() != 3 or h_t.dim() != 3 or (mask_a is None and mask_t is None) or h_a.shape[1] == 1:
        return tuple(model(h_a, h_t, mask_a, mask_t)[:3])
    E.require_cuda(h_a, "h_a")
    E.require_cuda(h_t, "h_t")
    B, T_a, T_t = h_a.shape[0], h_a.shape[1], h_t.shape[1]
    mask_a = E.check_mask(mask_a, B, T_a, "mask_a")
    mask_t = E.check_mask(mask_t, B, T_t, "mask_t")
    la = valid_lengths(mask_a, B, T_a)
    lt = valid_lengths(mask_t, B, T_t)
    order, buckets = bucket_plan(la, lt, T_a, T_t, rows_per_slab, max_utts)
    order_dev = order.to(h_a.device, non_blocking=True)
    d_a, d_t = h_a.shape[2], h_t.shape[2]
    outs = None
    for bk in buckets:
        utt = order_dev[bk.start:bk.end]
        a = ops.gather_utterances(h_a, utt, bk.T_a)
        t = ops.gather_utterances(h_t, utt, bk.T_t)
        ma = None if mask_a is None else ops.gather_masks(mask_a, utt, bk.T_a)
        mt = None if mask_t is None else ops.gather_masks(mask_t, utt, bk.T_t)
        # feature dims that are not multiples of 8 (MOSEI 74 / 300) come back zero-padded: hand the model
        # the logical columns; its own input cast pads them again
        res = model(a if a.shape[2] == d_a else a[:, :, :d_a].float(), t if t.shape[2] == d_t else t[:, :, :d_t].float(),
                    ma, mt)[:3]
        if outs is None:
            outs = [torch.empty((B,) + tuple(r.shape[1:]), dtype=torch.float32, device=h_a.device) for r in res]
        for dst, r in zip(outs, res):
            ops.scatter_rows(r.float() if r.dtype != torch.float32 else r, utt, dst)
    return tuple(outs)


# --------------------------------------------------------------------------- #
# host staging
# --------------------------------------------------------------------------- #
_STAGING = {}   # configuration -> staging buffers + events, reused across calls
# Pinned bf16 staging sets the converting worker rotates through.  With two, the third conversion of a call waited
# ~20 ms for the copy of the first to be sent (the copy engine was busy with fp32 slabs in between); three keep the
# host cores converting back to back (tools/e2e_pack_probe.py).
HOST_SETS = 3
# TwoEndedPlan.claim_back: the host side takes another slab while at least PLAN_MIN_LEFT are unclaimed and the copy side
# has (unclaimed - 1 + PLAN_RESERVE) decisions' worth of other work for the time of one conversion (PLAN_RESERVE: copies
# already queued in the engine when the decision is made).
PLAN_MIN_LEFT = 2
PLAN_RESERVE = 0.0
HOST_RESERVE = 2
# what the host-staged entry points moved, accumulated over calls (bench.py reads and resets it): bytes copied
# host -> device (counted from the tensors copied), slabs, and how many of them the host cores pre-cast
STATS = dict(h2d_bytes=0, slabs=0, host_cast_slabs=0, calls=0)


def reset_stats() -> None:
    for k in STATS:
        STATS[k] = 0


def _staging(dev, dtype_a, dtype_t, host_dtype, elems_a: int, elems_t: int, rows_a: int, rows_t: int, direct_sets: bool,
             host_sets: bool, with_ma: bool, with_mt: bool):
    """Staging sets for one configuration (device, dtypes).  Capacities only GROW: a stream of batches whose bucket
    plans (or last partial batch) need slightly different sizes reuses the same pinned / device buffers through
    slices instead of re-pinning hundreds of MB per call; a set is reallocated only when a call needs more room,
    or a kind of buffer (masks, host sets, direct sets) the resident one lacks."""
    key = (str(dev), dtype_a, dtype_t, host_dtype)
    need = dict(elems_a=elems_a, elems_t=elems_t, rows_a=rows_a, rows_t=rows_t)
    st = _STAGING.get(key)
    if st is not None:
        cap, has = st["cap"], st["has"]
        if (all(cap[k] >= v for k, v in need.items()) and (has["direct"] or not direct_sets) and (has["host"] or not host_sets)
                and (has["ma"] or not with_ma) and (has["mt"] or not with_mt)):
            return st
        # grow: keep every capacity and every kind the resident configuration already had
        need = {k: max(v, cap[k]) for k, v in need.items()}
        direct_sets, host_sets = direct_sets or has["direct"], host_sets or has["host"]
        with_ma, with_mt = with_ma or has["ma"], with_mt or has["mt"]
    if _STAGING:
        torch.cuda.synchronize(dev)       # copies in flight still read / write the buffers about to be dropped
    ea, et, ra, rt = need["elems_a"], need["elems_t"], need["rows_a"], need["rows_t"]

    def mk(da, dt):
        return dict(a=torch.empty((ea,), dtype=da, device=dev),
                    t=torch.empty((et,), dtype=dt, device=dev),
                    ma=torch.empty((ra,), dtype=torch.bool, device=dev) if with_ma else None,
                    mt=torch.empty((rt,), dtype=torch.bool, device=dev) if with_mt else None,
                    copied=torch.cuda.Event(), consumed=torch.cuda.Event())

    _STAGING.clear()      # keep one configuration resident
    st = dict(copy=torch.cuda.Stream(dev), cap=need, has=dict(direct=direct_sets, host=host_sets, ma=with_ma, mt=with_mt))
    if direct_sets:
        # landing sets of the source dtype: free again as soon as the GPU bf16 cast has read them
        st["direct"] = [mk(dtype_a, dtype_t), mk(dtype_a, dtype_t)]
    if host_sets:
        # slabs prepared by the host cores (bf16 pack of a padded batch, or rows of a shard): pinned staging
        # on the host (HOST_SETS of them in rotation) and landing buffers on the device (ping-pong, released
        # when the slab's forward is done)
        def mk_host():
            return dict(a=torch.empty((ea,), dtype=host_dtype).pin_memory(),
                        t=torch.empty((et,), dtype=host_dtype).pin_memory(),
                        sent=torch.cuda.Event())
        st["host16"] = [mk_host() for _ in range(HOST_SETS)]
        st["dev16"] = [mk(host_dtype, host_dtype), mk(host_dtype, host_dtype)]
    _STAGING[key] = st
    return st


@dataclass
class _Slab:
    lo: int                         # dense plan: utterances [lo, hi) of the batch
    hi: int                         # bucketed plan: positions [lo, hi) of the sorted order
    T_a: int
    T_t: int
    host_cast: bool                 # prepared by the host worker (pack / shard read) instead of copied as it is
    utt: Optional[torch.Tensor] = None      # bucketed: host indices of the slab's utterances
    utt_dev: Optional[torch.Tensor] = None  # the same as int32 on the device (mask gather, result scatter)

    @property
    def n(self) -> int:
        return self.hi - self.lo


class PendingResult:
    """Results of a forward_from_host / forward_from_shard call made with wait=False: the pinned host tensors
    are being filled by asynchronous D2H copies; wait() blocks until they are complete and returns them.
    The next call may be issued before wait(): its first copies then overlap this call's last slabs."""

    def __init__(self, tensors, event):
        self._tensors, self._event = tuple(tensors), event

    def wait(self):
        self._event.synchronize()
        return self._tensors


def _run_pipeline(model, dev, B: int, slabs: List[_Slab], d_a: int, d_t: int, dtype_a, dtype_t, host_dtype,
                  pack, direct_src, mask_a, mask_t, mask_a_dev, mask_t_dev, early: bool, out_device, trace,
                  wait: bool = True):
    """The staging engine behind forward_from_host / forward_from_shard: slab i+1 (and i+2) is prepared
    and copied on a side stream while slab i computes.
      pack(s, hb)      fills the pinned host set hb["a"], hb["t"] for a host-prepared slab (worker thread);
      direct_src(s)    -> flat host views (a, t) of a slab that is copied as it is;
      mask_a / mask_t  host masks of dense slabs (copied per slab); mask_*_dev: whole masks already on the
                       device for bucketed slabs (each slab gathers and trims its own)."""
    import threading

    rows_a = max(s.n * s.T_a for s in slabs)
    rows_t = max(s.n * s.T_t for s in slabs)
    any_host = any(s.host_cast for s in slabs)
    dense_masks = any(s.utt is None for s in slabs)
    st = _staging(dev, dtype_a, dtype_t, host_dtype, rows_a * d_a, rows_t * d_t, rows_a, rows_t,
                  any(not s.host_cast for s in slabs), any_host, mask_a is not None and dense_masks,
                  mask_t is not None and dense_masks)
    main = torch.cuda.current_stream(dev)
    copy = st["copy"]

    # ---- worker: host-side preparation of the designated slabs, in order
    ready = [threading.Event() for _ in slabs]
    enqueued = [threading.Event() for _ in slabs]   # the main thread has enqueued slab i's copies (and recorded `sent`)
    host_ids = [i for i, s in enumerate(slabs) if s.host_cast]
    failure = []

    def prepare():
        try:
            for k, i in enumerate(host_ids):
                hb = st["host16"][k % HOST_SETS]
                if k >= HOST_SETS:
                    # this host buffer last carried host slab k-HOST_SETS: its copy must have been ENQUEUED (the
                    # event below is re-recorded per use) before waiting for it to have been READ
                    enqueued[host_ids[k - HOST_SETS]].wait()
                    if failure:
                        return
                hb["sent"].synchronize()          # the copy that last read this host buffer is done
                pack(slabs[i], hb)
                ready[i].set()
        except Exception as e:   # surfaced on the main thread
            failure.append(e)
            for ev in ready:
                ev.set()

    worker = None
    if any_host:
        worker = threading.Thread(target=prepare, daemon=True)
        worker.start()

    n_host = [0]
    n_direct = [0]

    def stage(i):
        s = slabs[i]
        na, nt = s.n * s.T_a, s.n * s.T_t
        if s.host_cast:
            k = n_host[0]
            n_host[0] += 1
            hb, buf = st["host16"][k % HOST_SETS], st["dev16"][k % 2]
            ready[i].wait()
            if failure:
                raise failure[0]
            src_a, src_t = hb["a"][: na * d_a], hb["t"][: nt * d_t]
        else:
            hb, buf = None, st["direct"][n_direct[0] % 2]
            n_direct[0] += 1
            src_a, src_t = direct_src(s)
        with torch.cuda.stream(copy):
            copy.wait_event(buf["consumed"])     # the slab that last used this device set is done with it
            if trace is not None:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record(copy)
            buf["a"][: na * d_a].copy_(src_a, non_blocking=True)
            buf["t"][: nt * d_t].copy_(src_t, non_blocking=True)
            STATS["h2d_bytes"] += src_a.numel() * src_a.element_size() + src_t.numel() * src_t.element_size()
            STATS["slabs"] += 1
            STATS["host_cast_slabs"] += 1 if s.host_cast else 0
            if s.utt is None:
                if mask_a is not None:
                    buf["ma"][:na].copy_(mask_a[s.lo:s.hi].view(-1), non_blocking=True)
                if mask_t is not None:
                    buf["mt"][:nt].copy_(mask_t[s.lo:s.hi].view(-1), non_blocking=True)
            if hb is not None:
                hb["sent"].record(copy)
            if trace is not None:
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record(copy)
                trace.append(dict(slab=i, n=s.n, T_a=s.T_a, T_t=s.T_t, host_cast=s.host_cast, copy0=ev0, copy1=ev1))
            buf["copied"].record(copy)
        enqueued[i].set()
        return buf

    to_host = torch.device(out_device).type == "cpu"
    outs = []
    host_out = None   # pinned result tensors (from torch's caching host allocator)
    dev_out = None    # bucketed plan: full-size device results in the ORIGINAL order

    def emit(i, res):
        nonlocal host_out, dev_out
        s = slabs[i]
        if s.utt is not None:
            if dev_out is None:
                dev_out = [torch.empty((B,) + tuple(r.shape[1:]), dtype=torch.float32, device=dev) for r in res]
            for dst, r in zip(dev_out, res):
                ops.scatter_rows(r if r.dtype == torch.float32 else r.float(), s.utt_dev, dst)
            return
        if not to_host:
            outs.append(res)
            return
        if host_out is None:
            host_out = [torch.empty((B,) + tuple(r.shape[1:]), dtype=r.dtype, pin_memory=True) for r in res]
        for dst, r in zip(host_out, res):
            dst[s.lo:s.lo + r.shape[0]].copy_(r, non_blocking=True)   # D2H on the compute stream, behind this slab

    def run_slabs():
        nonlocal host_out
        staged = {0: stage(0)}
        if len(slabs) > 1:
            staged[1] = stage(1)
        for i, s in enumerate(slabs):
            n = s.n
            buf = staged.pop(i)
            main.wait_event(buf["copied"])
            if trace is not None:
                k0 = torch.cuda.Event(enable_timing=True)
                k0.record(main)
            if s.utt is not None:
                # the whole (small) masks went to the device once; each slab gathers and trims its own
                ma = None if mask_a_dev is None else ops.gather_masks(mask_a_dev, s.utt_dev, s.T_a)
                mt = None if mask_t_dev is None else ops.gather_masks(mask_t_dev, s.utt_dev, s.T_t)
            else:
                ma = None if mask_a is None else buf["ma"][: n * s.T_a].view(n, s.T_a).clone()
                mt = None if mask_t is None else buf["mt"][: n * s.T_t].view(n, s.T_t).clone()
            va = buf["a"][: n * s.T_a * d_a].view(n, s.T_a, d_a)
            vt = buf["t"][: n * s.T_t * d_t].view(n, s.T_t, d_t)
            if va.dtype == torch.float32 and early:
                # fp32 features are only read by the bf16 cast: after it the staging set is free again
                xa = E.to_seq(va, "h_a").x.view(n, s.T_a, -1)
                xt = E.to_seq(vt, "h_t").x.view(n, s.T_t, -1)
                buf["consumed"].record(main)
                emit(i, model(xa, xt, ma, mt)[:3])
            else:
                # bf16 landed on the device: the model reads it in place; the set is free after the forward
                emit(i, model(va, vt, ma, mt)[:3])
                buf["consumed"].record(main)
            if trace is not None:
                k1 = torch.cuda.Event(enable_timing=True)
                k1.record(main)
                trace[i]["comp0"], trace[i]["comp1"] = k0, k1
            if i + 2 < len(slabs):
                staged[i + 2] = stage(i + 2)
        if worker is not None:
            worker.join()
        if dev_out is not None:
            if not to_host:
                return tuple(o.to(out_device) for o in dev_out)
            host_out = [torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in dev_out]
            for dst, o in zip(host_out, dev_out):
                dst.copy_(o, non_blocking=True)
        if to_host:
            if not wait:
                done = torch.cuda.Event()
                done.record(main)
                return PendingResult(host_out, done)
            main.synchronize()
            return tuple(host_out)
        logits = torch.cat([o[0] for o in outs]).to(out_device)
        beta = torch.cat([o[1] for o in outs]).to(out_device)
        z = torch.cat([o[2] for o in outs]).to(out_device)
        return logits, beta, z

    try:
        STATS["calls"] += 1
        return run_slabs()
    except BaseException as e:
        failure.append(e)
        for ev in enqueued:
            ev.set()          # never leave the worker waiting on the main thread
        raise


class TwoEndedPlan:
    """The slabs of a dense batch, claimed from both ends: `next` (the copy side: fp32 slabs in order, or a slab
    the host side has finished) and `claim_back` (the host side: slabs to pre-cast, last first).  Every slab is
    handed out exactly once; the host side stops claiming when the slabs left are fewer than the copy side clears
    during one conversion (both rates measured as the plan runs), so the batch does not end waiting for the host.
    Thread-safe; no CUDA in here (tests/test_sharding_cpu.py drives it with plain threads)."""

    def __init__(self, n_slabs: int, t_pack: Optional[float] = None, t_step: Optional[float] = None):
        import threading
        from collections import deque
        self.cv = threading.Condition()
        self.front, self.back = 0, n_slabs
        self.ready = deque()         # (slab, k) converted by the host side, not yet sent
        self.busy = False            # the host side holds a claimed slab it has not published yet
        self.failure = None
        self.t_pack = t_pack         # seconds per conversion (last; a stream of calls hands its rates on)
        self.t_step = t_step         # seconds between PACED decisions of the copy side (running mean)
        self._last = None

    def claim_back(self):
        """-> slab index for the host side, or None when it should stop."""
        with self.cv:
            left = self.back - self.front
            if self.t_pack is None or self.t_step is None:
                ok = left > 2
            else:
                # worth it while the copy side has enough OTHER slabs to send for (most of) the time one conversion
                # takes: the slab then crosses PCIe at half the bytes and the engine never waited for it.  (The
                # earlier rule, left > ceil(t_pack / t_step), stopped one slab early when the two rates are close --
                # 16 host threads convert a slab in 17 ms while an fp32 slab takes the engine 16 ms -- and left the
                # copy engine the bound: 1.8 of 8 slabs pre-cast, 6.3 GB per step.)
                ok = left >= PLAN_MIN_LEFT and (left - 1 + PLAN_RESERVE) * self.t_step >= 0.75 * self.t_pack
            if not ok or self.failure is not None:
                return None
            self.back -= 1
            self.busy = True
            return self.back

    def publish(self, slab: int, k: int, seconds: float) -> None:
        with self.cv:
            self.t_pack = seconds
            self.ready.append((slab, k))
            self.busy = False
            self.cv.notify_all()

    def host_done(self, failure=None) -> None:
        with self.cv:
            if failure is not None and self.failure is None:
                self.failure = failure
            self.busy = False
            self.cv.notify_all()

    def next(self, paced: bool = True):
        """-> (slab, k) for the copy side: k >= 0 = the k-th slab the host side converted, k = -1 = send it as it
        is; None when every slab has been handed out.  Blocks while the host side still holds a slab.
        paced=False: the caller did not wait for the copy engine before this decision (the first two of a call, which
        only fill its queue) -- the interval in front of it says nothing about the engine's rate and is not measured.
        (Measured, it was: 0.2 ms between the first two decisions made every later call of a stream stop converting
        after one slab.)"""
        import time
        with self.cv:
            while True:
                if self.failure is not None:
                    raise self.failure
                if self.ready:
                    item = self.ready.popleft()
                    break
                if self.front < self.back:
                    item = (self.front, -1)
                    self.front += 1
                    break
                if not self.busy:
                    return None
                self.cv.wait(0.05)
            now = time.perf_counter()
            if paced and self._last is not None:
                dt = now - self._last
                self.t_step = dt if self.t_step is None else 0.5 * (self.t_step + dt)
            self._last = now if paced else None
            return item


def _run_dense_dynamic(model, dev, h_a, h_t, mask_a, mask_t, slab: int, out_device, threads: int, wait: bool,
                       trace: Optional[list] = None):
    """forward_from_host's default engine for a dense fp32 batch (host_cast_every="auto"): the copy engine and the
    host cores work through the slabs FROM BOTH ENDS and meet wherever their speeds put them.

      * The main thread sends fp32 slabs from the front of the batch, as they are, paced by the copy engine (it keeps
        two copies queued and blocks on the older one before deciding what to send next).
      * A worker thread converts slabs to bf16 from the BACK of the batch into pinned staging
        (hriemo_host_pack_bf16: half the bytes cross PCIe); a slab it has finished is sent next, ahead of the
        next fp32 one (`TwoEndedPlan`).
    The share of pre-cast slabs thereby follows what the process actually has: host threads (16 threads convert a
    512-utterance slab in ~8 ms, 2 threads in ~60 ms) and PCIe bandwidth (55 GB/s alone, ~23 GB/s per GPU when the
    ranks of one box share switch uplinks: DESIGN sec. 7) -- the fixed "every 2nd slab, only with >= 8 threads"
    plan switched the pre-cast off exactly where H2D was the bound.  Slabs are independent and both casts round to
    nearest even, so the results do not depend on which slabs the host converted (bit-identical, tested)."""
    import threading
    import time
    from collections import deque

    B, T_a, d_a = h_a.shape
    T_t, d_t = h_t.shape[1], h_t.shape[2]
    bounds = [(s, min(B, s + slab)) for s in range(0, B, slab)]
    n_sl = len(bounds)
    st = _staging(dev, h_a.dtype, h_t.dtype, torch.bfloat16, slab * T_a * d_a, slab * T_t * d_t, slab * T_a, slab * T_t,
                  True, True, mask_a is not None, mask_t is not None)
    main = torch.cuda.current_stream(dev)
    copy = st["copy"]
    rates = st.setdefault("rates", {})     # conversion / copy-side rates of the previous call on these staging sets
    plan = TwoEndedPlan(n_sl, rates.get((n_sl, "pack")), rates.get((n_sl, "step")))
    host_enq = [threading.Event() for _ in range(n_sl)]   # k-th host-prepared slab: its copy has been enqueued

    def prepare():
        k = 0
        try:
            while True:
                i = plan.claim_back()
                if i is None:
                    break
                hb = st["host16"][k % HOST_SETS]
                if k >= HOST_SETS:
                    host_enq[k - HOST_SETS].wait()   # the copy that last read this host set has been ENQUEUED ...
                    if plan.failure is not None:
                        break
                hb["sent"].synchronize()          # ... and is done
                t0 = time.perf_counter()
                lo, hi = bounds[i]
                ops.host_pack_bf16(h_a[lo:hi], hb["a"], T_a, threads=threads)
                ops.host_pack_bf16(h_t[lo:hi], hb["t"], T_t, threads=threads)
                plan.publish(i, k, time.perf_counter() - t0)
                k += 1
            plan.host_done()
        except BaseException as e:   # surfaced on the main thread
            plan.host_done(e)

    worker = threading.Thread(target=prepare, daemon=True)
    n_direct = [0]
    n_host = [0]
    copied_events = []

    def next_stage():
        """-> (slab index, device set, host-converted?) of the next slab to compute, its copies enqueued; None when
        every slab is out."""
        paced = len(copied_events) >= 2
        if paced:
            copied_events[-2].synchronize()       # pace the decisions by the copy engine: two copies queued at most
        item = plan.next(paced)
        if item is None:
            return None
        i, k = item
        lo, hi = bounds[i]
        n = hi - lo
        na, nt = n * T_a, n * T_t
        if k >= 0:
            hb, buf = st["host16"][k % HOST_SETS], st["dev16"][n_host[0] % 2]
            n_host[0] += 1
            src_a, src_t = hb["a"][: na * d_a], hb["t"][: nt * d_t]
        else:
            hb, buf = None, st["direct"][n_direct[0] % 2]
            n_direct[0] += 1
            src_a, src_t = h_a[lo:hi].view(-1), h_t[lo:hi].view(-1)
        with torch.cuda.stream(copy):
            copy.wait_event(buf["consumed"])      # the slab that last used this device set is done with it
            if trace is not None:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record(copy)
            buf["a"][: na * d_a].copy_(src_a, non_blocking=True)
            buf["t"][: nt * d_t].copy_(src_t, non_blocking=True)
            STATS["h2d_bytes"] += src_a.numel() * src_a.element_size() + src_t.numel() * src_t.element_size()
            STATS["slabs"] += 1
            STATS["host_cast_slabs"] += 1 if k >= 0 else 0
            if mask_a is not None:
                buf["ma"][:na].copy_(mask_a[lo:hi].view(-1), non_blocking=True)
            if mask_t is not None:
                buf["mt"][:nt].copy_(mask_t[lo:hi].view(-1), non_blocking=True)
            if hb is not None:
                hb["sent"].record(copy)
            if trace is not None:
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record(copy)
                trace.append(dict(slab=i, n=n, T_a=T_a, T_t=T_t, host_cast=k >= 0, copy0=ev0, copy1=ev1))
            ev = torch.cuda.Event()
            ev.record(copy)
            buf["copied"].record(copy)
        copied_events.append(ev)
        if k >= 0:
            host_enq[k].set()
        return i, buf, k >= 0

    to_host = torch.device(out_device).type == "cpu"
    outs = {}
    host_out = None

    def emit(i, res):
        nonlocal host_out
        lo = bounds[i][0]
        if not to_host:
            outs[i] = res
            return
        if host_out is None:
            host_out = [torch.empty((B,) + tuple(r.shape[1:]), dtype=r.dtype, pin_memory=True) for r in res]
        for dst, r in zip(host_out, res):
            dst[lo:lo + r.shape[0]].copy_(r, non_blocking=True)   # D2H on the compute stream, behind this slab

    try:
        worker.start()
        pending = deque()
        for _ in range(2):
            nxt = next_stage()
            if nxt is not None:
                pending.append(nxt)
        while pending:
            i, buf, from_host = pending.popleft()
            lo, hi = bounds[i]
            n = hi - lo
            main.wait_event(buf["copied"])
            if trace is not None:
                k0 = torch.cuda.Event(enable_timing=True)
                k0.record(main)
            ma = None if mask_a is None else buf["ma"][: n * T_a].view(n, T_a).clone()
            mt = None if mask_t is None else buf["mt"][: n * T_t].view(n, T_t).clone()
            va = buf["a"][: n * T_a * d_a].view(n, T_a, d_a)
            vt = buf["t"][: n * T_t * d_t].view(n, T_t, d_t)
            if not from_host:
                # fp32 features are only read by the bf16 cast: after it the staging set is free again
                xa = E.to_seq(va, "h_a").x.view(n, T_a, -1)
                xt = E.to_seq(vt, "h_t").x.view(n, T_t, -1)
                buf["consumed"].record(main)
                emit(i, model(xa, xt, ma, mt)[:3])
            else:
                # bf16 landed on the device: the model reads it in place; the set is free after the forward
                emit(i, model(va, vt, ma, mt)[:3])
                buf["consumed"].record(main)
            if trace is not None:
                k1 = torch.cuda.Event(enable_timing=True)
                k1.record(main)
                for row in trace:
                    if row["slab"] == i:
                        row["comp0"], row["comp1"] = k0, k1
            nxt = next_stage()
            if nxt is not None:
                pending.append(nxt)
        worker.join()
        STATS["calls"] += 1
        rates[(n_sl, "pack")], rates[(n_sl, "step")] = plan.t_pack, plan.t_step
    except BaseException as e:
        plan.host_done(e)
        for ev in host_enq:
            ev.set()          # never leave the worker waiting on the main thread
        raise
    if to_host:
        if not wait:
            done = torch.cuda.Event()
            done.record(main)
            return PendingResult(host_out, done)
        main.synchronize()
        return tuple(host_out)
    order = sorted(outs)
    return tuple(torch.cat([outs[i][j] for i in order]).to(out_device) for j in range(3))


@torch.no_grad()
def forward_from_host(model, h_a: torch.Tensor, h_t: torch.Tensor, mask_a: Optional[torch.Tensor] = None,
                      mask_t: Optional[torch.Tensor] = None, device="cuda", slab: int = 512,
                      out_device="cpu", host_cast_every=None, ramp: bool = False, bucket: bool = False,
                      trace: Optional[list] = None, wait: bool = True):
    """model(h_a, h_t, mask_a, mask_t) for HOST tensors with copy/compute overlap.

    The batch is cut into slabs of utterances that flow through fixed staging buffers (nothing is
    allocated or freed per slab for the inputs) on a side stream while earlier slabs compute:
      * fp32 slabs are copied as they are and cast to bf16 on the GPU; their staging set is free again
        as soon as the cast has read it;
      * host_cast_every=None / "auto" (default; dense fp32 batches of three or more slabs): the host cores
        convert slabs to bf16 from the BACK of the batch while the copy engine sends fp32 slabs from the
        front, and they meet wherever their speeds put them (`_run_dense_dynamic`); what crossed PCIe is
        accumulated in `pipeline.STATS`;
      * host_cast_every=k (an int; 0 = never): the fixed plan -- every k-th slab is converted to bf16 by the
        host cores (hriemo_host_pack_bf16 in a worker thread) into pinned bf16 staging and copied at half the
        bytes.  A step is bounded by the 55 GB/s H2D copy of the fp32 features (7.1 GB at the
        north-star batch); with every second slab pre-cast the copy drops under the compute time.
        The rounding is the same round-to-nearest-even as the GPU cast: results are bit-identical.
      * bucket=True (needs at least one mask): utterances are sorted by valid length and every slab is
        trimmed to its own maxima (`bucket_plan`, `slab` x T_a audio rows per slab, the first two slabs
        smaller).  Every slab is gathered, trimmed and cast by the host cores in one pass
        (hriemo_host_pack_bf16) and travels as bf16; the masks go to the device once and each slab
        gathers its own; results are scattered back to the original order on the device.
    Returns (logits, beta, z) on `out_device`; host results are fresh PINNED tensors filled by
    asynchronous D2H copies (a pageable `.cpu()` of the 50 MB of z cost 31 ms per step).
    wait=False (host outputs only) returns a PendingResult instead of synchronising: a stream of batches is
    then pipelined two deep -- the next call's first slab is copied while this call's last slabs compute, which
    hides the one latency a single call cannot (nothing computes until its first slab has landed).
    trace (optional list): receives one dict of CUDA events per slab (tools/e2e_timeline.py)."""
    dev = torch.device(device)
    B = h_a.shape[0]
    if h_a.dim() != 3 or h_t.dim() != 3:
        # utterance-level [B, d] inputs: nothing to pipeline
        out = model(h_a.to(dev), h_t.to(dev), None if mask_a is None else mask_a.to(dev),
                    None if mask_t is None else mask_t.to(dev))[:3]
        return tuple(o.to(out_device) for o in out)
    T_a, d_a = h_a.shape[1], h_a.shape[2]
    T_t, d_t = h_t.shape[1], h_t.shape[2]
    slab = max(1, min(slab, B))
    h_a, h_t = h_a.contiguous(), h_t.contiguous()
    early = (h_a.dtype == torch.float32 and h_t.dtype == torch.float32 and d_a % 8 == 0 and d_t % 8 == 0)
    bucket = bool(bucket and early and (mask_a is not None or mask_t is not None) and T_a > 1)
    mask_a = None if mask_a is None else mask_a.to(torch.bool).contiguous()
    mask_t = None if mask_t is None else mask_t.to(torch.bool).contiguous()
    threads = max(1, torch.get_num_threads())
    if host_cast_every is None or host_cast_every == "auto":
        if early and not bucket and not ramp and (B + slab - 1) // slab >= 3:
            # the converting threads leave HOST_RESERVE cores to the thread that issues the copies and launches (and to the
            # CUDA driver's own): with all 16 busy converting, the issuing thread was descheduled between launches
            conv = threads - HOST_RESERVE if threads >= 8 else threads
            return _run_dense_dynamic(model, dev, h_a, h_t, mask_a, mask_t, slab, out_device, conv, wait, trace)
        host_cast_every = default_host_cast_every(threads)
    mask_a_dev = mask_t_dev = None

    if bucket:
        len_a = valid_lengths(mask_a, B, T_a).clamp_(min=1)
        len_t = valid_lengths(mask_t, B, T_t).clamp_(min=1)
        order, buckets = bucket_plan(len_a, len_t, T_a, T_t, rows_per_slab=slab * T_a, max_utts=4 * slab, ramp=True)
        order_dev = order.to(dev, non_blocking=True)
        len_a_s, len_t_s = len_a[order.long()].contiguous(), len_t[order.long()].contiguous()
        # every slab is gathered, trimmed and cast by the host cores and travels as bf16.  (Sending every
        # second slab as fp32 with one async copy per utterance, valid rows only, was slower: 8 192 copies
        # of ~1 MB / ~0.15 MB per step run the copy engine at 30 GB/s instead of 55 --
        # profiles/r01_e2e_timeline_v11.txt.)
        slabs = [_Slab(b.start, b.end, b.T_a, b.T_t, True, order[b.start:b.end].contiguous(), order_dev[b.start:b.end])
                 for b in buckets]
        mask_a_dev = None if mask_a is None else mask_a.to(dev, non_blocking=True)
        mask_t_dev = None if mask_t is None else mask_t.to(dev, non_blocking=True)
    else:
        slabs = [_Slab(s, e, T_a, T_t, hc) for s, e, hc in slab_schedule(B, slab, host_cast_every if early else 0, ramp)]

    def pack(s, hb):
        if s.utt is None:
            ops.host_pack_bf16(h_a[s.lo:s.hi], hb["a"], s.T_a, threads=threads)
            ops.host_pack_bf16(h_t[s.lo:s.hi], hb["t"], s.T_t, threads=threads)
        else:
            ops.host_pack_bf16(h_a, hb["a"], s.T_a, s.utt, len_a_s[s.lo:s.hi], threads=threads)
            ops.host_pack_bf16(h_t, hb["t"], s.T_t, s.utt, len_t_s[s.lo:s.hi], threads=threads)

    def direct_src(s):
        return h_a[s.lo:s.hi].view(-1), h_t[s.lo:s.hi].view(-1)

    return _run_pipeline(model, dev, B, slabs, d_a, d_t, h_a.dtype, h_t.dtype, torch.bfloat16, pack, direct_src,
                         mask_a, mask_t, mask_a_dev, mask_t_dev, early, out_device, trace, wait)


@torch.no_grad()
def forward_from_shard(model, shard, device="cuda", slab_rows: int = 512 * 500, max_utts: int = 2048,
                       out_device="cpu", trace: Optional[list] = None, wait: bool = True):
    """The forward over every utterance of a packed feature shard (hriemo.shards.Shard),
    results in SHARD order (shard.original_order() maps back to the writer's input order).

    The shard stores each utterance's valid rows back to back, so nothing is converted on the way: the plan
    is `bucket_plan` over the stored lengths (a shard written with sort_by_length=True is already in plan
    order), a worker thread copies each slab's rows from the mapped file into pinned staging, zero-padded to
    the slab's own extents (hriemo_shard_read), and the slabs flow through the same staging engine as
    forward_from_host(bucket=True): bf16 shards are read by the model in place, fp32 shards are cast on the
    GPU.  The True = PAD masks are rebuilt from the stored PAD bytes once and gathered per slab on the device."""
    dev = torch.device(device)
    B = len(shard)
    T_a, T_t = max(1, shard.max_len_a), max(1, shard.max_len_t)
    d_a, d_t = shard.d_a, shard.d_t
    if d_a % 8 or d_t % 8:
        raise L.HriemoError("forward_from_shard: feature dims must be multiples of 8 (pad them when writing the shard)")
    threads = max(1, torch.get_num_threads())
    order, buckets = bucket_plan(shard.len_a, shard.len_t, T_a, T_t, rows_per_slab=slab_rows, max_utts=max_utts, ramp=True)
    order64 = order.long()
    order_dev = order.to(dev, non_blocking=True)
    slabs = [_Slab(b.start, b.end, b.T_a, b.T_t, True, order64[b.start:b.end].contiguous(), order_dev[b.start:b.end])
             for b in buckets]
    # the whole masks (rows x 1 byte) once: host -> pinned -> device
    ma_h = torch.empty((B, T_a), dtype=torch.bool).pin_memory()
    mt_h = torch.empty((B, T_t), dtype=torch.bool).pin_memory()
    shard.read(first=0, n=B, T_a=T_a, T_t=T_t, out_mask_a=ma_h, out_mask_t=mt_h, features=False, threads=threads)
    mask_a_dev, mask_t_dev = ma_h.to(dev, non_blocking=True), mt_h.to(dev, non_blocking=True)

    def pack(s, hb):
        shard.read(utt=s.utt, T_a=s.T_a, T_t=s.T_t, out_a=hb["a"], out_t=hb["t"], masks=False, threads=threads)

    return _run_pipeline(model, dev, B, slabs, d_a, d_t, shard.dtype, shard.dtype, shard.dtype, pack, None,
                         None, None, mask_a_dev, mask_t_dev, True, out_device, trace, wait)
