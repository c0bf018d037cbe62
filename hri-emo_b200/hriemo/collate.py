"""The reference's collate functions, producing PINNED host tensors.

Same signatures, outputs, dtypes and padding conventions as
  collate_seq_batch(batch, loss_type)   scripts/fusion/train_fusion_seq_level_decoder.py:191-232   (IEMOCAP)
  collate_seq_batch(batch)              scripts/infer/mosei_eval_infer.py:128-147                   (MOSEI)
-- zero-pad every utterance to the batch maximum, masks default to True = PAD.  Pass as `collate_fn=` to a DataLoader.

Pinning.  Called in the MAIN process (num_workers=0, or a collate applied by hand) the padded tensors are allocated
in page-locked memory (torch's caching host allocator), so hriemo.pipeline.forward_from_host and a plain
`.to(device, non_blocking=True)` copy them asynchronously.  Inside a DataLoader WORKER process (the reference runs
num_workers=4, train_fusion_seq_level_decoder.py:275-282) nothing is pinned: a page-locked allocation needs a CUDA
context, which a forked worker must not create ("Cannot re-initialize CUDA in forked subprocess"), and the property
would be lost anyway when the batch travels back through shared memory.  There the main process pins:
`DataLoader(..., num_workers=4, collate_fn=collate_seq_batch, pin_memory=True)` (torch's pin thread), or
`pin_batch(batch)` below."""
from __future__ import annotations

import torch
from torch.utils.data import get_worker_info


def _in_worker() -> bool:
    return get_worker_info() is not None


def _pinned(shape, dtype, fill):
    pin = (not _in_worker()) and torch.cuda.is_available()
    t = torch.empty(shape, dtype=dtype, pin_memory=pin)
    return t.fill_(fill)


def pin_batch(batch):
    """Main-process pin step for batches collated in worker processes: every tensor of the tuple that is not yet
    page-locked is copied into pinned memory (no-op without a CUDA device)."""
    if not torch.cuda.is_available():
        return batch
    return tuple(x.pin_memory() if isinstance(x, torch.Tensor) and not x.is_pinned() else x for x in batch)


def collate_seq_batch(batch, loss_type: str = "multi_label"):
    """list of (h_a [L_a,d], m_a [L_a] bool, h_t [L_t,d], m_t [L_t] bool, label) ->
    (h_a [B,L_a_max,d], mask_a [B,L_a_max], h_t [B,L_t_max,d], mask_t [B,L_t_max], labels)."""
    hs_a, ms_a, hs_t, ms_t, labels = zip(*batch)
    B = len(batch)
    d_a, d_t = hs_a[0].size(-1), hs_t[0].size(-1)
    La, Lt = max(x.size(0) for x in hs_a), max(x.size(0) for x in hs_t)
    h_a, h_t = _pinned((B, La, d_a), torch.float32, 0.0), _pinned((B, Lt, d_t), torch.float32, 0.0)
    m_a, m_t = _pinned((B, La), torch.bool, True), _pinned((B, Lt), torch.bool, True)
    for i in range(B):
        a, t = hs_a[i].size(0), hs_t[i].size(0)
        h_a[i, :a] = hs_a[i]
        m_a[i, :a] = ms_a[i]
        h_t[i, :t] = hs_t[i]
        m_t[i, :t] = ms_t[i]
    if loss_type == "single_label":
        labels = torch.tensor(labels, dtype=torch.long)
    else:
        labels = torch.stack(labels, dim=0)
    return h_a, m_a, h_t, m_t, labels


def collate_mosei_batch(batch):
    """list of (a [L_a,d_a] float32, t [L_t,d_t] float32, y [C]) -> (h_a, m_a, h_t, m_t, y); every stored row is
    valid (mask False), the padding is True (mosei_eval_infer.py:128-147)."""
    As, Ts, Ys = zip(*batch)
    B = len(batch)
    La, Lt = max(x.shape[0] for x in As), max(x.shape[0] for x in Ts)
    h_a, h_t = _pinned((B, La, As[0].shape[1]), torch.float32, 0.0), _pinned((B, Lt, Ts[0].shape[1]), torch.float32, 0.0)
    m_a, m_t = _pinned((B, La), torch.bool, True), _pinned((B, Lt), torch.bool, True)
    for i, (a, t) in enumerate(zip(As, Ts)):
        la, lt = a.shape[0], t.shape[0]
        h_a[i, :la] = a
        m_a[i, :la] = False
        h_t[i, :lt] = t
        m_t[i, :lt] = False
    return h_a, m_a, h_t, m_t, torch.stack(Ys, dim=0)
