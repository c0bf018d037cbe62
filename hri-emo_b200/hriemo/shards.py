"""Packed feature shards (format HRIEMOS1): writer, reference-format converter and reader.

The reference keeps one torch pickle per utterance and modality
(scripts/iemocap_feature_extraction_seq_level/extract_audio_feats_wavlm_seq.py:118-135:
{"hidden": [L, d] float32, "attention_mask": [L] long, 1 = valid}) and its Dataset unpickles two files
per sample (scripts/fusion/train_fusion_seq_level_decoder.py:139-156, :173-186).  A shard packs many
utterances into one file that the C-ABI reader (hri-emo_b200/csrc/host_shard.cpp) maps and copies, slab
by slab, straight into pinned staging memory; the layout is documented there.

* ``write_shard``  items -> file (bf16 by default: half the disk, page-cache and PCIe bytes; the rounding
  is the forward path's own input cast, so results are bit-identical to feeding the fp32 features).
* ``convert_reference_dirs``  the reference's <audio_dir>/<uid>.pt + <text_dir>/<uid>.pt -> shard.
* ``Shard``  reader: lengths, metadata, ``read`` of a slab as the collate's padded tensors + True = PAD masks.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import struct
from typing import Iterable, Optional, Sequence, Tuple

import torch

from . import lib as _l

MAGIC = b"HRIEMOS1"
DTYPES = {torch.bfloat16: 1, torch.float32: 2}
_ALIGN = 4096
_HEADER = struct.Struct("<8sIIQIIQQQQQQQQQIIQ8x")   # 128 bytes, see host_shard.cpp
assert _HEADER.size == 128


def _valid_len(pad: Optional[torch.Tensor], L: int) -> int:
    if pad is None:
        return L
    valid = (~pad.to(torch.bool)).nonzero()
    return int(valid[-1].item()) + 1 if valid.numel() else 0


def write_shard(path: str, items: Iterable[Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor, Optional[torch.Tensor]]],
                dtype: torch.dtype = torch.bfloat16, sort_by_length: bool = True, uids: Optional[Sequence[str]] = None,
                labels: Optional[Sequence] = None, meta: Optional[dict] = None) -> dict:
    """items: (h_a [L_a, d_a], pad_a [L_a] bool True = PAD or None, h_t [L_t, d_t], pad_t or None) per utterance --
    what the reference's Dataset.__getitem__ returns (train_fusion_seq_level_decoder.py:173-188).
    Each utterance is stored up to its last valid position (the PAD tail is dropped, holes are kept with their
    PAD bytes).  sort_by_length: store utterances in ascending (len_a, len_t) order, so that a contiguous range
    of the shard is a length bucket; the original position of every stored utterance is kept in the metadata
    ("order").  Returns the summary dict that is also stored as metadata."""
    if dtype not in DTYPES:
        raise ValueError("shard dtype must be torch.bfloat16 or torch.float32")
    items = list(items)
    n = len(items)
    if n == 0:
        raise ValueError("write_shard: no utterances")
    d_a, d_t = items[0][0].shape[-1], items[0][2].shape[-1]
    lens = []
    for h_a, p_a, h_t, p_t in items:
        if h_a.dim() != 2 or h_t.dim() != 2 or h_a.shape[1] != d_a or h_t.shape[1] != d_t:
            raise ValueError("write_shard: every utterance must be ([L_a, d_a], [L_t, d_t]) with fixed feature dims")
        lens.append((_valid_len(p_a, h_a.shape[0]), _valid_len(p_t, h_t.shape[0])))
    order = sorted(range(n), key=lambda i: lens[i]) if sort_by_length else list(range(n))
    rows_a = sum(lens[i][0] for i in order)
    rows_t = sum(lens[i][1] for i in order)
    elem = 2 if dtype == torch.bfloat16 else 4

    def up(x):
        return (x + _ALIGN - 1) // _ALIGN * _ALIGN

    info = dict(meta or {})
    info.update({"format": "HRIEMOS1", "n_utt": n, "d_a": d_a, "d_t": d_t, "dtype": str(dtype).replace("torch.", ""),
                 "order": order, "sorted_by_length": bool(sort_by_length)})
    if uids is not None:
        info["uids"] = [str(uids[i]) for i in order]
    if labels is not None:
        info["labels"] = [labels[i] if not isinstance(labels[i], torch.Tensor) else labels[i].tolist() for i in order]
    meta_bytes = json.dumps(info).encode()

    off_index = _HEADER.size
    off_audio = up(off_index + 32 * n)
    off_text = up(off_audio + rows_a * d_a * elem)
    off_mask_a = up(off_text + rows_t * d_t * elem)
    off_mask_t = up(off_mask_a + rows_a)
    off_meta = up(off_mask_t + rows_t)
    file_bytes = off_meta + len(meta_bytes)
    header = _HEADER.pack(MAGIC, 1, DTYPES[dtype], n, d_a, d_t, rows_a, rows_t, off_index, off_audio, off_text,
                          off_mask_a, off_mask_t, off_meta, file_bytes, max(l[0] for l in lens), max(l[1] for l in lens),
                          len(meta_bytes))
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(header)
        ra = rt = 0
        for i in order:
            f.write(struct.pack("<QQIIQ", ra, rt, lens[i][0], lens[i][1], 0))
            ra += lens[i][0]
            rt += lens[i][1]

        def section(off, chunks):
            f.seek(off)
            for c in chunks:
                f.write(c)

        def rows(k, li):
            for i in order:
                x = items[i][k][: lens[i][li]].detach().to("cpu").to(dtype).contiguous()
                yield x.view(torch.uint8).numpy().tobytes() if x.numel() else b""

        def pads(k, li):
            for i in order:
                p = items[i][k]
                L = lens[i][li]
                yield (bytes(L) if p is None else p[:L].to(torch.uint8).contiguous().numpy().tobytes())

        section(off_audio, rows(0, 0))
        section(off_text, rows(2, 1))
        section(off_mask_a, pads(1, 0))
        section(off_mask_t, pads(3, 1))
        section(off_meta, [meta_bytes])
        f.truncate(file_bytes)
    os.replace(tmp, path)
    return info


def load_reference_feature(path: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """One file of the reference's feature extractors -> (hidden [L, d] float32, pad [L] bool, True = PAD):
    the same conversion as SeqLevelDataset._load_seq_feat (train_fusion_seq_level_decoder.py:139-156)."""
    obj = torch.load(path, map_location="cpu")
    return obj["hidden"].float(), obj["attention_mask"].long() == 0


def convert_reference_dirs(audio_dir: str, text_dir: str, uids: Sequence[str], path: str, labels: Optional[Sequence] = None,
                           dtype: torch.dtype = torch.bfloat16, sort_by_length: bool = True) -> dict:
    """<audio_dir>/<uid>.pt and <text_dir>/<uid>.pt for every uid (the reference's on-disk layout) -> one shard.
    uids without both files are skipped, like the reference's Dataset does (:126-133)."""
    keep = [k for k, u in enumerate(uids) if os.path.isfile(os.path.join(audio_dir, f"{u}.pt"))
            and os.path.isfile(os.path.join(text_dir, f"{u}.pt"))]

    def gen():
        for k in keep:
            h_a, p_a = load_reference_feature(os.path.join(audio_dir, f"{uids[k]}.pt"))
            h_t, p_t = load_reference_feature(os.path.join(text_dir, f"{uids[k]}.pt"))
            yield h_a, p_a, h_t, p_t

    return write_shard(path, gen(), dtype=dtype, sort_by_length=sort_by_length, uids=[uids[k] for k in keep],
                       labels=None if labels is None else [labels[k] for k in keep],
                       meta={"source": {"audio_dir": str(audio_dir), "text_dir": str(text_dir)}})


class Shard:
    """A mapped shard.  Lengths and metadata are read once; ``read`` fills caller-provided host buffers."""

    def __init__(self, path: str):
        self._h = C.c_void_p()
        self._lib = _l.load()
        _l.check(self._lib.hriemo_shard_open(os.fsencode(path), C.byref(self._h)), "shard_open")
        info = _l.ShardInfo()
        _l.check(self._lib.hriemo_shard_info(self._h, C.byref(info)), "shard_info")
        self.path = path
        self.n_utt, self.d_a, self.d_t = int(info.n_utt), int(info.d_a), int(info.d_t)
        self.rows_a, self.rows_t = int(info.rows_a), int(info.rows_t)
        self.max_len_a, self.max_len_t = int(info.max_len_a), int(info.max_len_t)
        self.dtype = torch.bfloat16 if info.dtype == 1 else torch.float32
        self.len_a = torch.empty(self.n_utt, dtype=torch.int32)
        self.len_t = torch.empty(self.n_utt, dtype=torch.int32)
        _l.check(self._lib.hriemo_shard_lengths(self._h, self.len_a.data_ptr(), self.len_t.data_ptr()), "shard_lengths")
        buf = C.create_string_buffer(max(1, int(info.meta_bytes)))
        _l.check(self._lib.hriemo_shard_meta(self._h, buf, int(info.meta_bytes)), "shard_meta")
        self.meta = json.loads(buf.raw[: int(info.meta_bytes)].decode()) if info.meta_bytes else {}

    def __len__(self) -> int:
        return self.n_utt

    def close(self) -> None:
        if self._h:
            self._lib.hriemo_shard_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def read(self, utt=None, first: int = 0, n: Optional[int] = None, T_a: Optional[int] = None, T_t: Optional[int] = None,
             out_a: Optional[torch.Tensor] = None, out_t: Optional[torch.Tensor] = None,
             out_mask_a: Optional[torch.Tensor] = None, out_mask_t: Optional[torch.Tensor] = None,
             features: bool = True, masks: bool = True, threads: int = 0):
        """Utterances `utt` (int64 host vector of shard positions) or [first, first + n) as
        (h_a [n, T_a, d_a], h_t [n, T_t, d_t], mask_a [n, T_a], mask_t [n, T_t]); T defaults to the longest
        utterance requested.  out_*: flat or shaped host buffers (e.g. pinned staging) to fill instead of
        allocating; features / masks = False skips that half (returns None for it)."""
        if utt is not None:
            if utt.is_cuda or utt.dtype != torch.int64 or not utt.is_contiguous():
                raise _l.HriemoError("Shard.read: utt must be a contiguous host int64 vector")
            n = int(utt.shape[0])
            if n and (int(utt.min()) < 0 or int(utt.max()) >= self.n_utt):
                raise _l.HriemoError(f"Shard.read: utterance index out of range [0, {self.n_utt})")
            la, lt = self.len_a[utt], self.len_t[utt]
        else:
            n = self.n_utt - first if n is None else n
            if first < 0 or n < 0 or first + n > self.n_utt:
                raise _l.HriemoError(f"Shard.read: range [{first}, {first + n}) outside the shard's {self.n_utt} utterances")
            la, lt = self.len_a[first:first + n], self.len_t[first:first + n]
        T_a = max(1, int(la.max())) if T_a is None and n else (T_a or 1)
        T_t = max(1, int(lt.max())) if T_t is None and n else (T_t or 1)

        def buf(out, numel, dtype):
            if out is None:
                return torch.empty(numel, dtype=dtype)
            if out.is_cuda or out.dtype != dtype or not out.is_contiguous() or out.numel() < numel:
                raise _l.HriemoError(f"Shard.read: output buffer must be a contiguous host {dtype} tensor with >= {numel} elements")
            return out.view(-1)[:numel]

        a = buf(out_a, n * T_a * self.d_a, self.dtype) if features else None
        t = buf(out_t, n * T_t * self.d_t, self.dtype) if features else None
        ma = buf(out_mask_a, n * T_a, torch.bool) if masks else None
        mt = buf(out_mask_t, n * T_t, torch.bool) if masks else None
        threads = threads or max(1, min(32, os.cpu_count() or 1))
        ptr = lambda x: None if x is None else x.data_ptr()
        _l.check(self._lib.hriemo_shard_read(self._h, None if utt is None else utt.data_ptr(), first, n, T_a, T_t,
                                             ptr(a), ptr(t), ptr(ma), ptr(mt), threads), "shard_read")
        return (None if a is None else a.view(n, T_a, self.d_a), None if t is None else t.view(n, T_t, self.d_t),
                None if ma is None else ma.view(n, T_a), None if mt is None else mt.view(n, T_t))

    def original_order(self) -> torch.Tensor:
        """order[k] = position, in the writer's input sequence, of the utterance stored at shard position k."""
        return torch.tensor(self.meta.get("order", list(range(self.n_utt))), dtype=torch.int64)
