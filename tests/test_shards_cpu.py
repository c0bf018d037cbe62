"""Packed feature shards: the writer, the converter from the reference's
per-utterance pickles and the C-ABI reader, checked against the reference's own load + collate
(scripts/fusion/train_fusion_seq_level_decoder.py:139-156, :191-232) restated here."""
import os

import pytest
import torch

from hriemo import lib, shards


def _utterances(n, d_a, d_t, gen, max_a=40, max_t=12):
    items = []
    for i in range(n):
        La, Lt = int(torch.randint(1, max_a + 1, (1,), generator=gen)), int(torch.randint(1, max_t + 1, (1,), generator=gen))
        va, vt = int(torch.randint(1, La + 1, (1,), generator=gen)), int(torch.randint(1, Lt + 1, (1,), generator=gen))
        h_a, h_t = torch.randn(La, d_a, generator=gen), torch.randn(Lt, d_t, generator=gen)
        items.append((h_a, torch.arange(La) >= va, h_t, torch.arange(Lt) >= vt))
    return items


def _reference_collate(batch):
    """collate_seq_batch of the reference (zero-pad to the batch maximum, default mask True = PAD)."""
    hs_a, ms_a, hs_t, ms_t = zip(*batch)
    B, d_a, d_t = len(batch), hs_a[0].size(-1), hs_t[0].size(-1)
    La, Lt = max(x.size(0) for x in hs_a), max(x.size(0) for x in hs_t)
    pa, pt = torch.zeros(B, La, d_a), torch.zeros(B, Lt, d_t)
    ma, mt = torch.ones(B, La, dtype=torch.bool), torch.ones(B, Lt, dtype=torch.bool)
    for i in range(B):
        a, t = hs_a[i].size(0), hs_t[i].size(0)
        pa[i, :a], ma[i, :a], pt[i, :t], mt[i, :t] = hs_a[i], ms_a[i], hs_t[i], ms_t[i]
    return pa, ma, pt, mt


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("sort", [False, True])
def test_shard_roundtrip_equals_reference_collate(tmp_path, dtype, sort):
    g = torch.Generator().manual_seed(0)
    items = _utterances(23, 16, 24, g)
    items[3] = (items[3][0], torch.ones(items[3][0].shape[0], dtype=torch.bool), items[3][2], items[3][3])   # all PAD
    hole = items[5][1].clone(); hole[0] = True
    items[5] = (items[5][0], hole, items[5][2], items[5][3])                                                  # a hole
    items[7] = (items[7][0], None, items[7][2], None)                                                         # no masks
    path = str(tmp_path / "s.hriemo")
    info = shards.write_shard(path, items, dtype=dtype, sort_by_length=sort, uids=[f"u{i}" for i in range(23)],
                              labels=[i % 4 for i in range(23)])
    with shards.Shard(path) as sh:
        assert len(sh) == 23 and sh.d_a == 16 and sh.d_t == 24 and sh.dtype == dtype
        order = sh.original_order()
        assert sorted(order.tolist()) == list(range(23))
        assert sh.meta["uids"] == [f"u{i}" for i in order.tolist()] and sh.meta["labels"] == [i % 4 for i in order.tolist()]
        if sort:
            keys = [(int(a), int(t)) for a, t in zip(sh.len_a, sh.len_t)]
            assert keys == sorted(keys)
        # the whole shard as one padded batch == the reference's collate of the same utterances (shard order),
        # up to the PAD tail the shard does not store: extents are the longest VALID lengths
        a, t, ma, mt = sh.read()
        def full(it):
            h_a, p_a, h_t, p_t = it
            return (h_a, torch.zeros(h_a.shape[0], dtype=torch.bool) if p_a is None else p_a,
                    h_t, torch.zeros(h_t.shape[0], dtype=torch.bool) if p_t is None else p_t)
        batch = [full(items[i]) for i in order.tolist()]
        pa, pma, pt, pmt = _reference_collate(batch)
        Ta, Tt = a.shape[1], t.shape[1]
        assert Ta == sh.max_len_a and Tt == sh.max_len_t
        assert torch.equal(ma, pma[:, :Ta]) and torch.equal(mt, pmt[:, :Tt])
        assert bool(pma[:, Ta:].all()) and bool(pmt[:, Tt:].all())            # only PAD was dropped
        keep_a, keep_t = ~ma.unsqueeze(-1), ~mt.unsqueeze(-1)
        assert torch.equal(a.float() * keep_a, pa[:, :Ta].to(dtype).float() * keep_a)   # valid rows bit for bit
        assert torch.equal(t.float() * keep_t, pt[:, :Tt].to(dtype).float() * keep_t)
        # rows past an utterance's stored length are zero, like the collate's
        for k in range(23):
            assert bool((a[k, int(sh.len_a[k]):] == 0).all()) and bool((t[k, int(sh.len_t[k]):] == 0).all())
        # an index list, trimmed extents, caller-provided buffers
        utt = torch.tensor([22, 0, 7, 7], dtype=torch.int64)
        buf_a = torch.empty(4 * 9 * 16 + 5, dtype=dtype)
        a2, t2, ma2, mt2 = sh.read(utt=utt, T_a=9, T_t=30, out_a=buf_a, threads=3)
        assert a2.shape == (4, 9, 16) and t2.shape == (4, 30, 24) and a2.data_ptr() == buf_a.data_ptr()
        assert torch.equal(a2, a[utt][:, :9] if Ta >= 9 else torch.cat([a[utt], torch.zeros(4, 9 - Ta, 16, dtype=dtype)], 1))
        assert torch.equal(mt2[:, :Tt], mt[utt]) and bool(mt2[:, Tt:].all())
        feats_only = sh.read(first=2, n=3, masks=False)
        assert feats_only[2] is None and feats_only[0].shape[0] == 3
    assert info["n_utt"] == 23


def test_convert_reference_dirs(tmp_path):
    """Files written exactly like the reference's extractors write them (hidden + attention_mask, 1 = valid)."""
    g = torch.Generator().manual_seed(1)
    ad, td = tmp_path / "audio", tmp_path / "text"
    ad.mkdir(); td.mkdir()
    uids = [f"Ses01F_impro01_F{i:03d}" for i in range(6)]
    truth = {}
    for u in uids[:5]:   # the sixth uid has no features: skipped, like the reference's Dataset
        La, Lt = int(torch.randint(2, 30, (1,), generator=g)), int(torch.randint(2, 10, (1,), generator=g))
        ha, ht = torch.randn(La, 8, generator=g), torch.randn(Lt, 8, generator=g)
        am, tm = torch.ones(La, dtype=torch.long), torch.ones(Lt, dtype=torch.long)
        am[La - 1:] = 0
        torch.save({"hidden": ha, "attention_mask": am}, ad / f"{u}.pt")
        torch.save({"hidden": ht, "attention_mask": tm}, td / f"{u}.pt")
        truth[u] = (ha, am == 0, ht, tm == 0)
    path = str(tmp_path / "ref.hriemo")
    shards.convert_reference_dirs(str(ad), str(td), uids, path, labels=list(range(6)), dtype=torch.float32)
    with shards.Shard(path) as sh:
        assert len(sh) == 5 and sorted(sh.meta["uids"]) == uids[:5]
        a, t, ma, mt = sh.read()
        for k, u in enumerate(sh.meta["uids"]):
            ha, pa, ht, pt = truth[u]
            L = ha.shape[0] - 1                                   # the PAD tail is not stored
            assert int(sh.len_a[k]) == L and int(sh.len_t[k]) == ht.shape[0]
            assert torch.equal(a[k, :L], ha[:L]) and torch.equal(t[k, :ht.shape[0]], ht)
            assert not ma[k, :L].any() and bool(ma[k, L:].all())
            assert sh.meta["labels"][k] == uids.index(u)


def test_shard_reader_rejects_bad_files(tmp_path):
    g = torch.Generator().manual_seed(2)
    path = str(tmp_path / "ok.hriemo")
    shards.write_shard(path, _utterances(4, 8, 8, g))
    raw = open(path, "rb").read()
    for name, data in (("trunc", raw[:-10]), ("magic", b"NOTASHRD" + raw[8:]), ("tiny", raw[:50]),
                       ("rows", raw[:32] + (10 ** 9).to_bytes(8, "little") + raw[40:]),
                       # sizes whose 64-bit products wrap: rows_a * d_a * elem == 0 (mod 2^64), n_utt * 32 == 128 (mod 2^64)
                       ("wrap_rows", raw[:32] + (1 << 61).to_bytes(8, "little") + raw[40:]),
                       ("wrap_utts", raw[:16] + ((1 << 59) + 4).to_bytes(8, "little") + raw[24:]),
                       # an index entry whose row_a + len_a wraps past 2^64
                       ("wrap_index", raw[:128] + ((1 << 64) - 1).to_bytes(8, "little") + raw[136:])):
        bad = str(tmp_path / name)
        open(bad, "wb").write(data)
        with pytest.raises(lib.HriemoError):
            shards.Shard(bad)
    with pytest.raises(lib.HriemoError):
        shards.Shard(str(tmp_path / "missing"))
    with shards.Shard(path) as sh:
        with pytest.raises(lib.HriemoError):
            sh.read(first=3, n=2)
        with pytest.raises(lib.HriemoError):
            sh.read(utt=torch.tensor([4], dtype=torch.int64))


def test_collate_mirrors_the_reference_collates():
    from hriemo import collate

    g = torch.Generator().manual_seed(3)
    items = _utterances(7, 8, 12, g)
    labels = [torch.eye(4)[i % 4] for i in range(7)]
    h_a, m_a, h_t, m_t, y = collate.collate_seq_batch([it + (lab,) for it, lab in zip(items, labels)], "multi_label")
    pa, pma, pt, pmt = _reference_collate(items)
    assert torch.equal(h_a, pa) and torch.equal(m_a, pma) and torch.equal(h_t, pt) and torch.equal(m_t, pmt)
    assert h_a.dtype == torch.float32 and m_a.dtype == torch.bool and y.shape == (7, 4) and y.dtype == torch.float32
    y1 = collate.collate_seq_batch([it + (i % 4,) for i, it in enumerate(items)], "single_label")[4]
    assert y1.dtype == torch.long and y1.tolist() == [i % 4 for i in range(7)]
    mos = [(it[0], it[2], lab) for it, lab in zip(items, labels)]
    h_a2, m_a2, h_t2, m_t2, y2 = collate.collate_mosei_batch(mos)
    assert torch.equal(h_a2, pa) and torch.equal(h_t2, pt)
    for i, it in enumerate(items):
        assert not m_a2[i, :it[0].shape[0]].any() and bool(m_a2[i, it[0].shape[0]:].all())
        assert not m_t2[i, :it[2].shape[0]].any() and bool(m_t2[i, it[2].shape[0]:].all())


class _ListDataset(torch.utils.data.Dataset):
    def __init__(self, items):
        self.items = items

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


def test_collate_in_dataloader_worker_processes(monkeypatch):
    """The reference's loaders run the collate in worker PROCESSES (num_workers=4): there nothing may be pinned (a
    page-locked allocation would need a CUDA context in a forked child) even when the parent sees a GPU; the batches
    must equal the in-process collate's."""
    from functools import partial

    from hriemo import collate

    g = torch.Generator().manual_seed(5)
    items = _utterances(10, 8, 12, g)
    data = [it + (torch.eye(4)[i % 4],) for i, it in enumerate(items)]
    # pretend the parent process has a GPU: a worker must still not try to pin
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    seen_pin = []
    real_empty = torch.empty

    def spy_empty(*a, **kw):
        seen_pin.append(bool(kw.get("pin_memory", False)))
        kw["pin_memory"] = False
        return real_empty(*a, **kw)

    monkeypatch.setattr(collate, "_in_worker", lambda: True)
    monkeypatch.setattr(torch, "empty", spy_empty)
    collate.collate_seq_batch(data[:3], "multi_label")
    assert seen_pin and not any(seen_pin)                      # worker context: pageable allocations only
    monkeypatch.undo()
    loader = torch.utils.data.DataLoader(_ListDataset(data), batch_size=4, shuffle=False, num_workers=2,
                                         collate_fn=partial(collate.collate_seq_batch, loss_type="multi_label"))
    got = list(loader)
    assert len(got) == 3
    for k, batch in enumerate(got):
        want = collate.collate_seq_batch(data[4 * k:4 * k + 4], "multi_label")
        assert all(torch.equal(a, b) for a, b in zip(batch, want))
        assert collate.pin_batch(batch)[0].shape == want[0].shape
