"""Length-bucketed staging, host side: the plan, the host pack function of the
C-ABI library, and -- through the oracle, against the reference's own golden outputs -- the claim the
whole feature rests on: trimming a slab of utterances to its own maximum valid length changes no
logits / beta / z."""
import pytest
import torch

import golden_util as G
import hriemo_oracle as O
from hriemo import ops, pipeline


def _ragged(B, T, gen, lo=1):
    lens = torch.randint(lo, T + 1, (B,), generator=gen)
    return torch.arange(T)[None, :] >= lens[:, None], lens


def test_valid_lengths_host_masks():
    m = torch.tensor([[0, 0, 1, 1], [1, 1, 1, 1], [0, 1, 0, 1], [0, 0, 0, 0]], dtype=torch.bool)
    assert pipeline.valid_lengths(m, 4, 4).tolist() == [2, 0, 3, 4]       # holes count up to the last valid position
    assert pipeline.valid_lengths(None, 3, 7).tolist() == [7, 7, 7]


@pytest.mark.parametrize("B,T_a,T_t,rows,max_utts", [(1, 5, 3, 100, 8), (37, 300, 50, 3000, 16), (200, 500, 64, 512 * 500, 2048),
                                                      (64, 40, 128, 2000, 7)])
def test_bucket_plan_is_a_trimmed_partition(B, T_a, T_t, rows, max_utts):
    g = torch.Generator().manual_seed(B)
    _, la = _ragged(B, T_a, g, lo=0)
    _, lt = _ragged(B, T_t, g, lo=0)
    order, buckets = pipeline.bucket_plan(la, lt, T_a, T_t, rows, max_utts)
    assert sorted(order.tolist()) == list(range(B))
    assert buckets[0].start == 0 and buckets[-1].end == B
    assert all(a.end == b.start for a, b in zip(buckets, buckets[1:]))
    for bk in buckets:
        idx = order[bk.start:bk.end].long()
        n = bk.end - bk.start
        assert 0 < n <= max_utts
        assert int(la[idx].max()) <= bk.T_a <= T_a or bk.T_a == min(T_a, bk.T_t)   # nothing valid is cut
        assert int(lt[idx].max()) <= bk.T_t <= T_t
        assert bk.T_a >= 1 and bk.T_t >= 1
        assert bk.T_a >= min(bk.T_t, T_a)                                          # the gate needs T_a >= T_t
        assert n == 1 or n * max(int(la[idx].clamp(min=1).max()), int(lt[idx].clamp(min=1).max())) <= rows
    # ascending audio length: the plan opens with the cheap slabs
    las = la.clamp(1, T_a)[order.long()]
    assert bool((las[1:] >= las[:-1]).all())
    st = pipeline.bucket_stats(la, lt, T_a, T_t, buckets)
    assert st["valid_rows"] <= st["bucketed_rows"] + B * max(0, T_t - 1) and st["bucketed_rows"] <= st["padded_rows"] + B * T_t


def test_shard_by_length_balances_counts_and_work():
    g = torch.Generator().manual_seed(3)
    _, la = _ragged(1000, 500, g, lo=50)
    _, lt = _ragged(1000, 64, g, lo=4)
    shards = pipeline.shard_by_length(la, lt, 8)
    assert sorted(torch.cat(shards).tolist()) == list(range(1000))
    assert {len(s) for s in shards} == {125}
    work = [float(((la[s] + lt[s]).double() ** 2).sum()) for s in shards]
    assert max(work) / min(work) < 1.03
    # contiguous sharding of the same (unsorted) batch is worse or equal on the quadratic term only by luck;
    # the sorted deal also equalises the maximum length seen by every rank's buckets
    assert max(int(la[s].max()) for s in shards) - min(int(la[s].max()) for s in shards) <= 5


def test_host_pack_matches_torch_cast_and_trims():
    lib_ok = True
    try:
        from hriemo import lib
        lib.load()
    except Exception:
        lib_ok = False
    assert lib_ok, "libhriemo_b200.so must be built (python __graft_entry__.py)"
    g = torch.Generator().manual_seed(0)
    B, T, d = 9, 37, 72
    x = torch.randn(B, T, d, generator=g)
    x[0, 0, 0], x[0, 0, 1], x[0, 0, 2], x[0, 0, 3] = float("inf"), -0.0, 1e-40, 3.3895314e38
    ref = x.to(torch.bfloat16)
    dst = torch.empty(B * T * d, dtype=torch.bfloat16)
    out = ops.host_pack_bf16(x, dst, T, threads=3)
    assert torch.equal(out.view(torch.int16), ref.view(torch.int16))           # round to nearest even, bit for bit
    utt = torch.tensor([4, 0, 8, 8, 2], dtype=torch.int32)
    lens = torch.tensor([10, 37, 1, 20, 5], dtype=torch.int32)
    for T_out in (20, 37, 45):
        dst = torch.full((5 * T_out * d,), 7.0, dtype=torch.bfloat16)
        out = ops.host_pack_bf16(x, dst, T_out, utt, lens, threads=2)
        assert out.shape == (5, T_out, d)
        for i in range(5):
            keep = min(int(lens[i]), T_out, T)
            assert torch.equal(out[i, :keep].view(torch.int16), ref[utt[i], :keep].view(torch.int16))
            assert bool((out[i, keep:] == 0).all())
    with pytest.raises(Exception):
        ops.host_pack_bf16(x, torch.empty(10, dtype=torch.bfloat16), T)        # destination too small
    # NaN payloads become the canonical bf16 NaN of the GPU cast; a row count that is not a multiple of 16 and an
    # unaligned destination exercise the vector path's tail and its cached-store form
    y = torch.randn(3, 5, 40, generator=g)
    y[1, 2, 7] = float("nan")
    dst = torch.empty(3 * 5 * 40 + 8, dtype=torch.bfloat16)
    out = ops.host_pack_bf16(y, dst[3:], 5, threads=1)
    want = y.to(torch.bfloat16)
    ok = torch.isnan(want)
    assert torch.equal(out.view(torch.int16)[~ok], want.view(torch.int16)[~ok]) and bool(torch.isnan(out[ok]).all())
    # an utterance index outside the source batch is refused by the library itself (ABI check, not only the wrapper)
    from hriemo import lib
    bad = torch.tensor([0, 9], dtype=torch.int32)
    rc = lib.load().hriemo_host_pack_bf16(x.data_ptr(), d, T, d, bad.data_ptr(), None, dst.data_ptr(), d, 1, 2, B, 1)
    assert rc != 0 and b"outside the source batch" in lib.load().hriemo_last_error()


@pytest.mark.parametrize("name", ["cfg2_iemocap_ragged", "ns_500x64_ragged", "cfg3_mosei_default"])
def test_trimmed_buckets_reproduce_the_reference_outputs(name):
    """The reference's outputs for the padded batch (golden fixture) are reproduced by running each
    bucket of the plan on its own, trimmed to the bucket's maxima (oracle, fp64)."""
    fx = G.load(name)
    model, (h_a, h_t, m_a, m_t) = G.build_fusion(fx)
    sd = O.cast_state(model.state_dict(), torch.float64)
    H = G.n_heads_of(fx)
    fwd = O.mosei_fusion_with_emotion_decoder if fx["kind"] == "mosei" else O.fusion_with_emotion_decoder
    B, T_a, T_t = h_a.shape[0], h_a.shape[1], h_t.shape[1]
    la, lt = pipeline.valid_lengths(m_a, B, T_a), pipeline.valid_lengths(m_t, B, T_t)
    order, buckets = pipeline.bucket_plan(la, lt, T_a, T_t, rows_per_slab=B * T_a, max_utts=max(1, B // 2))
    assert len(buckets) >= 2
    assert any(b.T_a < T_a or b.T_t < T_t for b in buckets), "fixture has no padding to trim"
    lo = torch.empty_like(fx["logits"], dtype=torch.float64)
    be = torch.empty_like(fx["beta"], dtype=torch.float64)
    z = torch.empty_like(fx["z"], dtype=torch.float64)
    for bk in buckets:
        idx = order[bk.start:bk.end].long()
        r = fwd(sd, h_a[idx, :bk.T_a].double(), h_t[idx, :bk.T_t].double(), m_a[idx, :bk.T_a], m_t[idx, :bk.T_t], n_heads=H)
        lo[idx], be[idx], z[idx] = r[0], r[1], r[2]
    assert (lo - fx["logits"]).abs().max().item() < 2e-5
    assert (be - fx["beta"]).abs().max().item() < 2e-5
    assert (z - fx["z"]).abs().max().item() < 1e-4


def test_bucket_plan_ramp_opens_with_small_slabs():
    g = torch.Generator().manual_seed(9)
    _, la = _ragged(4096, 500, g, lo=250)
    _, lt = _ragged(4096, 64, g, lo=32)
    rows = 512 * 500
    order, plain = pipeline.bucket_plan(la, lt, 500, 64, rows, 2048)
    order2, ramped = pipeline.bucket_plan(la, lt, 500, 64, rows, 2048, ramp=True)
    assert torch.equal(order, order2)
    size = lambda b: (b.end - b.start) * b.T_a
    assert size(ramped[0]) <= rows // 4 and size(ramped[1]) <= rows // 2 and size(ramped[2]) > rows // 2
    assert size(plain[0]) > rows // 2
    assert ramped[-1].end == 4096 and all(a.end == b.start for a, b in zip(ramped, ramped[1:]))
    # a batch that fits a few slabs gets no ramp
    _, small = pipeline.bucket_plan(la[:600], lt[:600], 500, 64, rows, 2048, ramp=True)
    assert size(small[0]) > rows // 2
