"""Length-bucketed staging on the GPU: the staging kernels against torch
indexing (bit-exact: they move bytes or round once), and the bucketed forwards against the padded
forward, the reference's golden outputs and the fp64 oracle at the parity bars of test_model_gpu."""
import pytest
import torch

import golden_util as G
import hriemo_oracle as O
from hriemo import ops, pipeline

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOGIT_TOL, BETA_TOL, Z_TOL = 1e-2, 1e-4, 6e-2


def _tail_masks(B, T, gen, lo=1):
    lens = torch.randint(lo, T + 1, (B,), generator=gen)
    return torch.arange(T)[None, :] >= lens[:, None], lens


def test_mask_lengths_kernel():
    g = torch.Generator().manual_seed(1)
    m, _ = _tail_masks(300, 77, g, lo=0)
    m[5] = True                       # everything PAD
    m[6] = False
    m[7, 3] = True                    # a hole does not shorten the utterance
    got = ops.mask_lengths(m.to(DEV)).cpu()
    assert got.dtype == torch.int32 and torch.equal(got, pipeline.valid_lengths(m, 300, 77))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,cols,T_out", [(11, 40, 768, 23), (7, 33, 74, 33), (5, 9, 300, 14), (64, 500, 768, 311)])
def test_gather_utterances_and_masks(dtype, B, T, cols, T_out):
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, T, cols, generator=g).to(dtype).to(DEV)
    utt = torch.randperm(B, generator=g)[: max(1, B - 2)].to(torch.int32).to(DEV)
    out = ops.gather_utterances(x, utt, T_out)
    ld = (cols + 7) // 8 * 8
    assert out.shape == (utt.shape[0], T_out, ld) and out.dtype == torch.bfloat16
    keep = min(T, T_out)
    ref = x[utt.long(), :keep].to(torch.bfloat16)
    assert torch.equal(out[:, :keep, :cols].contiguous().view(torch.int16), ref.contiguous().view(torch.int16))
    assert bool((out[:, keep:] == 0).all()) and bool((out[:, :, cols:] == 0).all())
    m, _ = _tail_masks(B, T, g)
    mo = ops.gather_masks(m.to(DEV), utt, T_out)
    assert mo.dtype == torch.bool and torch.equal(mo[:, :keep].cpu(), m[utt.long().cpu(), :keep])
    assert bool(mo[:, keep:].all())


def test_scatter_rows_roundtrip():
    g = torch.Generator().manual_seed(2)
    src = torch.randn(37, 4, 96, generator=g).to(DEV)
    perm = torch.randperm(50, generator=g)[:37].to(torch.int32).to(DEV)
    dst = torch.zeros(50, 4, 96, device=DEV)
    ops.scatter_rows(src, perm, dst)
    assert torch.equal(dst[perm.long()], src)
    rest = torch.ones(50, dtype=torch.bool)
    rest[perm.long().cpu()] = False
    assert bool((dst[rest.to(DEV)] == 0).all())


def _check_against(lo, be, z, ref_lo, ref_be, ref_z):
    assert (lo - ref_lo).abs().max().item() <= LOGIT_TOL
    assert (be - ref_be).abs().max().item() <= BETA_TOL
    assert (z - ref_z).abs().max().item() <= Z_TOL
    assert torch.equal(be > 0.5, ref_be > 0.5)


@pytest.mark.parametrize("name", ["cfg2_iemocap_ragged", "ns_500x64_ragged", "cfg3_mosei_default"])
def test_bucketed_forwards_match_reference_golden(name):
    """Device-resident and host-staged bucketed forwards against the reference's outputs for the PADDED batch."""
    fx = G.load(name)
    model, ins = G.build_fusion(fx)
    model = model.to(DEV)
    dev_ins = [G.to_dev(x, DEV) for x in ins]
    B = ins[0].shape[0]
    lo, be, z = pipeline.forward_bucketed(model, *dev_ins, rows_per_slab=B * ins[0].shape[1], max_utts=max(1, B // 2))
    torch.cuda.synchronize()
    _check_against(lo.cpu(), be.cpu(), z.cpu(), fx["logits"], fx["beta"], fx["z"])
    if fx["kind"] != "mosei":   # the host path packs feature dims that are multiples of 8
        host = [x.clone().pin_memory() for x in ins]
        for _ in range(2):
            lo2, be2, z2 = pipeline.forward_from_host(model, *host, device=DEV, slab=max(1, B // 2), bucket=True)
        assert lo2.device.type == "cpu"
        _check_against(lo2, be2, z2, fx["logits"], fx["beta"], fx["z"])
        # both bucketed entries run the same kernels on the same trimmed slabs when the plans coincide
        assert (lo2 - lo.cpu()).abs().max().item() <= 2e-3


def test_bucketed_forward_ragged_batch_vs_padded_and_oracle():
    """A ragged batch of 96 utterances (lengths 1..T, one fully padded utterance, one mask with a hole):
    bucketed == padded within the bf16 path's own noise, NaN exactly where the padded forward is NaN,
    and an oracle spot check."""
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    torch.manual_seed(11)
    m = FusionWithEmotionDecoder(d_model=192, n_heads=2, beta_hidden=64, num_emotions=5).eval()
    sd = O.cast_state(m.state_dict(), torch.float64)
    m = m.to(DEV)
    g = torch.Generator().manual_seed(5)
    B, T_a, T_t = 96, 150, 40
    h_a, h_t = torch.randn(B, T_a, 192, generator=g), torch.randn(B, T_t, 192, generator=g)
    m_a, la = _tail_masks(B, T_a, g)
    m_t, lt = _tail_masks(B, T_t, g)
    m_a[3] = True                                   # fully padded audio: NaN outputs, like the reference
    m_a[4, 2] = True                                # a hole inside the valid region stays a hole
    m_a[5, :] = False                               # full length
    lens_a = pipeline.valid_lengths(m_a, B, T_a)
    order, buckets = pipeline.bucket_plan(lens_a, pipeline.valid_lengths(m_t, B, T_t), T_a, T_t, 16 * T_a, 32)
    st = pipeline.bucket_stats(lens_a, pipeline.valid_lengths(m_t, B, T_t), T_a, T_t, buckets)
    assert st["bucketed_rows"] < 0.75 * st["padded_rows"]
    ref = m(h_a.to(DEV), h_t.to(DEV), m_a.to(DEV), m_t.to(DEV))
    out = pipeline.forward_bucketed(m, h_a.to(DEV), h_t.to(DEV), m_a.to(DEV), m_t.to(DEV), rows_per_slab=16 * T_a, max_utts=32)
    host = pipeline.forward_from_host(m, h_a.pin_memory(), h_t.pin_memory(), m_a, m_t, device=DEV, slab=16, bucket=True)
    torch.cuda.synchronize()
    nan_rows = torch.isnan(ref[0]).any(dim=1).cpu()
    assert nan_rows.tolist() == [i == 3 for i in range(B)]
    for got in (out, host):
        lo, be, z = (t.cpu() for t in got)
        assert torch.equal(torch.isnan(lo).any(dim=1), nan_rows)
        ok = ~nan_rows
        assert (lo[ok] - ref[0].cpu()[ok]).abs().max().item() <= 2e-3       # same arithmetic, other tile shapes
        assert (be[ok] - ref[1].cpu()[ok]).abs().max().item() <= 1e-5
        assert (z[ok] - ref[2].cpu()[ok]).abs().max().item() <= 2e-2
    idx = torch.tensor([0, 4, 5, 17, 95])
    lo_o, be_o, z_o = O.fusion_with_emotion_decoder(sd, h_a[idx].double(), h_t[idx].double(), m_a[idx], m_t[idx], n_heads=2)
    _check_against(out[0].cpu()[idx], out[1].cpu()[idx], out[2].cpu()[idx], lo_o, be_o, z_o)


def test_bucketed_without_masks_falls_back_to_the_dense_plan():
    fx = G.load("cfg2_iemocap_nomask")
    model, ins = G.build_fusion(fx)
    model = model.to(DEV)
    host = [None if x is None else x.clone().pin_memory() for x in ins]
    lo, be, z = pipeline.forward_from_host(model, *host, device=DEV, slab=2, bucket=True)
    ref = model(*[G.to_dev(x, DEV) for x in ins])
    torch.cuda.synchronize()
    assert torch.equal(lo, ref[0].cpu()) and torch.equal(be, ref[1].cpu()) and torch.equal(z, ref[2].cpu())
    out = pipeline.forward_bucketed(model, *[G.to_dev(x, DEV) for x in ins])
    assert torch.equal(out[0], ref[0])


@pytest.mark.parametrize("name,dtype", [("cfg2_iemocap_ragged", torch.bfloat16), ("ns_500x64_ragged", torch.float32)])
def test_forward_from_shard_matches_reference_golden(tmp_path, name, dtype):
    """Packed shard -> pinned staging -> forward, against the reference's outputs for the
    same utterances; a bf16 shard holds exactly what the path's own input cast would have produced."""
    from hriemo import shards

    fx = G.load(name)
    model, (h_a, h_t, m_a, m_t) = G.build_fusion(fx)
    model = model.to(DEV)
    B = h_a.shape[0]
    path = str(tmp_path / "fx.hriemo")
    shards.write_shard(path, [(h_a[i], m_a[i], h_t[i], m_t[i]) for i in range(B)], dtype=dtype)
    with shards.Shard(path) as sh:
        for _ in range(2):
            lo, be, z = pipeline.forward_from_shard(model, sh, device=DEV, slab_rows=max(1, B // 2) * h_a.shape[1])
        back = sh.original_order()
    assert lo.device.type == "cpu" and lo.shape == fx["logits"].shape
    _check_against(lo, be, z, fx["logits"][back], fx["beta"][back], fx["z"][back])
    ref = model(*[G.to_dev(x, DEV) for x in (h_a, h_t, m_a, m_t)])
    torch.cuda.synchronize()
    assert (lo - ref[0].cpu()[back]).abs().max().item() <= 2e-3
