import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("hri-emo_b200", "oracle", "tests"): sys.path.insert(0, os.path.join(ROOT, p))
from hriemo import ops, pipeline
from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder
DEV = "cuda"
def tail(B, T, gen, lo=1):
    lens = torch.randint(lo, T + 1, (B,), generator=gen)
    return torch.arange(T)[None, :] >= lens[:, None], lens
torch.manual_seed(11)
m = FusionWithEmotionDecoder(d_model=192, n_heads=2, beta_hidden=64, num_emotions=5).eval().to(DEV)
g = torch.Generator().manual_seed(5)
B, T_a, T_t = 96, 150, 40
h_a, h_t = torch.randn(B, T_a, 192, generator=g), torch.randn(B, T_t, 192, generator=g)
m_a, la = tail(B, T_a, g); m_t, lt = tail(B, T_t, g)
variant = sys.argv[1] if len(sys.argv) > 1 else "full"
if variant == "full":
    m_a[3] = True; m_a[4, 2] = True; m_a[5, :] = False
la = pipeline.valid_lengths(m_a, B, T_a); lt = pipeline.valid_lengths(m_t, B, T_t)
order, buckets = pipeline.bucket_plan(la, lt, T_a, T_t, 16 * T_a, 32)
ref = m(h_a.to(DEV), h_t.to(DEV), m_a.to(DEV), m_t.to(DEV))
out = pipeline.forward_bucketed(m, h_a.to(DEV), h_t.to(DEV), m_a.to(DEV), m_t.to(DEV), rows_per_slab=16 * T_a, max_utts=32)
err = (out[0] - ref[0]).abs().amax(dim=1).cpu()
pos = {int(u): k for k, u in enumerate(order.tolist())}
for bi, bk in enumerate(buckets):
    print(f"bucket {bi}: n={bk.end-bk.start} T_a={bk.T_a} T_t={bk.T_t}")
    for k in range(bk.start, bk.end):
        u = int(order[k])
        if not (err[u] <= 2e-3):
            print(f"   utt {u:3d} pos {k-bk.start:2d} la={int(la[u]):3d} lt={int(lt[u]):2d} err={float(err[u]):.4f}")
# direct: run the padded model on each trimmed bucket with torch indexing (no gather kernels) and compare
for bi, bk in enumerate(buckets):
    idx = order[bk.start:bk.end].long()
    r = m(h_a[idx, :bk.T_a].contiguous().to(DEV), h_t[idx, :bk.T_t].contiguous().to(DEV), m_a[idx, :bk.T_a].contiguous().to(DEV), m_t[idx, :bk.T_t].contiguous().to(DEV))
    e2 = (r[0].cpu() - ref[0].cpu()[idx]).abs().amax(dim=1)
    e3 = (r[0].cpu() - out[0].cpu()[idx]).abs().amax(dim=1)
    print(f"bucket {bi}: torch-indexed trimmed vs padded max {float(e2[~torch.isnan(e2)].max()):.4f}; vs gathered-bucketed max {float(e3[~torch.isnan(e3)].max()):.6f}")
