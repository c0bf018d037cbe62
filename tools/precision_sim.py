"""Simulate the bf16 storage policy of the CUDA path on CPU (torch fp32 math with
explicit bf16 roundings) and report logits/beta error vs the fp32 oracle.
Design aid only; not used by the product path."""
import sys, math, torch
sys.path.insert(0, "oracle")
import hriemo_oracle as O

def r(x):  # round-to-bf16 storage
    return x.to(torch.bfloat16).to(torch.float32)

class Pol:
    preln_bf16 = True     # residual+bias+acc stored as bf16 before LN
    act_bf16 = True       # LN outputs stored bf16
    dec_fp32_stream = False

def lin(x, w, b):  # bf16 operands, fp32 accumulate
    return r(x) @ r(w).t() + b

def mha(sd, p, xq, xkv, H, pad):
    w, b = sd[p+"in_proj_weight"], sd[p+"in_proj_bias"]; d = w.shape[1]; dh = d//H
    B,Tq,_ = xq.shape; Tk = xkv.shape[1]
    q = r(lin(xq, w[:d], b[:d])); k = r(lin(xkv, w[d:2*d], b[d:2*d])); v = r(lin(xkv, w[2*d:], b[2*d:]))
    q = q.view(B,Tq,H,dh).transpose(1,2); k = k.view(B,Tk,H,dh).transpose(1,2); v = v.view(B,Tk,H,dh).transpose(1,2)
    s = q @ k.transpose(-1,-2) / math.sqrt(dh)
    if pad is not None: s = s.masked_fill(pad[:,None,None,:], float("-inf"))
    m = s.max(-1, keepdim=True).values
    e = torch.exp(s-m); l = e.sum(-1, keepdim=True)
    o = (r(e) @ v) / l
    o = r(o.transpose(1,2).reshape(B,Tq,d))
    return lin(o, sd[p+"out_proj.weight"], sd[p+"out_proj.bias"])

def ln(sd, p, x, pol):
    if pol.preln_bf16: x = r(x)
    y = O.layer_norm(x, sd[p+"weight"], sd[p+"bias"])
    return r(y) if pol.act_bf16 else y

def ffn(sd, p, x):
    h = r(torch.relu(lin(x, sd[p+"0.weight"], sd[p+"0.bias"])))
    return lin(h, sd[p+"2.weight"], sd[p+"2.bias"])

def forward(sd, a, t, ma, mt, H, pol):
    a = r(a); t = r(t)
    i = 0
    while f"cross_modal.layers.{i}.norm_a1.weight" in sd:
        p = f"cross_modal.layers.{i}."
        a_s = ln(sd, p+"self_norm_a.", a + mha(sd, p+"self_attn_a.", a, a, H, ma), pol)
        t_s = ln(sd, p+"self_norm_t.", t + mha(sd, p+"self_attn_t.", t, t, H, mt), pol)
        a1 = ln(sd, p+"norm_a1.", a_s + mha(sd, p+"attn_a2t.", a_s, t_s, H, mt), pol)
        a = ln(sd, p+"norm_a2.", a1 + ffn(sd, p+"ffn_a.", a1), pol)
        t1 = ln(sd, p+"norm_t1.", t_s + mha(sd, p+"attn_t2a.", t_s, a_s, H, ma), pol)
        t = ln(sd, p+"norm_t2.", t1 + ffn(sd, p+"ffn_t.", t1), pol)
        i += 1
    # gate: fp32 math on bf16-stored streams, fp32 MLP
    p = "beta_gate."
    a_n = O.layer_norm(a, sd[p+"norm_a.weight"], sd[p+"norm_a.bias"]); t_n = O.layer_norm(t, sd[p+"norm_t.weight"], sd[p+"norm_t.bias"])
    g = O._gate_input(O.masked_mean(a_n, ma), O.masked_mean(t_n, mt))
    w = torch.sigmoid(O._gate_mlp(sd, p, g)); beta = w.mean(-1, keepdim=True)
    L = t.shape[1]
    hf = r(w[:,None]*a_n[:, :L] + (1-w[:,None])*t_n)
    fm = O.build_fused_mask(ma, mt, L)
    p = "emotion_decoder."
    z = sd[p+"emotion_queries"].unsqueeze(0).expand(a.shape[0], -1, -1)
    class P2(Pol): pass
    pd = P2(); pd.preln_bf16 = pol.preln_bf16 and not pol.dec_fp32_stream; pd.act_bf16 = pol.act_bf16 and not pol.dec_fp32_stream
    i = 0
    while f"{p}layers.{i}.norm1.weight" in sd:
        q = f"{p}layers.{i}."
        z = ln(sd, q+"norm1.", z + mha(sd, q+"self_attn.", z, z, H, None), pd)
        z = ln(sd, q+"norm2.", z + mha(sd, q+"cross_attn.", z, hf, H, fm), pd)
        f = lin(r(torch.relu(lin(z, sd[q+"linear1.weight"], sd[q+"linear1.bias"]))), sd[q+"linear2.weight"], sd[q+"linear2.bias"])
        z = ln(sd, q+"norm3.", z + f, pd)
        i += 1
    logits = (z @ sd[p+"out_proj.weight"].t() + sd[p+"out_proj.bias"]).squeeze(-1)
    return logits, beta, z

if __name__ == "__main__":
    torch.manual_seed(1234)
    sys.path.insert(0, "/root/reference")
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder
    m = FusionWithEmotionDecoder().eval(); sd = {k: v.detach() for k, v in m.state_dict().items()}
    B, Ta, Tt = 24, 300, 50
    g = torch.Generator().manual_seed(1234)
    a = torch.randn(B,Ta,768,generator=g); t = torch.randn(B,Tt,768,generator=g)
    ma = O.ragged_masks(B,Ta,g); mt = O.ragged_masks(B,Tt,g)
    with torch.no_grad():
        for masks in [(None,None),(ma,mt)]:
            lo, be, z = m(a, t, *masks)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                lo_ac, be_ac, _ = m(a, t, *masks)
            print("torch autocast bf16: logits", (lo_ac.float()-lo).abs().max().item(), "beta", (be_ac.float()-be).abs().max().item())
            for name, kw in [("all-bf16", {}), ("preln fp32", dict(preln_bf16=False)), ("dec fp32 stream", dict(dec_fp32_stream=True)),
                             ("preln fp32 + dec fp32", dict(preln_bf16=False, dec_fp32_stream=True))]:
                pol = Pol()
                for k, v in kw.items(): setattr(pol, k, v)
                lo2, be2, z2 = forward(sd, a, t, *masks, 8, pol)
                print(f"{name:28s} logits {(lo2-lo).abs().max().item():.2e}  beta {(be2-be).abs().max().item():.2e}  z {(z2-z).abs().max().item():.2e}"
                      f"  thr {((lo2>0)==(lo>0)).float().mean().item():.4f} argmax {(lo2.argmax(-1)==lo.argmax(-1)).float().mean().item():.3f}"
                      f"  beta>0.5 {((be2>0.5)==(be>0.5)).float().mean().item():.3f} beta-argmax {int(be2.argmax())==int(be.argmax())}")
