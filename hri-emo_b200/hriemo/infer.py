"""Post-path outputs of the reference's inference script, same files and layouts:

  <out_dir>/<split>_y_prob.npy       sigmoid(logits)  float32 [N, C]      mosei_eval_infer.py:248, :263-266
  <out_dir>/<split>_y_true.npy       labels           as given [N, ...]   :249, :267
  <out_dir>/<split>_beta_mean.npy    beta reduced to one scalar per utterance [N]   :252-259, :271-274
  <out_dir>/<split>_attentions.pt    {"encoder": [per batch [per layer {name: ndarray [B,Tq,Tk]}]],
                                      "decoder": [per batch [per layer ndarray [B,N_e,L]]]}     :211-233, :277-284
  <out_dir>/<split>_y_pred.npy       (extra, only with thresholds) probs >= thresholds[c], the decision of
                                     scripts/analysis/mosei_summary_metrics.py:51, uint8 [N, C]

sigmoid and the threshold compare run in the hriemo_emotion_outputs kernel on the logits the forward left in
HBM; only the results cross PCIe.  `batches` is what the reference's DataLoader yields:
(h_a, m_a, h_t, m_t, y) per batch (collate_seq_batch, mosei_eval_infer.py:128-147)."""
from __future__ import annotations

import os
from typing import Iterable, Optional, Sequence

import numpy as np
import torch

from . import ops


@torch.no_grad()
def run_split(model, batches: Iterable, device, out_dir: str, split_name: str, dump_beta: bool = False,
              dump_attn: bool = False, attn_max_samples: int = 64, thresholds: Optional[Sequence[float]] = None) -> dict:
    dev = torch.device(device)
    model.eval()
    thr = None if thresholds is None else torch.as_tensor(list(thresholds), dtype=torch.float32, device=dev)
    probs, preds, labels, betas = [], [], [], []
    collected = {"encoder": [], "decoder": []}
    seen = 0
    for h_a, m_a, h_t, m_t, y in batches:
        h_a, h_t = h_a.to(dev, non_blocking=True), h_t.to(dev, non_blocking=True)
        m_a = None if m_a is None else m_a.to(dev, non_blocking=True)
        m_t = None if m_t is None else m_t.to(dev, non_blocking=True)
        if dump_attn and seen < attn_max_samples:
            logits, beta, _, pack = model(h_a, h_t, m_a, m_t, return_attention=True)
            collected["encoder"].append([{k: v.float().cpu().numpy() for k, v in layer.items()} for layer in pack["encoder"]])
            collected["decoder"].append([layer.float().cpu().numpy() for layer in pack["decoder"]])
            seen += h_a.shape[0]
        else:
            logits, beta, _ = model(h_a, h_t, m_a, m_t)
        p, dec = ops.emotion_outputs(logits, thr)
        probs.append(p.cpu().numpy())
        if thr is not None:
            preds.append(dec.to(torch.uint8).cpu().numpy())
        labels.append(y.numpy() if isinstance(y, torch.Tensor) else np.asarray(y))
        if dump_beta and beta is not None:
            b = beta.float().cpu()
            while b.ndim > 1:
                b = b.mean(dim=-1)
            betas.append(b.numpy())
    os.makedirs(out_dir, exist_ok=True)
    out = {"y_prob": np.concatenate(probs, axis=0), "y_true": np.concatenate(labels, axis=0)}
    np.save(os.path.join(out_dir, f"{split_name}_y_prob.npy"), out["y_prob"])
    np.save(os.path.join(out_dir, f"{split_name}_y_true.npy"), out["y_true"])
    if preds:
        out["y_pred"] = np.concatenate(preds, axis=0)
        np.save(os.path.join(out_dir, f"{split_name}_y_pred.npy"), out["y_pred"])
    if dump_beta and betas:
        out["beta_mean"] = np.concatenate(betas, axis=0)
        np.save(os.path.join(out_dir, f"{split_name}_beta_mean.npy"), out["beta_mean"])
    if dump_attn:
        out["attentions"] = collected
        out["attn_samples"] = seen
        torch.save(collected, os.path.join(out_dir, f"{split_name}_attentions.pt"))
    return out
