"""Drop-in `score` (reference: news_rec_utils/evaluation.py:57-98) computed on the GPU.

The reference maps `score_row` over a 4-process pool (sklearn `roc_auc_score` + numpy argsorts,
~2.2 ms per impression); here one kernel launch (`nrb_mind_metrics`) produces the per-impression
AUC / MRR / nDCG@5 / nDCG@10 and their sums.  Same return dict.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _lib, ops
from .synthetic import csr_offsets


def flatten_groups(groups: Sequence, dtype) -> tuple[np.ndarray, np.ndarray]:
    lens = np.fromiter((len(g) for g in groups), dtype=np.int64, count=len(groups))
    flat = np.concatenate([np.asarray(g) for g in groups]) if len(groups) else np.zeros(0)
    return np.ascontiguousarray(np.nan_to_num(flat, nan=0.0).astype(dtype)), lens


def score_device(ranks: torch.Tensor, labels: torch.Tensor, offsets: torch.Tensor, strict: bool = True) -> dict:
    """Device tensors in (ranks int32, labels int8, offsets int64) -> the reference's score dict."""
    per, sums = ops.mind_metrics(ranks, labels, offsets, want_per_impression=False)
    s = sums.cpu().numpy()
    n_imp = offsets.numel() - 1
    if strict and int(s[4]) != n_imp:
        # sklearn's roc_auc_score raises for single-class impressions (evaluation.py:49)
        raise ValueError(f"{n_imp - int(s[4])} impressions have only one class (or NaN ranks): AUC is undefined")
    n = max(float(s[4]), 1.0)
    return {"auc": float(s[0] / n), "mrr": float(s[1] / n), "ndcg5": float(s[2] / n), "ndcg10": float(s[3] / n),
            "num_samples": int(n_imp)}


def score(preds_input, labels_input, imp_ids: Sequence[str] = (), debug_dir=None) -> dict:
    """preds_input: per-impression dense ranks (`grouped_scores`), labels_input: per-impression labels."""
    dev = _lib.require_device()
    ranks, lens = flatten_groups(preds_input, np.int32)
    labels, lens2 = flatten_groups(labels_input, np.int8)
    assert np.array_equal(lens, lens2), "preds and labels must have the same group sizes"
    with torch.cuda.device(dev):
        return score_device(torch.from_numpy(ranks).to(dev), torch.from_numpy(labels).to(dev),
                            torch.from_numpy(csr_offsets(lens)).to(dev))
