"""Drop-in scoring seams (reference: news_rec_utils/data_model_helper.py).

  get_final_attention_eval            data_model_helper.py:112-131
  get_cos_sim_scores                  data_model_helper.py:174-239
  get_final_second_attention_score    data_model_helper.py:416-443

Same positional signatures and return types (CPU tensors / numpy) as the reference; the work is
one table upload + one dense row transform (cached per table/weights) + one fused CUDA launch
instead of a DataLoader, a per-impression Python loop and a per-impression scipy call.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .data_utils import rank_group_preds, ranks_to_object_array
from .engine import cached_engine


def get_final_attention_eval(history_rev_index: np.ndarray, history_len_list: np.ndarray,
                             news_embeddings: torch.Tensor, model: torch.nn.Module, precision=None) -> torch.Tensor:
    eng = cached_engine(news_embeddings, model, None, precision=precision)
    return eng.user_vectors(history_rev_index, history_len_list).cpu()


def get_cos_sim_scores(history_rev_index: np.ndarray, history_len_list: np.ndarray, news_rev_index: np.ndarray,
                       impression_len_list: np.ndarray, news_embeddings: torch.Tensor, model: torch.nn.Module,
                       query_news_embeddings: Optional[torch.Tensor] = None, precision=None) -> torch.Tensor:
    assert len(history_len_list) == len(impression_len_list), "Number of rows should be consistent"
    assert int(np.sum(impression_len_list)) == len(news_rev_index), \
        "Number of impressions should match length of impression list"
    q = query_news_embeddings if isinstance(query_news_embeddings, torch.Tensor) else None
    eng = cached_engine(news_embeddings, model, q, precision=precision)
    _, scores, _ = eng.score(history_rev_index, history_len_list, news_rev_index, impression_len_list,
                             want_ranks=False)
    return scores.cpu()


def get_final_second_attention_score(history_rev_index: np.ndarray, history_len_list: np.ndarray,
                                     news_rev_index: np.ndarray, impression_len_list: np.ndarray,
                                     news_embeddings: torch.Tensor, history_bool, attention_model: torch.nn.Module,
                                     precision=None) -> dict:
    """{"scores": float32[sum C], "grouped_scores": object[I] of dense-rank arrays}.

    Like the reference this expects `history_*` to describe exactly the impressions selected by
    `history_bool` (scripts/eval.py loads the WITH_HISTORY subset, so it is all True)."""
    hb = np.asarray(history_bool, dtype=bool)
    imp_len = np.asarray(impression_len_list)
    cand_idx = np.asarray(news_rev_index)[np.repeat(hb, imp_len)]
    cand_len = imp_len[hb]
    eng = cached_engine(news_embeddings, attention_model, None, precision=precision)
    _, scores, ranks = eng.score(history_rev_index, history_len_list, cand_idx, cand_len, want_ranks=True)
    scores_np = scores.cpu().numpy()
    if hb.all():
        grouped = ranks_to_object_array(ranks.cpu().numpy(), imp_len)
    else:
        # the reference ranks the filtered scores against the UNFILTERED length list
        # (data_model_helper.py:442); reproduce that grouping
        grouped = rank_group_preds(scores_np, imp_len)
    return {"scores": scores_np, "grouped_scores": grouped}
