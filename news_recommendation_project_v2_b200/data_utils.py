"""Drop-in data helpers on the hot path (reference: news_rec_utils/data_utils.py).

  group_items                      data_utils.py:400-411
  rank_group_preds                 data_utils.py:414-415   -> nrb_dense_rank
  pad_to_maxlen                    data_utils.py:723-750
  FinalAttentionEvalDataset        data_utils.py:501-509
  final_attention_eval_collate_fn  data_utils.py:784-791   -> nrb_gather_collate
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np
import torch

from . import _lib, ops
from .engine import resident_table
from .synthetic import csr_offsets


def group_items(items: np.ndarray, imp_counts: np.ndarray, func: Callable[[np.ndarray], np.ndarray] = lambda x: x):
    """Split a flat array by counts.  Always returns a 1-D object array (the reference's
    `np.array(list, dtype=object)` silently becomes 2-D when all counts are equal -- quirk a7)."""
    off = csr_offsets(np.asarray(imp_counts))
    out = np.empty(len(imp_counts), dtype=object)
    for i in range(len(imp_counts)):
        out[i] = func(items[off[i]:off[i + 1]])
    return out


def ranks_to_object_array(ranks: np.ndarray, imp_counts: np.ndarray) -> np.ndarray:
    """int32 flat dense ranks (0 = NaN group) -> the reference's object array of per-impression arrays."""
    r = ranks.astype(np.float32)
    if (ranks == 0).any():
        r[ranks == 0] = np.nan
    return group_items(r, imp_counts)


def rank_group_preds(pred_scores: np.ndarray, imp_counts: np.ndarray) -> np.ndarray:
    """scipy.stats.rankdata(-x, method='dense') per impression, computed by nrb_dense_rank."""
    dev = _lib.require_device()
    scores = np.asarray(pred_scores)
    # scipy ranks the array in its own dtype (data_utils.py:415): float64 stays float64, everything else is fp32
    scores = np.ascontiguousarray(scores, dtype=np.float64 if scores.dtype == np.float64 else np.float32)
    counts = np.asarray(imp_counts)
    off = csr_offsets(counts)
    assert off[-1] == scores.shape[0], "sum(imp_counts) must equal len(pred_scores)"
    with torch.cuda.device(dev):
        ranks = ops.dense_rank(torch.from_numpy(scores).to(dev), torch.from_numpy(off).to(dev)).cpu().numpy()
    return ranks_to_object_array(ranks, counts)


def pad_to_maxlen(grouped_items: Sequence[np.ndarray]) -> dict[str, np.ndarray]:
    lens = np.fromiter((len(g) for g in grouped_items), dtype=np.int64, count=len(grouped_items))
    mx = int(lens.max())
    idx = np.zeros((len(lens), mx), dtype=np.int32)
    msk = (np.arange(mx)[None, :] < lens[:, None]).astype(np.int32)
    for i, g in enumerate(grouped_items):
        idx[i, :lens[i]] = g
    return {"indices": idx, "attention_mask": msk}


class FinalAttentionEvalDataset(torch.utils.data.Dataset):
    def __init__(self, history_rev_index: np.ndarray, history_len_list: np.ndarray):
        self.group_history = group_items(history_rev_index, history_len_list)

    def __len__(self):
        return len(self.group_history)

    def __getitem__(self, idx):
        return self.group_history[idx]


def final_attention_eval_collate_fn(input: Sequence[np.ndarray], news_embeddings: torch.Tensor, device=None):
    """(table[indices] * mask[..., None], mask) for a batch of ragged history index lists.

    The gather runs on the GPU against the device-resident table (uploaded once).  Returns CPU
    tensors like the reference unless `device` is given (or the table already lives on the GPU)."""
    dev = _lib.require_device(news_embeddings.device if news_embeddings.is_cuda else None)
    lens = np.fromiter((len(g) for g in input), dtype=np.int64, count=len(input))
    off = csr_offsets(lens)
    flat = np.concatenate([np.asarray(g, dtype=np.int32) for g in input]) if off[-1] > 0 else np.zeros(1, np.int32)
    tdt = news_embeddings.dtype if news_embeddings.dtype in (torch.float32, torch.bfloat16) else torch.float32
    with torch.cuda.device(dev):
        table = resident_table(news_embeddings, tdt, dev)
        emb, mask = ops.gather_collate(table, torch.from_numpy(np.ascontiguousarray(flat, dtype=np.int32)).to(dev),
                                       torch.from_numpy(off).to(dev), int(lens.max()))
    if device is not None or news_embeddings.is_cuda:
        return emb, mask
    return emb.cpu(), mask.cpu()


def split_impressions_and_history(impressions: Sequence[str], history: Sequence) -> dict:
    """Behaviour strings -> table row ids + CSR index arrays (data_utils.py:168-232), built by the native
    host routine `nrb_csr_build` instead of a per-row Python loop.  Same keys / dtypes as the reference:
    news_list, impression_rev_ind_array int32[2, sum C], impression_len_list int32[I],
    history_rev_ind_array int32[2, sum H], history_len_list int32[rows with history], labels object[I]."""
    import ctypes as C

    assert len(impressions) > 0, "No Impressions given"
    lib = _lib.load()
    n = len(impressions)
    imp_buf = "\n".join(impressions).encode()
    hist_buf = "\n".join(h if isinstance(h, str) else "" for h in history).encode()
    handle = lib.nrb_csr_build(imp_buf, len(imp_buf), hist_buf, len(hist_buf), n)
    if not handle:
        raise _lib.NrbError("nrb_csr_build failed: " + lib.nrb_last_error().decode(errors="replace"))
    try:
        sizes = (C.c_int64 * 5)()
        _lib.check(lib.nrb_csr_sizes(handle, sizes), "nrb_csr_sizes")
        n_news, n_h, n_hrows, n_c, has_labels = (int(v) for v in sizes)
        hist = np.empty((2, n_h), dtype=np.int32)
        cand = np.empty((2, n_c), dtype=np.int32)
        hist_len = np.empty(n_hrows, dtype=np.int32)
        cand_len = np.empty(n, dtype=np.int32)
        labels_flat = np.empty(n_c if has_labels else 0, dtype=np.int8)
        p = lambda a: a.ctypes.data if a.size else None
        _lib.check(lib.nrb_csr_export(handle, p(hist[0]), p(hist[1]), p(hist_len), p(cand[0]), p(cand[1]),
                                      p(cand_len), p(labels_flat)), "nrb_csr_export")
        need = lib.nrb_csr_news_ids(handle, None, 0)
        buf = C.create_string_buffer(int(need) + 1)
        lib.nrb_csr_news_ids(handle, buf, need)
        news_list = np.array(buf.raw[:need].decode().split("\n")[:-1])
    finally:
        lib.nrb_csr_free(handle)
    labels = np.empty(n if has_labels else 0, dtype=object)
    if has_labels:
        off = csr_offsets(cand_len)
        for i in range(n):
            labels[i] = tuple(int(v) for v in labels_flat[off[i]:off[i + 1]])
    return {"news_list": news_list, "impression_rev_ind_array": cand, "impression_len_list": cand_len,
            "history_rev_ind_array": hist, "history_len_list": hist_len, "labels": labels}
