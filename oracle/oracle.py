"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch CPU restatement of the reference's embedding-and-scoring hot
path.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module,
and only as the checker / the CPU baseline -- never on the product path.

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF run in the build container: `oracle/make_golden.py` imports
the unmodified reference through `oracle/ref_harness.py`, runs it on seeded
inputs and commits the results under `tests/golden/`;
`tests/test_oracle_golden.py` checks every function below against them (and
against the live reference when `/root/reference` is present).

All arithmetic lives in third-party libraries that the reference does not pin
(pyproject.toml:8-23 lists bare names): torch (LayerNorm / Linear / SDPA /
F.normalize / F.cosine_similarity / F.gelu), scipy.stats.rankdata,
sklearn.metrics.roc_auc_score, numpy argsort.  Versions in this image:
torch 2.11.0, scipy 1.18.1, scikit-learn 1.9.0, numpy 2.3.5.  The functions
below restate the published algorithms of those calls with plain tensor
arithmetic so the oracle also runs in float64.

Every function takes plain arrays / a state_dict (the reference's key names,
SURVEY.md section 8a rows a1, a9) -- there are no nn.Modules here.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch

# --------------------------------------------------------------------------
# Stage A -- latent-attention pooling (reference: latent_attention.py:6-171)
# --------------------------------------------------------------------------

LN_EPS = 1e-5  # torch.nn.LayerNorm default, latent_attention.py:10-12


def _layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    # latent_attention.py:16,19 (PreNorm.forward): biased variance, eps inside sqrt
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def _gelu_erf(x: torch.Tensor) -> torch.Tensor:
    # latent_attention.py:27 -- F.gelu default = exact erf form
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def latent_block(
    sd: dict,
    x: torch.Tensor,
    heads: int = 8,
    dim_head: int = 512,
    dtype: torch.dtype = torch.float64,
) -> torch.Tensor:
    """Per-token output of the two cross_attend_blocks, [B,S,d] -> [B,S,d].

    Restates latent_attention.py:154-163: hiddens = attn(LN(h), LN_ctx(latents)) + h;
    hiddens = FF(LN(hiddens)) + hiddens.  K/V are projected ONCE from the latents
    (they are input independent; the reference recomputes them per batch row,
    latent_attention.py:161,67 -- same values).
    """
    g = lambda k: sd[k].detach().to(dtype)
    x = x.to(dtype)
    lat = g("latents")  # [L,d]
    inner = heads * dim_head
    # --- block 0: PreNorm(Attention) -- latent_attention.py:15-21, 63-74
    xn = _layer_norm(x, g("cross_attend_blocks.0.norm.weight"), g("cross_attend_blocks.0.norm.bias"))
    cn = _layer_norm(lat, g("cross_attend_blocks.0.norm_context.weight"), g("cross_attend_blocks.0.norm_context.bias"))
    q = xn @ g("cross_attend_blocks.0.fn.to_q.weight").T  # [B,S,I]
    kv = cn @ g("cross_attend_blocks.0.fn.to_kv.weight").T  # [L,2I]
    k, v = kv[:, :inner], kv[:, inner:]  # chunk(2): k = first I columns (:67)
    B, S, _ = x.shape
    L = lat.shape[0]
    qh = q.reshape(B, S, heads, dim_head)
    kh = k.reshape(L, heads, dim_head)
    vh = v.reshape(L, heads, dim_head)
    # SDPA default scale = dim_head ** -0.5 (:69-72); no mask; softmax over latents
    logits = torch.einsum("bshd,lhd->bshl", qh, kh) / math.sqrt(dim_head)
    logits = logits - logits.max(dim=-1, keepdim=True).values
    p = torch.exp(logits)
    p = p / p.sum(dim=-1, keepdim=True)
    o = torch.einsum("bshl,lhd->bshd", p, vh).reshape(B, S, inner)
    h1 = o @ g("cross_attend_blocks.0.fn.to_out.weight").T + x  # (:74, :162)
    # --- block 1: PreNorm(FeedForward) -- latent_attention.py:30-40, :163
    hn = _layer_norm(h1, g("cross_attend_blocks.1.norm.weight"), g("cross_attend_blocks.1.norm.bias"))
    f = hn @ g("cross_attend_blocks.1.fn.net.0.weight").T + g("cross_attend_blocks.1.fn.net.0.bias")
    half = f.shape[-1] // 2
    a, gate = f[..., :half], f[..., half:]  # GEGLU chunk(2) (:25-27)
    y = (a * _gelu_erf(gate)) @ g("cross_attend_blocks.1.fn.net.2.weight").T + g("cross_attend_blocks.1.fn.net.2.bias")
    return y + h1


def latent_pool(
    sd: dict,
    x: torch.Tensor,
    attention_mask: Optional[torch.Tensor],
    heads: int = 8,
    dim_head: int = 512,
    dtype: torch.dtype = torch.float64,
) -> torch.Tensor:
    """LatentAttentionModel.forward (latent_attention.py:134-171).

    mask None -> un-pooled [B,S,d]; else masked mean over S then L2 normalise
    (F.normalize p=2 eps=1e-12, :165-170).  An all-zero mask row gives NaN (0/0).
    """
    h = latent_block(sd, x, heads, dim_head, dtype)
    if attention_mask is None:
        return h
    m = attention_mask.to(dtype)
    s = (h * m.unsqueeze(-1)).sum(dim=1)
    d = m.sum(dim=1, keepdim=True)
    pooled = s / d
    nrm = pooled.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    return pooled / nrm


# --------------------------------------------------------------------------
# Stage B -- history gather + FinalAttention user encoder
# --------------------------------------------------------------------------


def group_items(items: np.ndarray, counts: np.ndarray) -> list:
    """data_utils.py:400-411 -- split a flat array by counts (exclusive cumsum)."""
    off = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)])
    return [items[off[i] : off[i + 1]] for i in range(len(counts))]


def pad_to_maxlen(groups: Sequence[np.ndarray]) -> tuple[np.ndarray, np.ndarray]:
    """data_utils.py:723-750 -- right-pad index lists with 0; int32 mask."""
    n = len(groups)
    mx = max(len(g) for g in groups)
    idx = np.zeros((n, mx), dtype=np.int32)
    msk = np.zeros((n, mx), dtype=np.int32)
    for i, g in enumerate(groups):
        idx[i, : len(g)] = g
        msk[i, : len(g)] = 1
    return idx, msk


def final_attention_eval_collate(groups: Sequence[np.ndarray], table: torch.Tensor):
    """data_utils.py:784-791 -- table[indices] * mask[..., None], mask."""
    idx, msk = pad_to_maxlen(groups)
    idx_t = torch.from_numpy(idx).long()
    msk_t = torch.from_numpy(msk)
    return table[idx_t] * msk_t.unsqueeze(-1), msk_t


def final_attention_rows(sd: dict, rows: torch.Tensor, dtype=torch.float64):
    """Per-row part of FinalAttention.forward (modeling_utils.py:218-222).

    Returns (x, logit): x = linear3(relu(linear2(relu(linear1(e))))),
    logit = linear5(relu(linear4(x))).  Dropout is inactive in eval.
    """
    g = lambda k: sd[k].detach().to(dtype)
    e = rows.to(dtype)
    x = torch.relu(e @ g("linear1.weight").T + g("linear1.bias"))
    x = torch.relu(x @ g("linear2.weight").T + g("linear2.bias"))
    x = x @ g("linear3.weight").T + g("linear3.bias")
    w = torch.relu(x @ g("linear4.weight").T + g("linear4.bias"))
    w = w @ g("linear5.weight").T
    return x, w


def final_attention(sd: dict, emb: torch.Tensor, mask: torch.Tensor, dtype=torch.float64) -> torch.Tensor:
    """FinalAttention.forward (modeling_utils.py:195-228) on a padded batch.

    weights = exp(logit) * mask (NO max subtraction, :224);
    weights /= sum_s(weights) + 1e-10 (:225); out = sum_s(x * weights) (:228).
    """
    x, w = final_attention_rows(sd, emb, dtype)
    w = torch.exp(w) * mask.to(dtype).unsqueeze(-1)
    w = w / (w.sum(dim=1, keepdim=True) + 1e-10)
    return (x * w).sum(dim=1)


def new_attention(sd: dict, emb: torch.Tensor, mask: torch.Tensor, num_layers: int = 1, dtype=torch.float64) -> torch.Tensor:
    """NewAttention.forward (attention.py:251-272) as it actually computes: every MyLayer returns
    g_mlp_layernorm(hidden_states) (attention.py:193, LayerNorm eps 1e-12; the attention / gated-MLP results
    are discarded), then per-dimension exp weights from linear1, masked normalise (+1e-10), weighted sum."""
    g = lambda k: sd[k].detach().to(dtype)
    x = emb.to(dtype)
    for i in range(num_layers):
        w, b = g(f"encoder.layer.{i}.g_mlp_layernorm.weight"), g(f"encoder.layer.{i}.g_mlp_layernorm.bias")
        mu = x.mean(dim=-1, keepdim=True)
        var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
        x = (x - mu) / torch.sqrt(var + 1e-12) * w + b
    wts = torch.exp(x @ g("linear1.weight").T + g("linear1.bias")) * mask.to(dtype).unsqueeze(-1)
    wts = wts / (wts.sum(dim=1, keepdim=True) + 1e-10)
    return (x * wts).sum(dim=1)


def user_vectors(sd: dict, table: torch.Tensor, hist_idx: np.ndarray, hist_len: np.ndarray,
                 batch: int = 64, dtype=torch.float64) -> torch.Tensor:
    """get_final_attention_eval (data_model_helper.py:112-131): batches of padded
    gathers through the user encoder, concatenated.  float32[I,d] in the reference."""
    groups = group_items(hist_idx, hist_len)
    out = []
    for s in range(0, len(groups), batch):
        emb, msk = final_attention_eval_collate(groups[s : s + batch], table)
        out.append(final_attention(sd, emb, msk, dtype))
    return torch.cat(out)


def latent_user_vectors(sd: dict, table: torch.Tensor, hist_idx: np.ndarray, hist_len: np.ndarray,
                        batch: int = 16, dtype=torch.float64, heads: int = 8, dim_head: int = 512) -> torch.Tensor:
    """get_final_attention_eval (data_model_helper.py:112-131) with LatentAttentionModel as the user encoder
    (the slot components.py:504, 675 show): padded history gather -> latent_attention.py:134-171 per batch."""
    groups = group_items(hist_idx, hist_len)
    out = []
    for s in range(0, len(groups), batch):
        emb, msk = final_attention_eval_collate(groups[s : s + batch], table)
        out.append(latent_pool(sd, emb, msk, heads=heads, dim_head=dim_head, dtype=dtype))
    return torch.cat(out)


def latent_second_attention_score(sd: dict, table: torch.Tensor, hist_idx, hist_len, cand_idx, cand_len,
                                  dtype=torch.float64, batch: int = 16, heads: int = 8, dim_head: int = 512):
    """get_final_second_attention_score (data_model_helper.py:416-443) with the latent-attention user encoder
    (BASELINE configs[4])."""
    u = latent_user_vectors(sd, table, hist_idx, hist_len, batch=batch, dtype=dtype, heads=heads, dim_head=dim_head)
    s = cosine_scores(u, table, cand_idx, cand_len, dtype=dtype)
    s_np = s.detach().numpy()
    return {"user": u, "scores": s_np, "grouped_scores": rank_group_preds(s_np, cand_len)}


# --------------------------------------------------------------------------
# Stage C -- cosine scoring + per-impression dense rank
# --------------------------------------------------------------------------


def cosine_scores(user: torch.Tensor, table: torch.Tensor, cand_idx: np.ndarray, cand_len: np.ndarray,
                  dtype=torch.float64, eps: float = 1e-8) -> torch.Tensor:
    """get_cos_sim_scores' loop (data_model_helper.py:200-230).

    F.cosine_similarity(u[d], c[C,d]) in torch 2.x normalises FIRST:
    sum_k (u_k / max(|u|,eps)) * (c_k / max(|c|,eps)), eps = 1e-8.
    """
    off = np.concatenate([[0], np.cumsum(cand_len, dtype=np.int64)])
    res = []
    t = table.to(dtype)
    for i in range(len(cand_len)):
        u = user[i].to(dtype)
        c = t[torch.from_numpy(np.asarray(cand_idx[off[i] : off[i + 1]])).long()]
        un = u / u.norm().clamp_min(eps)
        cn = c / c.norm(dim=-1, keepdim=True).clamp_min(eps)
        res.append((cn * un).sum(dim=-1))
    return torch.cat(res) if res else torch.zeros(0, dtype=dtype)


def dense_rank_desc(scores: np.ndarray) -> np.ndarray:
    """scipy.stats.rankdata(-x, method="dense") restated (data_utils.py:414-415).

    Rank 1 = highest score; equal scores share a rank; ranks are consecutive.
    -0.0 == 0.0.  Any NaN -> the whole group is NaN (scipy nan_policy
    'propagate').  Returns float64 (integer valued, or NaN)."""
    x = np.asarray(scores, dtype=np.float64)
    if x.size == 0:
        return np.zeros(0, dtype=np.float64)
    if np.isnan(x).any():
        return np.full(x.shape, np.nan)
    uniq = np.unique(x)  # ascending, exact equality (0.0 and -0.0 merge)
    pos = np.searchsorted(uniq, x)
    return (len(uniq) - pos).astype(np.float64)


def rank_group_preds(scores: np.ndarray, counts: np.ndarray) -> list:
    return [dense_rank_desc(g) for g in group_items(np.asarray(scores), counts)]


def final_second_attention_score(sd: dict, table: torch.Tensor, hist_idx, hist_len, cand_idx, cand_len,
                                 dtype=torch.float64, batch: int = 64):
    """get_final_second_attention_score (data_model_helper.py:416-443) for the
    WITH_HISTORY subset (history_bool all True, as scripts/eval.py:41-52 loads)."""
    u = user_vectors(sd, table, hist_idx, hist_len, batch=batch, dtype=dtype)
    s = cosine_scores(u, table, cand_idx, cand_len, dtype=dtype)
    s_np = s.detach().numpy()
    return {"user": u, "scores": s_np, "grouped_scores": rank_group_preds(s_np, cand_len)}


def classification_head(sd: dict, rows: torch.Tensor, dtype=torch.float64) -> torch.Tensor:
    """ClassificationHead.forward (modeling_utils.py:106-116): linear_3(relu(linear_2(relu(linear_1(e)))))."""
    g = lambda k: sd[k].detach().to(dtype)
    x = torch.relu(rows.to(dtype) @ g("linear_1.weight").T + g("linear_1.bias"))
    x = torch.relu(x @ g("linear_2.weight").T + g("linear_2.bias"))
    return x @ g("linear_3.weight").T + g("linear_3.bias")


def final_score(sd: dict, table: torch.Tensor, hist_idx, hist_len, cand_idx, cand_len, history_bool,
                classification_score: np.ndarray, alpha_param: float, dtype=torch.float64):
    """get_final_score (data_model_helper.py:272-301) with FinalAttention as the attention model:
    scores = classification_score[cand]; rows with history are overwritten by
    sigmoid(alpha)*cos + (1-sigmoid(alpha))*classification_score[cand] (WeightedSumModel, modeling_utils.py:158-165)."""
    hb = np.asarray(history_bool, dtype=bool)
    cand_len = np.asarray(cand_len)
    scores = np.asarray(classification_score, dtype=np.float64)[np.asarray(cand_idx)].copy()
    keep = np.repeat(hb, cand_len)
    u = user_vectors(sd, table, hist_idx, hist_len, dtype=dtype)
    cos = cosine_scores(u, table, np.asarray(cand_idx)[keep], cand_len[hb], dtype=dtype).numpy()
    a = 1.0 / (1.0 + math.exp(-alpha_param))
    scores[keep] = cos * a + scores[keep] * (1.0 - a)
    return {"scores": scores, "grouped_scores": rank_group_preds(scores, cand_len)}


# --------------------------------------------------------------------------
# Consumer -- MIND metrics (evaluation.py:13-98)
# --------------------------------------------------------------------------


def _auc_tie_aware(y_true: np.ndarray, y_score: np.ndarray) -> float:
    """sklearn.metrics.roc_auc_score for binary labels = Mann-Whitney U with
    ties counted 1/2 (trapezoid over distinct thresholds)."""
    pos = y_score[y_true > 0.5]
    neg = y_score[y_true <= 0.5]
    if len(pos) == 0 or len(neg) == 0:
        raise ValueError("Only one class present in y_true")
    gt = (pos[:, None] > neg[None, :]).sum()
    eq = (pos[:, None] == neg[None, :]).sum()
    return float((gt + 0.5 * eq) / (len(pos) * len(neg)))


def _order_desc(y_score: np.ndarray) -> np.ndarray:
    # evaluation.py:14,28 -- np.argsort(y_score)[::-1] (default quicksort kind)
    return np.argsort(y_score)[::-1]


def score_row(labels: Sequence[int], ranks: Sequence[float]):
    """evaluation.py:34-54: y_score = 1/rank, then AUC, MRR, nDCG@5, nDCG@10."""
    y_true = np.asarray(labels, dtype=np.float32)
    y_score = np.array([1.0 / r for r in ranks], dtype=np.float64)
    auc = _auc_tie_aware(y_true, y_score)
    order = _order_desc(y_score)
    yt = np.take(y_true, order)
    mrr = float(np.sum(yt / (np.arange(len(yt)) + 1)) / np.sum(y_true))

    def dcg(y, sc, k):
        o = np.argsort(sc)[::-1]
        t = np.take(y, o[:k])
        return np.sum((2**t - 1) / np.log2(np.arange(len(t)) + 2))

    nd5 = float(dcg(y_true, y_score, 5) / dcg(y_true, y_true, 5))
    nd10 = float(dcg(y_true, y_score, 10) / dcg(y_true, y_true, 10))
    return auc, mrr, nd5, nd10


def score(grouped_ranks, labels) -> dict:
    rows = [score_row(l, r) for l, r in zip(labels, grouped_ranks)]
    a = np.asarray(rows, dtype=np.float64)
    return {"auc": float(a[:, 0].mean()), "mrr": float(a[:, 1].mean()),
            "ndcg5": float(a[:, 2].mean()), "ndcg10": float(a[:, 3].mean()),
            "num_samples": len(rows)}


# --------------------------------------------------------------------------
# Input layout producer (data_utils.py:168-232) -- "next" row, SURVEY 8f.2
# --------------------------------------------------------------------------


def split_impressions_and_history(impressions: Sequence[str], history: Sequence[str]) -> dict:
    """First-appearance news ids over history-then-impression tokens; flat int32
    row-id arrays with owner row; int32 length lists; labels as object array."""
    label_present = "-" in impressions[0]
    pos: dict[str, int] = {}
    news_list: list[str] = []
    imp_ids: list[int] = []
    hist_ids: list[int] = []
    labels = []
    hist_len: list[int] = []
    imp_len: list[int] = []

    def rid(tok: str) -> int:
        r = pos.get(tok)
        if r is None:
            r = len(news_list)
            pos[tok] = r
            news_list.append(tok)
        return r

    for imp_row, hist_row in zip(impressions, history):
        if hist_row:
            toks = hist_row.split()
            hist_len.append(len(toks))
            hist_ids.extend(rid(t) for t in toks)
        toks = imp_row.split()
        if label_present:
            pairs = [t.split("-") for t in toks]
            labels.append(tuple(int(p[1]) for p in pairs))
            toks = [p[0] for p in pairs]
        imp_len.append(len(toks))
        imp_ids.extend(rid(t) for t in toks)
    owner = lambda lens: np.repeat(np.arange(len(lens), dtype=np.int32), lens).astype(np.int32)
    lab = np.empty(len(labels), dtype=object)
    for i, l in enumerate(labels):
        lab[i] = l
    return {
        "news_list": np.array(news_list),
        "impression_rev_ind_array": np.stack([np.array(imp_ids, dtype=np.int32), owner(imp_len)]),
        "impression_len_list": np.array(imp_len, dtype=np.int32),
        "history_rev_ind_array": np.stack([np.array(hist_ids, dtype=np.int32), owner(hist_len)]),
        "history_len_list": np.array(hist_len, dtype=np.int32),
        "labels": lab,
    }
