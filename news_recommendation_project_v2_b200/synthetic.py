"""Synthetic MIND-shaped inputs.

Seeds follow the reference: `torch.manual_seed(1234)` (config.py:55) and
`np.random.default_rng(1234)` (scripts/eval.py:38).  Everything is generated on
the host with numpy so that the same bytes are produced on every box; the bulk
generators used by bench.py at full size have device-side variants there.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class Impressions:
    """CSR layout of split_impressions_and_history (data_utils.py:168-232)."""

    hist_idx: np.ndarray  # int32 [sum H]  -> history_rev_ind_array[0]
    hist_len: np.ndarray  # int32 [I]      -> history_len_list
    cand_idx: np.ndarray  # int32 [sum C]  -> impression_rev_ind_array[0]
    cand_len: np.ndarray  # int32 [I]      -> impression_len_list
    labels: np.ndarray  # object [I] of int tuples

    @property
    def n(self) -> int:
        return int(self.hist_len.shape[0])


def make_table(n_rows: int, dim: int, seed: int = 1234) -> torch.Tensor:
    """L2-normalised random table, fp32 CPU (save_emb.py writes a normalised
    table: data_model_helper.py:65-78)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(n_rows, dim, generator=g, dtype=torch.float32)
    return torch.nn.functional.normalize(t, p=2, dim=-1)


def make_impressions(n_imp: int, n_rows: int, h_max: int = 50, cand: str = "small",
                     seed: int = 1234, with_labels: bool = True) -> Impressions:
    """H_i = clip(Geometric(1/32), 1, h_max); C_i uniform[3,7] ('small', cfg 1/2)
    or clip(round(LogNormal(ln 30, 0.7)), 2, 300) ('large', cfg 4/5).  Counts are
    forced non-uniform (reference quirk a7: equal counts crash group_items'
    object array at data_model_helper.py:226)."""
    rng = np.random.default_rng(seed)
    hist_len = np.clip(rng.geometric(1.0 / 32.0, size=n_imp), 1, h_max).astype(np.int32)
    if cand == "small":
        cand_len = rng.integers(3, 8, size=n_imp).astype(np.int32)
    elif cand == "large":
        cand_len = np.clip(np.rint(rng.lognormal(np.log(30.0), 0.7, size=n_imp)), 2, 300).astype(np.int32)
    else:
        raise ValueError(cand)
    if n_imp > 1 and np.all(cand_len == cand_len[0]):
        cand_len[0] += 1
    if n_imp > 1 and np.all(hist_len == hist_len[0]):
        hist_len[0] = max(1, hist_len[0] - 1) if hist_len[0] > 1 else 2
    hist_idx = rng.integers(0, n_rows, size=int(hist_len.sum()), dtype=np.int64).astype(np.int32)
    cand_idx = rng.integers(0, n_rows, size=int(cand_len.sum()), dtype=np.int64).astype(np.int32)
    labels = np.empty(n_imp, dtype=object)
    if with_labels:
        for i in range(n_imp):
            c = int(cand_len[i])
            lab = (rng.random(c) < 0.04).astype(np.int64)
            lab[rng.integers(0, c)] = 1
            if lab.sum() == c:  # need at least one negative (evaluation.py:49)
                lab[(int(np.argmax(lab)) + 1) % c] = 0
            labels[i] = tuple(int(v) for v in lab)
    return Impressions(hist_idx, hist_len, cand_idx, cand_len, labels)


def make_long_history_impressions(n_imp: int, n_rows: int, h_max: int, seed: int) -> Impressions:
    """BASELINE configs[4] shaped impressions: the usual generator with the first impressions forced to the
    extremes (H = h_max, 1, h_max - 1) so that the longest history the config allows is always present."""
    imp = make_impressions(n_imp, n_rows, h_max=h_max, cand="large", seed=seed)
    rng = np.random.default_rng(seed + 1000)
    hist_len = imp.hist_len.copy()
    hist_len[0], hist_len[1], hist_len[2] = h_max, 1, h_max - 1
    hist_idx = rng.integers(0, n_rows, size=int(hist_len.sum()), dtype=np.int64).astype(np.int32)
    return Impressions(hist_idx, hist_len, imp.cand_idx, imp.cand_len, imp.labels)


def make_token_batch(batch: int, seq: int, dim: int, seed: int = 1234, min_len: int = 8):
    """Stage A input: x ~ randn(B,S,d) fp32, len_b ~ U[min_len, S], int32 mask."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, seq, dim, generator=g, dtype=torch.float32)
    lens = torch.randint(min(min_len, seq), seq + 1, (batch,), generator=g)
    mask = (torch.arange(seq)[None, :] < lens[:, None]).to(torch.int32)
    return x, mask


def csr_offsets(lengths: np.ndarray) -> np.ndarray:
    """Exclusive prefix sum, int64 [I+1]."""
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=off[1:])
    return off


def _linear_init(g: torch.Generator, out_f: int, in_f: int, bias: bool):
    # same scale as torch.nn.Linear's default init (uniform +-1/sqrt(in)), own stream
    bound = 1.0 / (in_f ** 0.5)
    w = (torch.rand(out_f, in_f, generator=g, dtype=torch.float32) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=g, dtype=torch.float32) * 2 - 1) * bound if bias else None
    return w, b


def make_latent_state_dict(dim: int, num_latents: int, heads: int = 8, dim_head: int = 512,
                           seed: int = 1234) -> dict:
    """Random weights under the reference's state_dict keys.
    LayerNorm affine is perturbed away from (1, 0) so that it is exercised."""
    g = torch.Generator().manual_seed(seed)
    inner = heads * dim_head
    sd = {}
    sd["latents"] = torch.randn(num_latents, dim, generator=g, dtype=torch.float32)
    p = "cross_attend_blocks.0."
    sd[p + "fn.to_q.weight"], _ = _linear_init(g, inner, dim, False)
    sd[p + "fn.to_kv.weight"], _ = _linear_init(g, 2 * inner, dim, False)
    sd[p + "fn.to_out.weight"], _ = _linear_init(g, dim, inner, False)
    for nm in ("norm", "norm_context"):
        sd[p + nm + ".weight"] = 1.0 + 0.1 * torch.randn(dim, generator=g)
        sd[p + nm + ".bias"] = 0.1 * torch.randn(dim, generator=g)
    p = "cross_attend_blocks.1."
    sd[p + "fn.net.0.weight"], sd[p + "fn.net.0.bias"] = _linear_init(g, 8 * dim, dim, True)
    sd[p + "fn.net.2.weight"], sd[p + "fn.net.2.bias"] = _linear_init(g, dim, 4 * dim, True)
    sd[p + "norm.weight"] = 1.0 + 0.1 * torch.randn(dim, generator=g)
    sd[p + "norm.bias"] = 0.1 * torch.randn(dim, generator=g)
    return sd


def make_final_attention_state_dict(dim: int, hidden: int = 4096, seed: int = 1234) -> dict:
    """Random weights under FinalAttention's keys."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    sd["linear1.weight"], sd["linear1.bias"] = _linear_init(g, hidden, dim, True)
    sd["linear2.weight"], sd["linear2.bias"] = _linear_init(g, hidden, hidden, True)
    sd["linear3.weight"], sd["linear3.bias"] = _linear_init(g, dim, hidden, True)
    sd["linear4.weight"], sd["linear4.bias"] = _linear_init(g, hidden, dim, True)
    sd["linear5.weight"], _ = _linear_init(g, dim, hidden, False)
    return sd
