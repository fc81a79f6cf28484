"""Drop-in `NewAttention` user encoder (reference: news_rec_utils/attention.py:209-279).

In the reference every `MyLayer.forward` discards its attention / gated-MLP result and returns
`g_mlp_layernorm(hidden_states)` (attention.py:193), so the encoder is a chain of LayerNorms
(eps 1e-12) applied to each history slot independently; the pooling is the same per-dimension
exp-weighting as FinalAttention:  res = LN_k(...LN_1(e)); w = exp(linear1(res)) * mask;
out = sum_s res*w / (sum_s w + 1e-10)  (attention.py:266-272).  That is again a per-row transform
(X = res, E = exp(linear1(res))) followed by `nrb_score_rank`'s FINAL_ATTENTION pooling.

The dead attention / MLP parameters are kept (same names and shapes) so that reference checkpoints
load with `strict=True`.  Inference only.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import torch
from torch import nn

from . import _lib, config, ops


class _MyAttentionWeights(nn.Module):
    def __init__(self, hidden_size: int, num_attention_heads: int = 8):
        super().__init__()
        self.qkv_proj = nn.Linear(hidden_size, hidden_size * 3, bias=True)
        self.dropout = nn.Dropout(0)
        self.o_proj = nn.Linear(hidden_size, hidden_size, bias=True)


class _GatedMLPWeights(nn.Module):
    def __init__(self, hidden_size: int, intermediate_size: int = 3072):
        super().__init__()
        self.up_gate_proj = nn.Linear(hidden_size, intermediate_size * 2, bias=False)
        self.down_proj = nn.Linear(intermediate_size, hidden_size, bias=True)


class _MyLayerWeights(nn.Module):
    """Parameter container of attention.py:150-194; only g_mlp_layernorm is live."""

    def __init__(self, hidden_size: int, layer_norm_eps: float = 1e-12):
        super().__init__()
        self.attention = _MyAttentionWeights(hidden_size)
        self.g_mlp = _GatedMLPWeights(hidden_size)
        self.attn_layernorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.g_mlp_layernorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)


class _MyEncoderWeights(nn.Module):
    def __init__(self, hidden_size: int, num_hidden_layers: int):
        super().__init__()
        self.layer = nn.ModuleList([_MyLayerWeights(hidden_size) for _ in range(num_hidden_layers)])


class NewAttention(nn.Module):
    def __init__(self, hidden_size: Optional[int] = None, num_hidden_layers: int = 1, precision=None):
        super().__init__()
        d = config.REDUCED_DIM if hidden_size is None else hidden_size
        self.encoder = _MyEncoderWeights(d, num_hidden_layers)
        self.linear1 = nn.Linear(d, d)
        self.precision = precision

    def row_tables(self, rows: torch.Tensor, dtype: torch.dtype):
        """rows [n, d] on the device -> (X, E) tables in `dtype` (the per-row part of forward)."""
        dev = rows.device
        x = rows
        for layer in self.encoder.layer:
            ln = layer.g_mlp_layernorm
            x = ops.layer_norm(x.contiguous(), ln.weight.detach().to(dev, torch.float32).contiguous(),
                               ln.bias.detach().to(dev, torch.float32).contiguous(), ln.eps, out_dtype=dtype)
        x = x.to(dtype).contiguous()
        w = self.linear1.weight.detach().to(dev, dtype).contiguous()
        b = self.linear1.bias.detach().to(dev, torch.float32).contiguous()
        e = ops.linear(x, w, b, _lib.EPI_EXP, None, dtype)
        return x, e

    @torch.no_grad()
    def forward(self, embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise _lib.NrbError("NewAttention (nrb200) is inference only: call model.eval()")
        in_dev = embeddings.device
        dev = _lib.require_device(in_dev if in_dev.type == "cuda" else None)
        dtype = config.precision_dtype(self.precision)
        B, H, d = embeddings.shape
        with torch.cuda.device(dev):
            rows = embeddings.detach().to(device=dev, dtype=dtype).reshape(B * H, d).contiguous()
            x, e = self.row_tables(rows, dtype)
            user = ops.pool_masked_rows(x, e, attention_mask)
        return user if in_dev.type == "cuda" else user.to(in_dev)


def get_new_attention_model(model_path: Optional[Path] = None) -> NewAttention:
    """modeling_utils.py:426-431."""
    model = NewAttention(hidden_size=config.REDUCED_DIM)
    if model_path:
        model.load_state_dict(torch.load(model_path, weights_only=True))
    return model.to(config.DEVICE).eval()
