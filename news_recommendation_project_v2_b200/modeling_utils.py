"""Drop-in user encoder + model factories (reference: news_rec_utils/modeling_utils.py).

  FinalAttention              modeling_utils.py:175-228
  get_final_attention_model   modeling_utils.py:274-279
  get_latent_attention_model  modeling_utils.py:151-155
  get_model_eval              modeling_utils.py:402-417

Same constructor signature, parameter names and state_dict keys as the reference.
`forward` runs on the CUDA kernels: the five Linear layers go through
`nrb_final_attention_rows` (tcgen05 / FFMA), the exp-weighted masked pooling
through `nrb_score_rank`'s FINAL_ATTENTION pooling.  Inference only.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import torch
from torch import nn

from . import _lib, config, ops
from .latent_attention import LatentAttentionModel


class FinalAttention(nn.Module):
    def __init__(self, reduced_dim: int, hidden_dim: int, precision=None):
        super().__init__()
        self.linear1 = nn.Linear(reduced_dim, hidden_dim)
        self.dropout1 = nn.Dropout(0.1)  # inactive in eval; kept for attribute parity
        self.linear2 = nn.Linear(hidden_dim, hidden_dim)
        self.dropout2 = nn.Dropout(0.1)
        self.linear3 = nn.Linear(hidden_dim, reduced_dim)
        self.linear4 = nn.Linear(reduced_dim, hidden_dim)
        self.dropout3 = nn.Dropout(0.1)
        self.linear5 = nn.Linear(hidden_dim, reduced_dim, bias=False)
        self.precision = precision

    @torch.no_grad()
    def forward(self, embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        """embeddings [B,H,d] (already masked), attention_mask [B,H] -> [B,d] (fp32)."""
        if self.training:
            raise _lib.NrbError("FinalAttention (nrb200) is inference only: call model.eval() (dropout / autograd "
                                "are out of scope)")
        from .engine import _final_attention_weights

        in_dev = embeddings.device
        dev = _lib.require_device(in_dev if in_dev.type == "cuda" else None)
        dtype = config.precision_dtype(self.precision)
        B, H, d = embeddings.shape
        with torch.cuda.device(dev):
            rows = embeddings.detach().to(device=dev, dtype=dtype).reshape(B * H, d).contiguous()
            w = _final_attention_weights(self, dtype, dev)
            # the MLPs act on each history slot independently (modeling_utils.py:218-222)
            x, e = ops.final_attention_rows(rows, w, dtype)
            user = ops.pool_masked_rows(x, e, attention_mask)
        return user if in_dev.type == "cuda" else user.to(in_dev)


class ClassificationHead(nn.Module):
    """Click classifier for news without user history (modeling_utils.py:106-116): three Linear layers with
    ReLU in between, out_dim scores per row.  Same parameter names (linear_1/2/3)."""

    def __init__(self, in_dim: int, hidden_dim: int, out_dim: int, precision=None):
        super().__init__()
        self.linear_1 = nn.Linear(in_features=in_dim, out_features=hidden_dim)
        self.linear_2 = nn.Linear(in_features=hidden_dim, out_features=hidden_dim)
        self.linear_3 = nn.Linear(in_features=hidden_dim, out_features=out_dim)
        self.precision = precision

    @torch.no_grad()
    def forward(self, embeddings: torch.Tensor) -> torch.Tensor:
        in_dev = embeddings.device
        dev = _lib.require_device(in_dev if in_dev.type == "cuda" else None)
        dtype = config.precision_dtype(self.precision)
        lead = embeddings.shape[:-1]
        with torch.cuda.device(dev):
            x = embeddings.detach().to(device=dev, dtype=dtype).reshape(-1, embeddings.shape[-1]).contiguous()
            wt = lambda l: l.weight.detach().to(dev, dtype).contiguous()
            bs = lambda l: l.bias.detach().to(dev, torch.float32).contiguous()
            h = ops.linear(x, wt(self.linear_1), bs(self.linear_1), _lib.EPI_RELU, None, dtype)
            h = ops.linear(h, wt(self.linear_2), bs(self.linear_2), _lib.EPI_RELU, None, dtype)
            out_dim = self.linear_3.out_features
            pad = (-out_dim) % 32  # the dense kernels want N % 32 == 0: zero rows are appended and sliced away
            w3, b3 = wt(self.linear_3), bs(self.linear_3)
            if pad:
                w3 = torch.cat([w3, torch.zeros(pad, w3.shape[1], dtype=dtype, device=dev)]).contiguous()
                b3 = torch.cat([b3, torch.zeros(pad, dtype=torch.float32, device=dev)]).contiguous()
            y = ops.linear(h, w3, b3, _lib.EPI_NONE, None, torch.float32)[:, :out_dim]
        y = y.reshape(*lead, out_dim)
        return y if in_dev.type == "cuda" else y.to(in_dev)


def get_classification_head(model_path: Optional[Path] = None) -> ClassificationHead:
    """modeling_utils.py:139-148."""
    model = ClassificationHead(in_dim=config.EMBEDDING_DIM, hidden_dim=config.EMBEDDING_DIM, out_dim=1)
    if model_path:
        model.load_state_dict(torch.load(model_path, weights_only=True))
    return model.to(config.DEVICE).eval()


class WeightedSumModel(nn.Module):
    """sigmoid(alpha) blend of the cosine score and the classification baseline (modeling_utils.py:158-165).
    The blend itself is fused into `nrb_score_rank`; `forward` is kept for API parity on small tensors."""

    def __init__(self):
        super().__init__()
        self.alpha = nn.Parameter(torch.tensor(0.0))

    def blend_alpha(self) -> float:
        return float(torch.sigmoid(self.alpha.detach().float()).item())

    @torch.no_grad()
    def forward(self, cos_sim, baseline):
        alpha = torch.sigmoid(self.alpha)
        return cos_sim * alpha + baseline * (1 - alpha)


def get_weighted_sum_model(model_path: Optional[Path] = None) -> WeightedSumModel:
    model = WeightedSumModel()
    if model_path:
        model.load_state_dict(torch.load(model_path, weights_only=True))
    return model.to(config.DEVICE)


def get_final_attention_model(model_path: Optional[Path] = None) -> FinalAttention:
    model = FinalAttention(reduced_dim=config.REDUCED_DIM, hidden_dim=4096)
    if model_path:
        model.load_state_dict(torch.load(model_path, weights_only=True))
    return model.to(config.DEVICE).eval()


def get_latent_attention_model(model_path: Optional[Path] = None) -> LatentAttentionModel:
    model = LatentAttentionModel()
    if model_path:
        model.load_state_dict(torch.load(model_path, weights_only=True))
    return model.to(config.DEVICE).eval()


def get_model_eval(dataloader, model: nn.Module) -> torch.Tensor:
    """Run `model` over a loader, results concatenated on the host (modeling_utils.py:402-417)."""
    outs = []
    model.eval()
    with torch.no_grad():
        for item in dataloader:
            if isinstance(item, (tuple, list)):
                res = model(*[t.to(config.DEVICE) for t in item])
            else:
                res = model(item.to(config.DEVICE))
            outs.append(res.detach().cpu())
    return torch.cat(outs)
