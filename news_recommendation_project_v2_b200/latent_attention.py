"""Drop-in `LatentAttentionModel` (reference: news_rec_utils/latent_attention.py:77-171).

Same constructor behaviour (dims from the config globals, 8 heads x 512, 64
latents unless EMBEDDING_DIM == 4096), same `forward(embeddings, attention_mask)`
contract and the SAME state_dict keys, so `get_latent_attention_model`'s
`load_state_dict(torch.load(path, weights_only=True))` keeps working
(modeling_utils.py:151-155).  The parameter containers below only hold weights;
the arithmetic is the fused CUDA path of `nrb_latent_forward` (tokens attend to
the learned latents, softmax, value product, GEGLU MLP, masked mean, L2 norm).

Inference only: outputs carry no autograd graph (training is out of scope).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib, config, ops


class _AttentionWeights(nn.Module):
    """Parameters of the cross-attention (keys: to_q / to_kv / to_out .weight)."""

    def __init__(self, query_dim: int, context_dim: int, heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_kv = nn.Linear(context_dim, 2 * inner, bias=False)
        self.to_out = nn.Linear(inner, query_dim, bias=False)


class _FeedForwardWeights(nn.Module):
    """Parameters of the GEGLU MLP (keys: net.0.{weight,bias}, net.2.{weight,bias})."""

    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([nn.Linear(dim, dim * mult * 2), nn.Identity(), nn.Linear(dim * mult, dim)])


class _PreNormBlock(nn.Module):
    """Parameters of one pre-norm residual block (keys: fn.*, norm.*, norm_context.*)."""

    def __init__(self, dim: int, fn: nn.Module, context_dim: Optional[int] = None):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)
        self.norm_context = nn.LayerNorm(context_dim) if context_dim is not None else None


class LatentAttentionModel(nn.Module):
    def __init__(self, dim: Optional[int] = None, num_latents: Optional[int] = None, heads: Optional[int] = None,
                 dim_head: Optional[int] = None, precision=None):
        super().__init__()
        d = config.REDUCED_DIM if dim is None else dim
        # latent_attention.py:91-104 -- the 4096-d branch uses 32 latents, 2 heads x 32
        if dim is None and config.EMBEDDING_DIM == 4096:
            dflt = (32, 2, 32)
        else:
            dflt = (64, 8, 512)
        L = dflt[0] if num_latents is None else num_latents
        h = dflt[1] if heads is None else heads
        dh = dflt[2] if dim_head is None else dim_head
        self.cross_attend_blocks = nn.ModuleList([
            _PreNormBlock(d, _AttentionWeights(d, d, h, dh), context_dim=d),
            _PreNormBlock(d, _FeedForwardWeights(d)),
        ])
        self.output_normalize = True
        self.register_parameter("latents", nn.Parameter(torch.randn(L, d)))
        self.precision = precision
        self._folded = None
        self._folded_key = None

    # -- kernel-ready weights, refolded whenever a parameter changes -------------------------------
    def _fingerprint(self, dtype, device) -> tuple:
        return (str(dtype), str(device)) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def folded(self, precision=None, device=None) -> ops.FoldedLatent:
        dev = _lib.require_device(device)
        dtype = config.precision_dtype(self.precision if precision is None else precision)
        key = self._fingerprint(dtype, dev)
        if self._folded is None or self._folded_key != key:
            attn = self.cross_attend_blocks[0].fn
            with torch.cuda.device(dev):
                self._folded = ops.latent_fold(self.state_dict(), attn.heads, attn.dim_head, dtype, dev)
            self._folded_key = key
        return self._folded

    @torch.no_grad()
    def forward_packed(self, tokens: torch.Tensor, item_offsets: torch.Tensor) -> torch.Tensor:
        """Varlen variant: `tokens` [T, d] holds only real tokens, `item_offsets` [B+1] their CSR offsets ->
        [B, d] unit-norm vectors (the layout of a packed token store; no padding, no mask)."""
        in_dev = tokens.device
        dev = _lib.require_device(in_dev if in_dev.type == "cuda" else None)
        fw = self.folded(None, dev)
        with torch.cuda.device(dev):
            x = tokens.detach().to(dev)
            if x.dtype not in (torch.float32, torch.bfloat16):
                x = x.float()
            out = ops.latent_forward_packed(fw, x.contiguous(), item_offsets, max_tokens=config.LATENT_MAX_TOKENS)
        return out if in_dev.type == "cuda" else out.to(in_dev)

    @torch.no_grad()
    def forward(self, embeddings: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B,S,d] (+ mask [B,S]) -> [B,d] unit-norm, or un-pooled [B,S,d] when the mask is None."""
        if not self.output_normalize:
            raise _lib.NrbError("output_normalize=False is not supported by the fused kernel")
        in_dev = embeddings.device
        dev = _lib.require_device(in_dev if in_dev.type == "cuda" else None)
        fw = self.folded(None, dev)
        with torch.cuda.device(dev):
            x = embeddings.detach().to(dev)
            if x.dtype not in (torch.float32, torch.bfloat16):
                x = x.float()
            x = x.contiguous()
            m = None
            if attention_mask is not None:
                m = (attention_mask.to(dev) != 0).to(torch.int32).contiguous()
            out = ops.latent_forward(fw, x, m, max_tokens=config.LATENT_MAX_TOKENS)
        return out if in_dev.type == "cuda" else out.to(in_dev)
