"""Module constants, mirroring the reference's `news_rec_utils/config.py:19-43`.

The reference reads model dimensions from these globals at construction time
(latent_attention.py:91-113); the drop-in modules accept keyword overrides and
fall back to these values.
"""
from __future__ import annotations

import os
from enum import Enum

import torch


class NewsDataset(Enum):  # config.py:5-10 -- names of the cached tables (`<value>.pt`)
    MINDsmall_train = "MINDsmall_train"
    MINDsmall_dev = "MINDsmall_dev"
    MINDlarge_train = "MINDlarge_train"
    MINDlarge_dev = "MINDlarge_dev"
    MINDlarge_test = "MINDlarge_test"


class DataSubset(Enum):  # config.py:13-16
    WITH_HISTORY = "with_history"
    WITHOUT_HISTORY = "without_history"
    ALL = "all"

DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")  # config.py:19
EMBEDDING_DIM = 1024  # config.py:29
REDUCED_DIM = EMBEDDING_DIM  # config.py:31
IMPRESSION_MAXLEN = 600  # config.py:33
NUM_WORKERS = 4  # config.py:43
TORCH_DTYPE = torch.float32  # config.py:39

# Arithmetic of the CUDA hot path:
#   "bf16": bf16 tables / weights, tcgen05 tensor cores with fp32 accumulation (throughput path)
#   "fp32": fp32 tables / weights, FFMA accumulation (the reference's own arithmetic; parity path)
#   "fp32x3": fp32 tables; the FinalAttention row transform runs on the tensor cores as split-bf16 GEMMs
#             (hi/lo operand pairs, three products accumulated in fp32: ~2^-17 relative per operand); every other
#             dense op of that mode uses the fp32 FFMA kernels
PRECISION = os.environ.get("NRB200_PRECISION", "bf16")

# Tokens processed per latent-attention chunk (bounds the workspace: ~23 KB / token in bf16 at d=768, L=512, i.e.
# 12 GB at the default; smaller calls allocate only what they need).  Larger chunks amortise the ~130 us fixed cost
# of the ten kernels of a chunk: 926 TFLOP/s at 49 k tokens, 980 at 98 k, 997 at 262 k (round 1, same box); round 2,
# same box: 1,064-1,067 at 262 k, 1,080-1,083 at 524 k, no further gain at 1.3 M.
LATENT_MAX_TOKENS = int(os.environ.get("NRB200_LATENT_MAX_TOKENS", "524288"))


def precision_dtype(precision: str | torch.dtype | None = None) -> torch.dtype:
    p = PRECISION if precision is None else precision
    if isinstance(p, torch.dtype):
        if p in (torch.float32, torch.bfloat16):
            return p
        raise ValueError(f"unsupported precision {p}")
    if p in ("bf16", "bfloat16"):
        return torch.bfloat16
    if p in ("fp32", "float32", "fp32x3"):
        return torch.float32
    raise ValueError(f"unsupported precision {p!r} (use 'bf16', 'fp32' or 'fp32x3')")


def precision_is_split(precision: str | torch.dtype | None = None) -> bool:
    return (PRECISION if precision is None else precision) == "fp32x3"
