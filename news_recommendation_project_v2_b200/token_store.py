"""Token-state store -> packed varlen layout (the data format in front of Stage A).

The reference keeps one `torch.save`d `[n_tok, d]` tensor of valid-token hidden states per news item in a
sqlite table `tensors(id INTEGER PRIMARY KEY, data BLOB)` (modeling_utils.py:456-473) and, per batch, reads
the blobs back, pads them to the batch maximum and builds a mask on the CPU
(data_utils.py:878-890, 929-933, 753-781).  Here the store is read ONCE into a packed token matrix plus CSR
offsets, which `LatentAttentionModel.forward_packed` consumes directly -- no padding, no mask.

Host-side I/O only (sqlite3 + torch.load); all arithmetic stays in the CUDA path.
"""
from __future__ import annotations

import io
import sqlite3
from typing import Iterable, Optional

import torch


def read_token_store(db_name: str, ids: Optional[Iterable[int]] = None, dtype: torch.dtype = torch.bfloat16):
    """Rows of the reference's token store (0-based item indices, `id = index + 1` as in get_embeds_from_db)
    -> (tokens [T, d] in `dtype`, offsets int64 [B+1]) in the order of `ids` (default: all, ascending id)."""
    conn = sqlite3.connect(db_name)
    try:
        if ids is None:
            rows = conn.execute("SELECT id, data FROM tensors ORDER BY id;").fetchall()
            blobs = [r[1] for r in rows]
        else:
            ids = [int(i) for i in ids]
            blobs = []
            for i in ids:  # keep the caller's order (and duplicates), unlike `id IN (...)`
                row = conn.execute("SELECT data FROM tensors WHERE id = ?;", (i + 1,)).fetchone()
                if row is None:
                    raise IndexError(f"token store has no item {i}")
                blobs.append(row[0])
    finally:
        conn.close()
    parts = [torch.load(io.BytesIO(b), weights_only=True) for b in blobs]
    offsets = torch.zeros(len(parts) + 1, dtype=torch.int64)
    if parts:
        offsets[1:] = torch.cumsum(torch.tensor([p.shape[0] for p in parts], dtype=torch.int64), 0)
        tokens = torch.cat([p.to(dtype) for p in parts], dim=0).contiguous()
    else:
        tokens = torch.zeros(0, 0, dtype=dtype)
    return tokens, offsets


def write_token_store(db_name: str, items: Iterable[torch.Tensor]) -> None:
    """Write `[n_tok, d]` tensors in the reference's format (store_text_embed_full_eval, modeling_utils.py:456-473)."""
    with sqlite3.connect(db_name) as conn:
        conn.execute("DROP TABLE IF EXISTS tensors;")
        conn.execute("CREATE TABLE tensors (id INTEGER PRIMARY KEY, data BLOB)")
        for t in items:
            buf = io.BytesIO()
            torch.save(t.detach().cpu(), buf)
            conn.execute("INSERT INTO tensors (data) VALUES (?)", (buf.getvalue(),))
    conn.close()


def apply_token_attn(model, db_name: str, num_samples: int, chunk_items: int = 4096) -> torch.Tensor:
    """Drop-in for data_model_helper.py:390-413: pooled vectors [num_samples, d] (CPU) for items 0..num_samples-1
    of the store, through `model.forward_packed` (a LatentAttentionModel)."""
    outs = []
    for i0 in range(0, num_samples, chunk_items):
        ids = range(i0, min(num_samples, i0 + chunk_items))
        tokens, offsets = read_token_store(db_name, ids)
        outs.append(model.forward_packed(tokens, offsets).cpu())
    return torch.cat(outs) if outs else torch.zeros(0)
