"""Impression sharding across the GPUs of one box.

Impressions are independent units: each rank scores a contiguous block against its replica of the
table, so there is NO collective on the data path.  Blocks are balanced by the bytes they read
(`r*H_i + C_i` rows), not by impression count.  After scoring, the only exchanges are an ordered
variable-size gather of the flat score / rank arrays (when the host wants them all) and a 5-scalar
SUM all-reduce of the metric sums.  Works with any torch.distributed backend (NCCL on GPUs, gloo in
the CPU tests).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch
import torch.distributed as dist

from .synthetic import csr_offsets


def partition_impressions(hist_len: np.ndarray, cand_len: np.ndarray, world: int, rows_per_slot: int = 2):
    """Contiguous [start, end) blocks with (almost) equal sum of rows_per_slot*H_i + C_i."""
    cost = rows_per_slot * np.asarray(hist_len, dtype=np.int64) + np.asarray(cand_len, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(cost)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(cum, total * r / world, side="left")))
    bounds.append(len(cost))
    bounds = np.maximum.accumulate(np.minimum(bounds, len(cost)))
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world)]


def shard_impressions(hist_idx, hist_len, cand_idx, cand_len, start: int, end: int):
    """Views of the CSR arrays for impressions [start, end)."""
    h_off, c_off = csr_offsets(np.asarray(hist_len)), csr_offsets(np.asarray(cand_len))
    return (hist_idx[h_off[start]:h_off[end]], hist_len[start:end],
            cand_idx[c_off[start]:c_off[end]], cand_len[start:end])


def gather_ordered(local: torch.Tensor) -> torch.Tensor:
    """Concatenate variable-length 1-D tensors from all ranks in rank order (on every rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes) if sizes else 0
    buf = torch.zeros(mx, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])


def reduce_sums(sums: Sequence[float], device=None) -> list:
    """SUM all-reduce of a few float64 scalars (metric sums + counts)."""
    t = torch.tensor(list(sums), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def table_shard_bounds(n_rows: int, world: int):
    """Row ownership of the row-sharded table: rank g owns [g*Ns, min(N, (g+1)*Ns)), Ns = ceil(N / world)."""
    ns = (n_rows + world - 1) // world
    return [(min(n_rows, g * ns), min(n_rows, (g + 1) * ns)) for g in range(world)]
