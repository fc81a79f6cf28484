"""Row-sharded news-embedding table across the GPUs of one box (BASELINE configs[4]).

Rank g owns rows [g*Ns, min(N, (g+1)*Ns)).  Both user encoders on the path are per-row functions
followed by a pooled reduction (FinalAttention is separable per row; latent-attention tokens are
independent), so each rank applies the dense per-row transform to ITS OWN shard -- tensor-core work,
perfectly partitioned -- and the transformed rows (and the raw rows candidates are scored against) are
all-gathered so that the bandwidth-bound gather/score/rank then runs locally on the rank's impression
shard against a full local copy.

all-gather = `nrb_push_rows`: the chunk just produced is written straight into every peer's copy of the
full table with 128-bit stores on peer-mapped symmetric memory (NVLink 5 / NVSwitch), launched on the
compute stream right behind the chunk's GEMMs so the transfer of chunk i overlaps the transform of chunk
i+1.  `gather="nccl"` keeps the plain `all_gather_into_tensor` variant for comparison / non-P2P setups.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from .config import LATENT_MAX_TOKENS, precision_dtype
from .engine import ScoringEngine, _final_attention_weights
from .sharding import table_shard_bounds


class ShardedTableEngine(ScoringEngine):
    def __init__(self, local_rows: torch.Tensor, n_rows_total: int, model: torch.nn.Module, precision=None,
                 device: Optional[torch.device] = None, group=None, gather: str = "p2p", chunk_rows: int = 16384):
        from .latent_attention import LatentAttentionModel

        if not dist.is_initialized():
            raise _lib.NrbError("ShardedTableEngine needs torch.distributed to be initialised")
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = _lib.require_device(device)
        self.dtype = precision_dtype(precision)
        self.model = model
        self.n_rows, self.dim = int(n_rows_total), int(local_rows.shape[1])
        self._streams = None
        r0, r1 = table_shard_bounds(self.n_rows, self.world)[self.rank]
        if local_rows.shape[0] != r1 - r0:
            raise _lib.NrbError(f"rank {self.rank} must hold rows [{r0},{r1}) = {r1 - r0} rows, got {local_rows.shape[0]}")
        latent = isinstance(model, LatentAttentionModel)
        self._latent = latent
        self.pool_mode = _lib.POOL_MEAN_L2 if latent else _lib.POOL_FINAL_ATTENTION
        self.gather, self.chunk_rows = gather, chunk_rows
        self._bounds = (r0, r1)
        n, d, dev = self.n_rows, self.dim, self.device
        with torch.cuda.device(dev):
            self._names = ["cand", "hist_x"] + ([] if latent else ["hist_e"])
            if gather == "p2p":
                import torch.distributed._symmetric_memory as symm_mem

                self._full, self._ptrs = {}, {}
                for nm in self._names:
                    t = symm_mem.empty((n, d), dtype=self.dtype, device=dev)
                    h = symm_mem.rendezvous(t, self.group)
                    self._full[nm], self._ptrs[nm] = t, list(h.buffer_ptrs)
            elif gather == "nccl":
                ns = table_shard_bounds(n, self.world)[0][1]
                self._full = {nm: torch.empty(self.world * ns, d, dtype=self.dtype, device=dev) for nm in self._names}
                self._ptrs = None
            else:
                raise ValueError(gather)
        self.build(local_rows)

    def build(self, local_rows: torch.Tensor) -> None:
        """Transform this rank's shard chunk by chunk and all-gather the chunks (collective: every rank calls it)."""
        r0, r1 = self._bounds
        n, d, dev = self.n_rows, self.dim, self.device
        latent, names, full, ptrs = self._latent, self._names, self._full, self._ptrs
        ns = table_shard_bounds(n, self.world)[0][1]
        with torch.cuda.device(dev):
            # nobody may still be scoring against the previous contents when peers start overwriting them
            torch.cuda.current_stream().synchronize()
            dist.barrier(group=self.group)
            if latent:
                fw = self.model.folded(self.dtype, dev)
            else:
                w = _final_attention_weights(self.model, self.dtype, dev)
            for c0 in range(0, r1 - r0, self.chunk_rows):
                c1 = min(r1 - r0, c0 + self.chunk_rows)
                chunk = local_rows[c0:c1].to(dev, non_blocking=True)
                chunk = chunk.to(self.dtype).contiguous()
                if latent:
                    outs = {"hist_x": ops.latent_forward(fw, chunk.view(c1 - c0, 1, d), None,
                                                         max_tokens=LATENT_MAX_TOKENS).view(c1 - c0, d)}  # fp32
                else:
                    x, e = ops.final_attention_rows(chunk, w, self.dtype)
                    outs = {"hist_x": x, "hist_e": e}
                outs["cand"] = chunk
                for nm in names:
                    if ptrs is not None:
                        ops.push_rows(outs[nm], ptrs[nm], self.dtype, r0 + c0, d)
                    else:
                        full[nm][self.rank * ns + c0:self.rank * ns + c1].copy_(outs[nm])
            if ptrs is None:
                for nm in names:
                    dist.all_gather_into_tensor(full[nm], full[nm][self.rank * ns:(self.rank + 1) * ns].clone(),
                                                group=self.group)
            # every rank's pushes have landed before anybody scores
            torch.cuda.current_stream().synchronize()
            dist.barrier(group=self.group)
            self.cand, self.hist_x = full["cand"][:n], full["hist_x"][:n]
            self.hist_e = None if latent else full["hist_e"][:n]
