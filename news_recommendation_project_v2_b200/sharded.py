"""Row-sharded news-embedding table across the GPUs of one box (BASELINE configs[4]).

Rank g owns rows [g*Ns, min(N, (g+1)*Ns)).  Both user encoders on the path are per-row functions
followed by a pooled reduction (FinalAttention is separable per row; latent-attention tokens are
independent), so each rank applies the dense per-row transform to ITS OWN shard -- tensor-core work,
perfectly partitioned -- and the transformed rows (and the raw rows candidates are scored against) are
all-gathered so that the bandwidth-bound gather/score/rank then runs locally on the rank's impression
shard against a full local copy.

all-gather = `nrb_push_rows`: the chunk just produced is written straight into every peer's copy of the
full table with 128-bit stores on peer-mapped symmetric memory (NVLink 5 / NVSwitch), launched on a
high-priority side stream as soon as the chunk's GEMMs finish, so the transfer of chunk i overlaps the
transform of chunk i+1 (the store-only kernel co-resides with the persistent GEMM CTAs where their
register footprint leaves room).  `gather="dma"` = `nrb_push_bytes`: the transform writes its chunk into
the local copy and the copy engines forward it to the peers, which needs no SM resources at all.
`gather="nccl"` keeps the plain `all_gather_into_tensor` variant for comparison / non-P2P setups.
`gather="nvls"` is the fused compute + collective form: the rows of chunk i are multicast to every rank (one
`multimem.st` per 16 bytes through the NVSwitch) by the SPARE CONTROL WARP of the persistent tcgen05 GEMM CTAs that
compute chunk i+1 (`nrb_push_attach`), so the all-gather costs no copy engine, no extra kernel and no SM.
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
                 device: Optional[torch.device] = None, group=None, gather: str = "auto", chunk_rows: int = 32768,
                 cand_table: Optional[torch.Tensor] = None):
        """`cand_table`: the full RAW table, already resident on this device in the engine dtype (BASELINE
        configs[3] sharded over ranks: the table is replicated, only its per-row transform is partitioned).
        Then only the transformed rows are all-gathered."""
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
        self._comm = None
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
            self._names = ([] if cand_table is not None else ["cand"]) + ["hist_x"] + ([] if latent else ["hist_e"])
            self._cand_table = cand_table
            if cand_table is not None and (tuple(cand_table.shape) != (n, d) or cand_table.dtype != self.dtype
                                           or cand_table.device != dev):
                raise _lib.NrbError("cand_table must be the full [n_rows, dim] table on the engine device / dtype")
            self._tick = torch.zeros(1, dtype=torch.int32, device=dev)
            if gather in ("p2p", "dma", "none", "nvls", "auto"):
                import torch.distributed._symmetric_memory as symm_mem

                self._full, self._ptrs, self._mc = {}, {}, {}
                for nm in self._names:
                    t = symm_mem.empty((n, d), dtype=self.dtype, device=dev)
                    h = symm_mem.rendezvous(t, self.group)
                    self._full[nm], self._ptrs[nm] = t, list(h.buffer_ptrs)
                    self._mc[nm] = int(getattr(h, "multicast_ptr", 0) or 0)
                if gather == "nvls" and not all(self._mc.values()):
                    raise _lib.NrbError("gather='nvls' needs NVLink multicast (NVLS) support for symmetric memory; "
                                        "use gather='dma'")
                if gather == "auto":  # fused in-GEMM multicast where the fabric offers it, copy engines otherwise
                    gather = "nvls" if all(self._mc.values()) else "dma"
                    self.gather = gather
            elif gather == "nccl":
                ns = table_shard_bounds(n, self.world)[0][1]
                self._full = {nm: torch.empty(self.world * ns, d, dtype=self.dtype, device=dev) for nm in self._names}
                self._ptrs = None
            else:
                raise ValueError(gather)
        self.build(local_rows)

    def _comm_streams(self) -> list:
        if self._comm is None:
            n = 1 if self.gather == "p2p" else max(1, min(self.world - 1, 2))  # >2 concurrent copies slow the GEMMs (push_bench)
            self._comm = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(n)]
        return self._comm

    def build(self, local_rows: torch.Tensor) -> None:
        """Transform this rank's shard chunk by chunk and all-gather the chunks (collective: every rank calls it)."""
        r0, r1 = self._bounds
        n, d, dev = self.n_rows, self.dim, self.device
        latent, names, full, ptrs = self._latent, self._names, self._full, self._ptrs
        ns = table_shard_bounds(n, self.world)[0][1]
        with torch.cuda.device(dev):
            # nobody may still be scoring against the previous contents when peers start overwriting them
            self._stream_barrier()
            if latent:
                fw = self.model.folded(self.dtype, dev)
            else:
                from .engine import _model_fingerprint

                fp = _model_fingerprint(self.model)  # kernel-ready weights stay resident until a parameter changes
                if getattr(self, "_fa_weights_key", None) != fp:
                    self._fa_weights = _final_attention_weights(self.model, self.dtype, dev)
                    self._fa_weights_key = fp
                w = self._fa_weights
            compute = torch.cuda.current_stream()
            if self.gather == "nvls":
                self._build_nvls(local_rows, fw if latent else None, None if latent else w)
                self._stream_barrier()
                self._padded = None
                self.cand = self._cand_table if self._cand_table is not None else full["cand"][:n]
                self.hist_x = full["hist_x"][:n]
                self.hist_e = None if latent else full["hist_e"][:n]
                return
            dma = self.gather in ("dma", "none")
            if self.gather == "none":  # diagnosis only: peers are never written (every destination = local copy)
                ptrs = {nm: [p[self.rank]] * self.world for nm, p in ptrs.items()}
            comms = self._comm_streams() if ptrs is not None else []
            row_bytes = d * torch.empty(0, dtype=self.dtype).element_size()

            def after_compute():
                ready = torch.cuda.Event()
                ready.record(compute)
                for st in comms:
                    st.wait_event(ready)

            def push(nm, src, row0):
                # "p2p": store kernel on a high-priority side stream -- it co-resides with the next chunk's persistent
                # GEMM CTAs wherever their register footprint leaves room
                after_compute()
                src.record_stream(comms[0])
                with torch.cuda.stream(comms[0]):
                    ops.push_rows(src, ptrs[nm], self.dtype, row0, d)

            def push_dma(nm, src, row0, transient=True):
                # "dma": one peer copy per destination on the copy engines (no SM resources at all), destinations
                # rotated by rank so that no GPU is everybody's target at the same moment
                after_compute()
                for k in range(self.world):
                    g = (self.rank + 1 + k) % self.world
                    st = comms[k % len(comms)]
                    if transient:  # allocator-owned temporaries must outlive the copy
                        src.record_stream(st)
                    with torch.cuda.stream(st):
                        ops.push_bytes(src, [ptrs[nm][g]], row0 * row_bytes)

            for c0 in range(0, r1 - r0, self.chunk_rows):
                c1 = min(r1 - r0, c0 + self.chunk_rows)
                chunk = local_rows[c0:c1].to(dev, non_blocking=True)
                chunk = chunk.to(self.dtype).contiguous()
                g0, g1 = r0 + c0, r0 + c1
                if "cand" not in names:
                    pass  # raw table replicated: nothing to gather for the candidates
                elif ptrs is None:
                    full["cand"][self.rank * ns + c0:self.rank * ns + c1].copy_(chunk)
                else:
                    (push_dma if dma else push)("cand", chunk, g0)  # raw rows do not wait for the transform
                if latent:
                    outs = {"hist_x": ops.latent_forward(fw, chunk.view(c1 - c0, 1, d), None,
                                                         max_tokens=LATENT_MAX_TOKENS).view(c1 - c0, d)}  # fp32
                    if dma:  # fp32 -> table dtype into the local copy; the copy engines take it from there
                        ops.push_rows(outs["hist_x"], [ptrs["hist_x"][self.rank]], self.dtype, g0, d)
                elif dma:
                    ops.final_attention_rows(chunk, w, self.dtype, x_out=full["hist_x"][g0:g1], e_out=full["hist_e"][g0:g1])
                else:
                    x, e = ops.final_attention_rows(chunk, w, self.dtype)
                    outs = {"hist_x": x, "hist_e": e}
                for nm in names:
                    if nm == "cand":
                        continue
                    if ptrs is None:
                        full[nm][self.rank * ns + c0:self.rank * ns + c1].copy_(outs[nm])
                    elif dma:
                        push_dma(nm, full[nm][g0:g1], g0, transient=False)
                    else:
                        push(nm, outs[nm], g0)
            for st in comms:
                done = torch.cuda.Event()
                done.record(st)
                compute.wait_event(done)
            if ptrs is None:
                for nm in names:
                    dist.all_gather_into_tensor(full[nm], full[nm][self.rank * ns:(self.rank + 1) * ns].clone(),
                                                group=self.group)
            # every rank's pushes have landed before anybody scores
            self._stream_barrier()
            self._padded = None
            self.cand = self._cand_table if self._cand_table is not None else full["cand"][:n]
            self.hist_x = full["hist_x"][:n]
            self.hist_e = None if latent else full["hist_e"][:n]

    def _build_nvls(self, local_rows: torch.Tensor, fw, w) -> None:
        """Fused compute + collective build: chunk i's rows are multicast to every rank by the spare warp of the GEMM
        CTAs that compute chunk i+1; the last chunk leaves through the store-only flush kernel."""
        r0, r1 = self._bounds
        d, dev, mc = self.dim, self.device, self._mc
        try:
            self._build_nvls_chunks(local_rows, fw, w, r0, r1, d, dev, mc)
        except BaseException:
            ops.push_cancel()  # nothing stale may ride in a later, unrelated GEMM launch
            raise

    def _build_nvls_chunks(self, local_rows, fw, w, r0, r1, d, dev, mc) -> None:
        prev = None
        for c0 in range(0, r1 - r0, self.chunk_rows):
            c1 = min(r1 - r0, c0 + self.chunk_rows)
            chunk = local_rows[c0:c1].to(dev, non_blocking=True)
            if chunk.dtype != self.dtype:
                chunk = ops.convert_rows(chunk.contiguous(), self.dtype)
            chunk = chunk.contiguous()
            g0 = r0 + c0
            if prev is not None:
                ops.push_attach(prev, 4 if self._latent else 5)  # the GEMM launches of one internal row block
            segs = [(chunk, mc["cand"], self.dtype, g0, d)] if "cand" in self._names else []
            if self._latent:
                out = ops.latent_forward(fw, chunk.view(c1 - c0, 1, d), None, max_tokens=LATENT_MAX_TOKENS)
                segs.append((out.view(c1 - c0, d), mc["hist_x"], self.dtype, g0, d))  # fp32 -> table dtype on the way
            else:
                x, e = ops.final_attention_rows(chunk, w, self.dtype)
                segs += [(x, mc["hist_x"], self.dtype, g0, d), (e, mc["hist_e"], self.dtype, g0, d)]
            ops.push_flush()  # rows of the previous chunk that no GEMM picked up
            prev = segs  # keeps the sources alive until the next chunk's kernels are queued behind them
        if prev is not None:
            ops.push_attach(prev, 1)
            ops.push_flush()

    def _stream_barrier(self) -> None:
        """Device-side barrier: a 4-byte NCCL all-reduce ordered on the CURRENT stream.  It completes on a rank only
        after every rank's stream has reached it, i.e. after all stream-ordered work in front of it (peer copies,
        store kernels, scoring launches) has finished everywhere -- with no host synchronisation, so the host keeps
        queueing the next chunk while the exchange drains."""
        dist.all_reduce(self._tick, group=self.group)
