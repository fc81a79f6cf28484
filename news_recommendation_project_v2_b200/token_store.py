"""Token-state store -> packed varlen layout (the data format in front of Stage A).

Two layouts: the reference's sqlite store (read / written here for compatibility) and the packed token FILE
(`write_packed_tokens`, `convert_token_store`, `PackedTokenFile`, `apply_token_attn_packed`) that Stage A streams
from with pinned double buffering.

The reference keeps one `torch.save`d `[n_tok, d]` tensor of valid-token hidden states per news item in a
sqlite table `tensors(id INTEGER PRIMARY KEY, data BLOB)` (modeling_utils.py:456-473) and, per batch, reads
the blobs back, pads them to the batch maximum and builds a mask on the CPU
(data_utils.py:878-890, 929-933, 753-781).  Here the store is read ONCE into a packed token matrix plus CSR
offsets, which `LatentAttentionModel.forward_packed` consumes directly -- no padding, no mask.

Host-side I/O only (sqlite3 + torch.load); all arithmetic stays in the CUDA path.
"""
from __future__ import annotations

import io
import os
import sqlite3
import struct
from typing import Iterable, Optional

import numpy as np
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


# ---------------------------------------------------------------------------------------------------------------
# Packed varlen token FILE: the on-disk form of what `forward_packed` consumes.
#
#   [0, 4096)            header: magic "NRBTOK01", version, dtype code, dim, n_items, n_tokens, section offsets
#   [tokens_off, ...)    tokens  [n_tokens, dim]  bf16 (or fp32), item after item, valid tokens only
#   [offsets_off, ...)   offsets int64 [n_items + 1] (CSR: item i owns token rows offsets[i] .. offsets[i+1])
#
# Both sections are 4096-byte aligned, so the file is mmap-able as is; a chunk of items is one contiguous byte
# range.  The reference's store (one pickled tensor per sqlite row, modeling_utils.py:456-473) needs a SQL query, a
# torch.load and a pad-to-batch-max per batch (data_utils.py:878-933, 753-781); this layout needs none of them.
# ---------------------------------------------------------------------------------------------------------------
_MAGIC = b"NRBTOK01"
_HEADER = struct.Struct("<8sIIQQQQQ")  # magic, version, dtype, dim, n_items, n_tokens, tokens_off, offsets_off
_ALIGN = 4096
_DTYPES = {0: (torch.float32, np.float32, 4), 1: (torch.bfloat16, np.uint16, 2)}


def write_packed_tokens(path: str, items: Iterable[torch.Tensor], dim: int, dtype: torch.dtype = torch.bfloat16):
    """Stream `[n_tok, dim]` tensors (valid tokens of one news item each) into a packed token file.
    Returns (n_items, n_tokens)."""
    code = 1 if dtype == torch.bfloat16 else 0
    if dtype not in (torch.bfloat16, torch.float32):
        raise ValueError("packed token files hold bf16 or fp32 tokens")
    lens = []
    n_tok = 0
    with open(path, "wb") as f:
        f.write(b"\0" * _ALIGN)
        for t in items:
            if t.dim() != 2 or t.shape[1] != dim:
                raise ValueError(f"item has shape {tuple(t.shape)}, expected [n_tok, {dim}]")
            t = t.detach().to(device="cpu", dtype=dtype).contiguous()
            raw = t.view(torch.int16).numpy() if code == 1 else t.numpy()
            f.write(raw.tobytes())
            lens.append(t.shape[0])
            n_tok += t.shape[0]
        pos = f.tell()
        off_pos = (pos + _ALIGN - 1) // _ALIGN * _ALIGN
        f.write(b"\0" * (off_pos - pos))
        off = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(np.asarray(lens, dtype=np.int64), out=off[1:])
        f.write(off.tobytes())
        f.seek(0)
        f.write(_HEADER.pack(_MAGIC, 1, code, dim, len(lens), n_tok, _ALIGN, off_pos))
    return len(lens), n_tok


def convert_token_store(db_name: str, out_path: str, dtype: torch.dtype = torch.bfloat16, batch: int = 2048):
    """The reference's sqlite token store (modeling_utils.py:456-473) -> packed token file, in id order."""
    conn = sqlite3.connect(db_name)
    try:
        first = conn.execute("SELECT data FROM tensors ORDER BY id LIMIT 1;").fetchone()
        if first is None:
            raise ValueError("empty token store")
        dim = int(torch.load(io.BytesIO(first[0]), weights_only=True).shape[1])

        def rows():
            cur = conn.execute("SELECT data FROM tensors ORDER BY id;")
            while True:
                got = cur.fetchmany(batch)
                if not got:
                    return
                for (blob,) in got:
                    yield torch.load(io.BytesIO(blob), weights_only=True)

        return write_packed_tokens(out_path, rows(), dim, dtype)
    finally:
        conn.close()


def _clear_cuda_last_error() -> None:
    """cudaGetLastError() in the CUDA runtime torch itself is linked against (torch exposes no binding for it): a
    failed cudaHostRegister otherwise stays pending and is reported by the next unrelated torch call."""
    import ctypes

    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            ctypes.CDLL(name).cudaGetLastError()
            return
        except OSError:
            continue


class PackedTokenFile:
    """mmap view of a packed token file: `.offsets` int64 [n_items + 1], `.tokens_raw` [n_tokens, dim] (uint16 bit
    patterns for bf16), `.tokens(a, b)` -> torch view of token rows [a, b) in the file dtype."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            head = f.read(_HEADER.size)
        magic, version, code, dim, n_items, n_tok, tok_off, off_off = _HEADER.unpack(head)
        if magic != _MAGIC or version != 1 or code not in _DTYPES:
            raise ValueError(f"{path} is not a packed token file")
        self.path, self.dim, self.n_items, self.n_tokens = path, int(dim), int(n_items), int(n_tok)
        self.dtype, npdt, self.elem_size = _DTYPES[code]
        self._tok_off = int(tok_off)
        self.offsets = np.memmap(path, dtype=np.int64, mode="r", offset=off_off, shape=(self.n_items + 1,))
        self.tokens_raw = np.memmap(path, dtype=npdt, mode="r", offset=tok_off, shape=(self.n_tokens, self.dim)) \
            if self.n_tokens else np.zeros((0, self.dim), dtype=npdt)

    def register(self) -> bool:
        """Page-lock the mapped token section for the GPU (cudaHostRegister): chunks then cross PCIe by DMA straight
        from the page cache, with no staging copy.  Tried first on the read-only mapping (cudaHostRegisterReadOnly),
        then -- where the driver refuses read-only registrations and the file is writable -- on a shared read-write
        mapping of the same pages (nothing is ever written through it).  Returns False when both are refused (then
        `apply_token_attn_packed` stages through pinned buffers); `register_error` keeps the CUDA error codes."""
        if getattr(self, "_registered", False):
            return True
        if not self.n_tokens or not torch.cuda.is_available():
            return False
        self.register_error = []
        attempts = [(self.tokens_raw, 8)]  # cudaHostRegisterReadOnly on the existing mapping
        if os.access(self.path, os.W_OK):
            attempts.append((None, 0))  # shared read-write mapping, default flags
        for mapping, flags in attempts:
            try:
                if mapping is None:
                    mapping = np.memmap(self.path, dtype=self.tokens_raw.dtype, mode="r+", offset=self._tok_off,
                                        shape=(self.n_tokens, self.dim))
                rt = torch.cuda.cudart()
                addr, nbytes = mapping.ctypes.data, mapping.nbytes
                err = rt.cudaHostRegister(addr, nbytes, flags)
                code = int(err[0]) if isinstance(err, tuple) else int(err)
                ok = code == 0
                if ok:
                    import warnings

                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")  # read-only mapping: torch never writes through this view
                        ok = bool(torch.from_numpy(mapping[:1]).is_pinned())
                    if not ok:
                        rt.cudaHostUnregister(addr)
                        code = -1
                self.register_error.append((flags, code))
                if ok:
                    self.tokens_raw = mapping
                    self._registered = True
                    return True
            except Exception as exc:  # noqa: BLE001 -- any refusal means "stage through pinned buffers"
                self.register_error.append((flags, repr(exc)[:120]))
            _clear_cuda_last_error()  # a refused registration must not surface later as somebody else's error
        self._registered = False
        return False

    def unregister(self) -> None:
        if getattr(self, "_registered", False):
            try:
                torch.cuda.cudart().cudaHostUnregister(self.tokens_raw.ctypes.data)
            except Exception:
                pass
            self._registered = False

    def tokens(self, a: int, b: int) -> torch.Tensor:
        t = torch.from_numpy(np.array(self.tokens_raw[a:b]))  # a private, writable copy of the mapped rows
        return t.view(torch.bfloat16) if self.dtype == torch.bfloat16 else t

    def chunk_bounds(self, max_tokens: int) -> list:
        """Largest runs of whole items with at most `max_tokens` tokens: [(item0, item1), ...]."""
        off, out, i0 = self.offsets, [], 0
        while i0 < self.n_items:
            i1 = int(np.searchsorted(off, off[i0] + max_tokens, side="right")) - 1
            i1 = min(max(i1, i0 + 1), self.n_items)
            out.append((i0, i1))
            i0 = i1
        return out


def apply_token_attn_packed(model, tok_file, chunk_tokens: Optional[int] = None, copy_threads: int = 8,
                            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pooled vectors [n_items, d] (pinned CPU fp32) for every item of a packed token file, chunk by chunk with
    pinned double buffering: while the GPU pools chunk i (`nrb_latent_forward_packed`), chunk i+1 moves from the
    page cache into a pinned staging buffer (parallel memcpy) and over PCIe on a copy stream, and the vectors of
    chunk i-1 travel back on a third stream.  Drop-in for data_model_helper.py:390-413 on the packed layout."""
    from . import _lib, config, ops

    tf = tok_file if isinstance(tok_file, PackedTokenFile) else PackedTokenFile(tok_file)
    dev = _lib.require_device(None)
    fw = model.folded(None, dev)
    d = tf.dim
    if d != fw.dim:
        raise _lib.NrbError(f"token dim {d} != model dim {fw.dim}")
    chunk_tokens = int(chunk_tokens or min(config.LATENT_MAX_TOKENS, 262144))  # finer chunks: a fuller pipeline
    longest = int(np.max(np.diff(tf.offsets))) if tf.n_items else 0
    chunk_tokens = max(chunk_tokens, longest, 1)
    bounds = tf.chunk_bounds(chunk_tokens)
    npdt = tf.tokens_raw.dtype
    tdt = torch.int16 if tf.dtype == torch.bfloat16 else torch.float32
    if out is None:
        out = torch.empty(tf.n_items, d, dtype=torch.float32).pin_memory()
    with torch.cuda.device(dev):
        pinned = [torch.empty(chunk_tokens, d, dtype=tdt).pin_memory() for _ in range(2)]
        dbuf = [torch.empty(chunk_tokens, d, dtype=tdt, device=dev) for _ in range(2)]
        obuf = [None, None]
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        cur = torch.cuda.current_stream()
        h2d_done = [None, None]
        compute_done = [None, None]
        d2h_done = [None, None]
        lib = _lib.load()

        direct = bool(getattr(tf, "_registered", False))  # mapped section page-locked: DMA straight from the mapping

        def stage(k):
            i0, i1 = bounds[k]
            t0, t1 = int(tf.offsets[i0]), int(tf.offsets[i1])
            b = k & 1
            n = t1 - t0
            if direct:
                if compute_done[b] is not None:
                    s_in.wait_event(compute_done[b])
                with torch.cuda.stream(s_in):
                    if n:
                        import warnings

                        with warnings.catch_warnings():
                            warnings.simplefilter("ignore")  # read-only mapping: torch never writes through this view
                            src = torch.from_numpy(tf.tokens_raw[t0:t1])
                        dbuf[b][:n].copy_(src.view(tdt) if src.dtype != tdt else src, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(s_in)
                h2d_done[b] = ev
                return n
            if h2d_done[b] is not None:
                h2d_done[b].synchronize()  # the pinned buffer is free once its previous upload has finished
            if n:
                # page cache -> pinned staging buffer: native multi-threaded memcpy (no GIL, whole pages per thread)
                src = tf.tokens_raw[t0:t1]
                _lib.check(lib.nrb_host_copy(pinned[b].data_ptr(), src.ctypes.data, src.nbytes, copy_threads),
                           "nrb_host_copy")
            if compute_done[b] is not None:
                s_in.wait_event(compute_done[b])  # the device buffer is free once chunk k-2 has been pooled
            with torch.cuda.stream(s_in):
                dbuf[b][:n].copy_(pinned[b][:n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
            h2d_done[b] = ev
            return n

        try:
            n_next = stage(0) if bounds else 0
            for k, (i0, i1) in enumerate(bounds):
                b, n = k & 1, n_next
                cur.wait_event(h2d_done[b])
                off_local = torch.from_numpy((np.asarray(tf.offsets[i0:i1 + 1]) - int(tf.offsets[i0])).astype(np.int64))
                toks = dbuf[b][:n]
                toks = toks.view(torch.bfloat16) if tf.dtype == torch.bfloat16 else toks
                if d2h_done[b] is not None:
                    cur.wait_event(d2h_done[b])
                obuf[b] = ops.latent_forward_packed(fw, toks, off_local, max_tokens=chunk_tokens)
                ev = torch.cuda.Event()
                ev.record(cur)
                compute_done[b] = ev
                s_out.wait_event(ev)
                with torch.cuda.stream(s_out):
                    out[i0:i1].copy_(obuf[b], non_blocking=True)
                    e2 = torch.cuda.Event()
                    e2.record(s_out)
                d2h_done[b] = e2
                if k + 1 < len(bounds):
                    n_next = stage(k + 1)  # overlaps the GPU work just queued
            s_out.synchronize()
            cur.synchronize()
        finally:
            torch.cuda.current_stream().synchronize()
    return out
