"""ShardedIndex -- a DeviceIndex-shaped facade over G row shards, one process per GPU.

SURVEY.md section 8(e): "queries replicated to every GPU (broadcast) ... each GPU runs K1/K2 on its
shard ... merge ... adds route to row % G".  The reference is ONE uvicorn process
(backend/run.py:10-14) whose routes call ``collection.add/query/update/delete``; with this class
rank 0 keeps being that process: it builds an ordinary ``Collection`` on top of a ``ShardedIndex``
(ids, metadata, persistence and the filter pass stay on rank 0, unchanged), while ranks 1..G-1 sit
in ``serve()`` and execute the same device operations on their shards.

Layout: global row g lives on shard ``g % G`` at local row ``g // G`` (appends stay balanced without
any directory).  ``remove(row)`` keeps the Collection's contract (the LAST global row moves into the
hole): the last row's vector and filter bits are fetched from its shard, written over the deleted
row's slot on ITS shard (``vs_set_row_host``), and the last row's shard shrinks by one.
Every shard's kernels report TRUE global rows (``vs_set_row_map(shard, G)``: reported row =
shard + local * G), so the device-side exchange/merge orders exact score ties by global row and the
answer is identical to a single index holding all rows.

Every operation is a collective driven by rank 0: a small header goes out with
``broadcast_object_list``, array payloads with ``broadcast``; query results come back through the
``ShardedSearcher`` (fused peer exchange over NVLink, or all-gather + merge).  The class only needs
the group's backend to move tensors living on ``comm_device`` ("cuda" with NCCL, "cpu" with gloo in
the tests, where the shard index and the search functions are injected).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np


class ShardedIndex:
    def __init__(self, dim: int, dtype: str = "bf16", device: Optional[int] = None, group=None,
                 exchange: str = "p2p", mode: str = "auto", b_max: int = 1024, k_max: int = 128,
                 index_factory: Optional[Callable] = None, searcher_factory: Optional[Callable] = None,
                 comm_device: Optional[str] = None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise ValueError("at most 8 shards (one NVSwitch domain)")
        self.dim, self.dtype = int(dim), ("bf16" if dtype in ("bf16", "bfloat16") else "f32")
        self.device = int(device) if device is not None else 0
        self.comm_device = comm_device or ("cuda:%d" % self.device if dist.get_backend(group) == "nccl" else "cpu")
        if index_factory is None:
            from .index import DeviceIndex as index_factory
        self.local = index_factory(self.dim, self.dtype, self.device, 0, self.rank, self.world)   # row map (shard, G)
        if searcher_factory is None:
            from .sharded import ShardedSearcher

            def searcher_factory(ix):
                return ShardedSearcher.for_index(ix, group=group, mode=mode, exchange=exchange, b_max=b_max, k_max=k_max)
        self.searcher = searcher_factory(self.local)
        self.n = 0                                  # global row count (same on every rank)
        self._src = dist.get_global_rank(group, 0) if group is not None else 0
        self._closed = False

    # ------------------------------------------------------------------ plumbing
    def _header(self, obj=None):
        box = [obj]
        self._dist.broadcast_object_list(box, src=self._src, group=self.group)
        return box[0]

    def _bcast(self, arr: Optional[np.ndarray], shape, dtype):
        torch = self._torch
        if self.rank == 0:
            t = torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype).reshape(shape)).to(self.comm_device)
        else:
            t = torch.empty(shape, dtype=torch.from_numpy(np.zeros(0, dtype)).dtype, device=self.comm_device)
        self._dist.broadcast(t, src=self._src, group=self.group)
        return t

    def _owner(self, row: int):
        return row % self.world, row // self.world

    def _require_front(self):
        if self.rank != 0:
            raise RuntimeError("ShardedIndex operations are driven by rank 0; other ranks call serve()")
        if self._closed:
            raise RuntimeError("ShardedIndex is closed")

    # ------------------------------------------------------------------ the collective bodies (all ranks)
    def _do_add(self, rows_t):
        first = self.n
        m = rows_t.shape[0]
        sel = rows_t[(self.rank - first) % self.world::self.world]      # global row g lives on shard g % G
        if sel.shape[0] > 0:
            self.local.add(sel.contiguous() if sel.is_cuda else sel.contiguous().numpy())
        self.n += m
        return first

    def _do_query(self, q_t, k: int, require_bits):
        return self.searcher.search(q_t, k, require_bits=require_bits)

    def _do_remove(self, row: int):
        last = self.n - 1
        so, lo = self._owner(row)
        sl, ll = self._owner(last)
        moved = -1 if row == last else last
        if row != last:
            # ship the last row's vector (+ filter bits) from its shard to the deleted row's shard
            vec = self.local.get_rows(ll, 1)[0] if self.rank == sl else None
            bits = self.local.get_filter_bits(ll) if self.rank == sl else None
            box = [(vec, bits)]
            self._dist.broadcast_object_list(box, src=self._dist.get_global_rank(self.group, sl) if self.group is not None else sl,
                                             group=self.group)
            vec, bits = box[0]
            if self.rank == so:
                self.local.set_row(lo, vec)
                self.local.set_filter_bits(lo, bits)
        if self.rank == sl:
            assert len(self.local) - 1 == ll, "shard bookkeeping diverged"
            self.local.remove(ll)                   # the shard's last local row: just shrinks
        self.n = last
        return moved

    def _do_get_rows(self, first: int, n: int):
        out = np.zeros((n, self.dim), np.float32)
        for j in range(n):
            s, l = self._owner(first + j)
            if s == self.rank:
                out[j] = self.local.get_rows(l, 1)[0]
        t = self._torch.from_numpy(out).to(self.comm_device)
        self._dist.all_reduce(t, group=self.group)  # every row is non-zero on exactly one rank
        return t.cpu().numpy()

    def _do_bits(self, row: int, bits):
        s, l = self._owner(row)
        if bits is None:                            # get
            box = [self.local.get_filter_bits(l) if self.rank == s else None]
            self._dist.broadcast_object_list(box, src=self._dist.get_global_rank(self.group, s) if self.group is not None else s,
                                             group=self.group)
            return box[0]
        if self.rank == s:
            self.local.set_filter_bits(l, bits)
        return None

    def _do_bits_range(self, first: int, bits_lists):
        j0 = (self.rank - first) % self.world            # my rows are every G-th one
        mine = bits_lists[j0::self.world]
        if mine:
            self.local.set_filter_bits_range((first + j0) // self.world, mine)

    def _dispatch(self, h):
        op = h["op"]
        if op == "add":
            return self._do_add(self._bcast(None, (h["n"], self.dim), np.float32))
        if op == "query":
            return self._do_query(self._bcast(None, (h["B"], self.dim), np.float32), h["k"], h.get("bits"))
        if op == "remove":
            return self._do_remove(h["row"])
        if op == "get_rows":
            return self._do_get_rows(h["first"], h["n"])
        if op == "bits":
            return self._do_bits(h["row"], h.get("bits"))
        if op == "bits_range":
            return self._do_bits_range(h["first"], h["bits"])
        if op == "clear":
            self.local.clear()
            self.n = 0
            return None
        if op == "close":
            self._closed = True
            self.local.close()
            return None
        raise RuntimeError(f"unknown sharded op {op!r}")

    def serve(self):
        """Ranks != 0: execute rank 0's operations until it closes the index."""
        if self.rank == 0:
            raise RuntimeError("rank 0 is the front end")
        while not self._closed:
            self._dispatch(self._header())

    # ------------------------------------------------------------------ DeviceIndex surface (rank 0)
    def __len__(self):
        return self.n

    count = __len__

    def add(self, rows) -> int:
        self._require_front()
        a = rows.detach().cpu().numpy() if hasattr(rows, "detach") else np.asarray(rows)
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, self.dim)
        self._header({"op": "add", "n": int(a.shape[0])})
        return self._do_add(self._bcast(a, a.shape, np.float32))

    def query(self, q, k: int, require_bits: Optional[Sequence[int]] = None, mode: str = "auto"):
        self._require_front()
        a = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.dim)
        self._header({"op": "query", "B": int(a.shape[0]), "k": int(k), "bits": require_bits})
        s, r = self._do_query(self._bcast(a, a.shape, np.float32), int(k), require_bits)
        return s.cpu().numpy(), r.cpu().numpy().astype(np.int64)

    def query_dev(self, q, k: int, out_scores=None, out_rows=None, require_bits=None, mode: str = "auto", stream=None):
        """Device-tensor flavour of :meth:`query` (the queries are re-broadcast to the other ranks)."""
        s, r = self.query(q.detach().cpu().numpy(), k, require_bits, mode)
        return self._torch.from_numpy(s), self._torch.from_numpy(r)

    def query_multimodal(self, img, txt, w, k: int, require_bits=None, mode: str = "auto"):
        """search_multimodal's blend (backend/app/main.py:850-860), then the sharded query.  The blend
        of B vectors is done once on rank 0's host copy (same statements as the reference)."""
        img = np.ascontiguousarray(img, np.float32).reshape(-1, self.dim)
        txt = np.ascontiguousarray(txt, np.float32).reshape(-1, self.dim)
        ww = np.broadcast_to(np.asarray(w, dtype=np.float64), (img.shape[0],))
        i_n = img / np.linalg.norm(img, axis=1, keepdims=True)
        t_n = txt / np.linalg.norm(txt, axis=1, keepdims=True)
        c = np.float32(1) * ww[:, None].astype(np.float32) * i_n + (1.0 - ww)[:, None].astype(np.float32) * t_n
        c = c / np.linalg.norm(c, axis=1, keepdims=True)
        return self.query(c.astype(np.float32), k, require_bits, mode)

    def remove(self, row: int) -> int:
        self._require_front()
        if not 0 <= row < self.n:
            raise ValueError(f"row {row} out of range [0,{self.n})")
        self._header({"op": "remove", "row": int(row)})
        return self._do_remove(int(row))

    def get_rows(self, first: int, n: int) -> np.ndarray:
        self._require_front()
        self._header({"op": "get_rows", "first": int(first), "n": int(n)})
        return self._do_get_rows(int(first), int(n))

    def set_filter_bits(self, row: int, bits: Sequence[int]):
        self._require_front()
        self._header({"op": "bits", "row": int(row), "bits": list(bits)})
        self._do_bits(int(row), list(bits))

    def set_filter_bits_range(self, first: int, bits_lists):
        self._require_front()
        bits_lists = [list(b) for b in bits_lists]
        self._header({"op": "bits_range", "first": int(first), "bits": bits_lists})
        self._do_bits_range(int(first), bits_lists)

    def get_filter_bits(self, row: int):
        self._require_front()
        self._header({"op": "bits", "row": int(row)})
        return self._do_bits(int(row), None)

    def clear(self):
        self._require_front()
        self._header({"op": "clear"})
        self.local.clear()
        self.n = 0

    def close(self):
        if self.rank == 0 and not self._closed:
            self._header({"op": "close"})
            self._closed = True
            self.local.close()
