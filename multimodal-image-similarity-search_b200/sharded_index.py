"""ShardedIndex -- a DeviceIndex-shaped facade over G row shards, ONE PROCESS PER GPU (torchrun).

SURVEY.md section 8(e): "queries replicated to every GPU (broadcast) ... each GPU runs K1/K2 on its
shard ... merge ... adds route to row % G".  With this class rank 0 is the reference's server process
(backend/run.py:10-14): it builds an ordinary ``Collection`` on top of a ``ShardedIndex`` (ids, metadata,
persistence and the filter pass stay on rank 0, unchanged), while ranks 1..G-1 sit in ``serve()`` and
execute the same device operations on their shards.  (``GroupIndex`` is the same collection in a single
process; use this class when the deployment is already one process per GPU.)

Layout: global row g lives on shard ``g % G`` at local row ``g // G`` (appends stay balanced without
any directory).  ``remove(row)`` keeps the Collection's contract (the LAST global row moves into the
hole).  Every shard's kernels report TRUE global rows (``vs_set_row_map(shard, G)``), so the device-side
exchange/merge orders exact score ties by global row and the answer is identical to a single index.

Every operation is a collective driven by rank 0.  Nothing is pickled: an operation is announced with
ONE small int64 tensor (op code + scalars + the four filter-bit words) and its payloads (rows, queries,
bit words) travel as tensors through ``broadcast``; query results come back through the
``ShardedSearcher`` (fused peer exchange over NVLink, or all-gather + merge).  The multimodal blend, the
filter sweep and the all-pairs pass run on every rank's device.  The class only needs the group's backend
to move tensors living on ``comm_device`` ("cuda" with NCCL, "cpu" with gloo in the tests, where the shard
index and the search functions are injected).
"""
from __future__ import annotations

import struct
from typing import Callable, Optional, Sequence

import numpy as np

from . import _native as N
from .index import bits_to_words
from .sharded import triangle_bounds

_OPS = ["add", "query", "query_mm", "remove", "get_rows", "bits_get", "bits_set", "bits_range", "clear", "close",
        "sweep", "sweep_apply", "dedup"]
_OP = {name: i + 1 for i, name in enumerate(_OPS)}


def _f2i(x: float) -> int:
    return struct.unpack("<q", struct.pack("<d", float(x)))[0]


def _i2f(i: int) -> float:
    return struct.unpack("<d", struct.pack("<q", int(i)))[0]


def _words_of(bits: Optional[Sequence[int]]) -> np.ndarray:
    return bits_to_words([list(bits)])[0] if bits else np.zeros(N.MASK_WORDS, dtype=np.uint64)


def _bits_of(words: np.ndarray) -> list:
    return [b for b in range(64 * N.MASK_WORDS) if (int(words[b // 64]) >> (b % 64)) & 1]


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
        self._index_factory = index_factory
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
    def _grank(self, r: int) -> int:
        return self._dist.get_global_rank(self.group, r) if self.group is not None else r

    def _header(self, op: Optional[str] = None, a: int = 0, b: int = 0, c: int = 0, words: Optional[np.ndarray] = None):
        """rank 0 announces an operation; every rank returns (op, a, b, c, words[4] uint64)."""
        torch = self._torch
        h = torch.zeros(8, dtype=torch.int64)
        if self.rank == 0:
            w = np.zeros(N.MASK_WORDS, dtype=np.uint64) if words is None else np.asarray(words, dtype=np.uint64)
            h = torch.tensor([_OP[op], int(a), int(b), int(c), *w.view(np.int64).tolist()], dtype=torch.int64)
        h = h.to(self.comm_device)
        self._dist.broadcast(h, src=self._src, group=self.group)
        v = h.cpu().numpy()
        return _OPS[int(v[0]) - 1], int(v[1]), int(v[2]), int(v[3]), v[4:8].copy().view(np.uint64)

    def _bcast(self, arr, shape, dtype, src: int = 0):
        torch = self._torch
        if self.rank == src and arr is not None:
            t = torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype).reshape(shape)).to(self.comm_device)
        else:
            t = torch.empty(shape, dtype=torch.from_numpy(np.zeros(0, dtype)).dtype, device=self.comm_device)
        self._dist.broadcast(t, src=self._grank(src), group=self.group)
        return t

    def _owner(self, row: int):
        return row % self.world, row // self.world

    def _local_count(self, n: int, s: int) -> int:
        return (n - s + self.world - 1) // self.world if n > s else 0

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

    def _do_query(self, q_t, k: int, words):
        bits = _bits_of(words) if words.any() else None
        return self.searcher.search(q_t, k, require_bits=bits)

    def _do_query_mm(self, B: int, k: int, words, img=None, txt=None, w=None):
        """search_multimodal's blend (backend/app/main.py:850-860) on EVERY rank's device, then the sharded query."""
        img_t = self._bcast(img, (B, self.dim), np.float32)
        txt_t = self._bcast(txt, (B, self.dim), np.float32)
        w_t = self._bcast(w, (B,), np.float64)
        return self._do_query(self.local.blend_dev(img_t, txt_t, w_t), k, words)

    def _do_remove(self, row: int):
        last = self.n - 1
        so, lo = self._owner(row)
        sl, ll = self._owner(last)
        moved = -1 if row == last else last
        if row != last:
            # ship the last row's vector (+ filter-bit words) from its shard to the deleted row's shard
            vec = self.local.get_rows(ll, 1)[0] if self.rank == sl else None
            wrd = _words_of(self.local.get_filter_bits(ll)).view(np.int64) if self.rank == sl else None
            vec_t = self._bcast(vec, (self.dim,), np.float32, src=sl)
            wrd_t = self._bcast(wrd, (N.MASK_WORDS,), np.int64, src=sl)
            if self.rank == so:
                self.local.set_row(lo, vec_t.cpu().numpy())
                self.local.set_filter_bits(lo, _bits_of(wrd_t.cpu().numpy().view(np.uint64)))
        if self.rank == sl:
            assert len(self.local) - 1 == ll, "shard bookkeeping diverged"
            self.local.remove(ll)                   # the shard's last local row: just shrinks
        self.n = last
        return moved

    def _do_get_rows(self, first: int, n: int):
        out = np.zeros((n, self.dim), np.float32)
        j0 = (self.rank - first) % self.world
        m = len(range(j0, n, self.world))
        if m:
            out[j0::self.world] = self.local.get_rows((first + j0) // self.world, m)
        t = self._torch.from_numpy(out).to(self.comm_device)
        self._dist.all_reduce(t, group=self.group)  # every row is non-zero on exactly one rank
        return t.cpu().numpy()

    def _do_bits_get(self, row: int):
        s, l = self._owner(row)
        wrd = _words_of(self.local.get_filter_bits(l)).view(np.int64) if self.rank == s else None
        return _bits_of(self._bcast(wrd, (N.MASK_WORDS,), np.int64, src=s).cpu().numpy().view(np.uint64))

    def _do_bits_set(self, row: int, words):
        s, l = self._owner(row)
        if self.rank == s:
            self.local.set_filter_bits(l, _bits_of(words))

    def _do_bits_range(self, first: int, n: int, words=None):
        w = self._bcast(None if words is None else np.asarray(words, np.uint64).view(np.int64), (n, N.MASK_WORDS), np.int64)
        w = w.cpu().numpy().view(np.uint64)
        j0 = (self.rank - first) % self.world            # my rows are every G-th one
        mine = w[j0::self.world]
        if mine.shape[0]:
            if hasattr(self.local, "set_filter_words_range"):
                self.local.set_filter_words_range((first + j0) // self.world, np.ascontiguousarray(mine))
            else:
                self.local.set_filter_bits_range((first + j0) // self.world, [_bits_of(x) for x in mine])

    def _do_sweep(self, F: int, tau: float, prompts=None):
        """config 4: every shard sweeps its rows (no exchange); the bit rows are gathered and interleaved on rank 0."""
        torch = self._torch
        p = self._bcast(prompts, (F, self.dim), np.float32).cpu().numpy()
        wmax = (self._local_count(self.n, 0) + 255) // 256 * 8
        mine = np.zeros((F, max(wmax, 1)), dtype=np.uint32)
        if len(self.local):
            b = self.local.filter_sweep(p, tau)
            mine[:, :b.shape[1]] = b
        t = torch.from_numpy(mine.view(np.int32)).to(self.comm_device)
        out = torch.empty((self.world * F, mine.shape[1]), dtype=torch.int32, device=self.comm_device)
        self._dist.all_gather_into_tensor(out, t, group=self.group)
        if self.rank != 0:
            return None
        parts = out.cpu().numpy().view(np.uint32).reshape(self.world, F, -1)
        glob = np.zeros((F, self.filter_words() * 32), dtype=np.uint8)
        for s in range(self.world):
            ns = self._local_count(self.n, s)
            if ns:
                glob[:, s:self.n:self.world] = np.unpackbits(parts[s].view(np.uint8), axis=1, bitorder="little")[:, :ns]
        return np.packbits(glob, axis=1, bitorder="little").view(np.uint32)

    def _do_sweep_apply(self, bit: int, tau: float, prompt=None):
        torch = self._torch
        p = self._bcast(prompt, (self.dim,), np.float32).cpu().numpy()
        cnt = int(self.local.apply_filter_sweep(p, tau, bit)) if len(self.local) else 0
        t = torch.tensor([cnt], dtype=torch.int64, device=self.comm_device)
        self._dist.all_reduce(t, group=self.group)
        return int(t.item())

    def _do_dedup(self, tau: float, capacity: int):
        """config 5: every rank assembles a full replica (shard-major row order), works on its equal-work slice of the
        triangle, and the pairs -- mapped back to global rows -- are gathered on rank 0."""
        torch = self._torch
        counts = [self._local_count(self.n, s) for s in range(self.world)]
        starts = np.concatenate([[0], np.cumsum(counts)])
        full = self._index_factory(self.dim, self.dtype, self.device, self.n)
        chunk = 1 << 16
        for s in range(self.world):
            for c0 in range(0, counts[s], chunk):
                m = min(chunk, counts[s] - c0)
                if self.rank == s and hasattr(self.local, "get_rows_dev") and self.comm_device != "cpu":
                    view = self.local.get_rows_dev(c0, m)
                    self._dist.broadcast(view, src=self._grank(s), group=self.group)
                elif self.comm_device != "cpu":
                    view = torch.empty((m, self.dim), dtype=torch.float32, device=self.comm_device)
                    self._dist.broadcast(view, src=self._grank(s), group=self.group)
                else:
                    view = self._bcast(self.local.get_rows(c0, m) if self.rank == s else None, (m, self.dim), np.float32, src=s)
                full.add(view if view.is_cuda else view.numpy())
        lo, hi = triangle_bounds(self.n, self.world, self.rank)
        if hi > lo:
            i, j, sc = full.dedup(tau, lo, hi, capacity=capacity)
        else:
            i, j, sc = np.empty(0, np.int64), np.empty(0, np.int64), np.empty(0, np.float32)
        full.close()

        def to_global(p):                           # replica row -> (shard, local) -> global row
            s = np.searchsorted(starts, p, side="right") - 1
            return (p - starts[s]) * self.world + s
        gi, gj = to_global(np.asarray(i, np.int64)), to_global(np.asarray(j, np.int64))
        a, b = np.minimum(gi, gj), np.maximum(gi, gj)
        cnt = torch.tensor([len(a)], dtype=torch.int64, device=self.comm_device)
        cnts = torch.empty(self.world, dtype=torch.int64, device=self.comm_device)
        self._dist.all_gather_into_tensor(cnts, cnt, group=self.group)
        mx = max(1, int(cnts.max().item()))
        pad = np.zeros((3, mx), dtype=np.float64)
        pad[0, :len(a)], pad[1, :len(a)], pad[2, :len(a)] = a, b, sc
        out = torch.empty((self.world * 3, mx), dtype=torch.float64, device=self.comm_device)
        self._dist.all_gather_into_tensor(out, torch.from_numpy(pad).to(self.comm_device), group=self.group)
        if self.rank != 0:
            return None
        o = out.cpu().numpy().reshape(self.world, 3, mx)
        c = cnts.cpu().numpy()
        i = np.concatenate([o[s, 0, :c[s]] for s in range(self.world)]).astype(np.int64)
        j = np.concatenate([o[s, 1, :c[s]] for s in range(self.world)]).astype(np.int64)
        sc = np.concatenate([o[s, 2, :c[s]] for s in range(self.world)]).astype(np.float32)
        order = np.lexsort((j, i))
        return i[order], j[order], sc[order]

    def _dispatch(self, hdr):
        op, a, b, c, words = hdr
        if op == "add":
            return self._do_add(self._bcast(None, (a, self.dim), np.float32))
        if op == "query":
            return self._do_query(self._bcast(None, (a, self.dim), np.float32), b, words)
        if op == "query_mm":
            return self._do_query_mm(a, b, words)
        if op == "remove":
            return self._do_remove(a)
        if op == "get_rows":
            return self._do_get_rows(a, b)
        if op == "bits_get":
            return self._do_bits_get(a)
        if op == "bits_set":
            return self._do_bits_set(a, words)
        if op == "bits_range":
            return self._do_bits_range(a, b)
        if op == "sweep":
            return self._do_sweep(a, _i2f(b))
        if op == "sweep_apply":
            return self._do_sweep_apply(a, _i2f(b))
        if op == "dedup":
            return self._do_dedup(_i2f(a), b)
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
        self._header("add", a.shape[0])
        return self._do_add(self._bcast(a, a.shape, np.float32))

    def _finish(self, s, r):
        s, r = s.cpu().numpy(), r.cpu().numpy().astype(np.int64)      # (synchronises)
        if hasattr(self.local, "exchange_error") and self.local.exchange_error():
            self.local.exchange_clear_error()
            raise RuntimeError("peer exchange timed out: a shard never delivered its candidates (results were empty)")
        return s, r

    def query(self, q, k: int, require_bits: Optional[Sequence[int]] = None, mode: str = "auto"):
        self._require_front()
        a = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.dim)
        words = _words_of(require_bits)
        self._header("query", a.shape[0], int(k), words=words)
        return self._finish(*self._do_query(self._bcast(a, a.shape, np.float32), int(k), words))

    def query_dev(self, q, k: int, out_scores=None, out_rows=None, require_bits=None, mode: str = "auto", stream=None):
        """Device-tensor flavour of :meth:`query`: a CUDA tensor on the communication device is broadcast to the other
        ranks as it is (NCCL, device to device) and the result stays on the device -- nothing bounces through the host.
        (A peer-exchange time-out then surfaces on the next call: the sticky error word, see vs_exchange_error.)"""
        self._require_front()
        torch = self._torch
        t = q.detach()
        if not t.is_cuda or self.comm_device == "cpu":
            s, r = self.query(t.cpu().numpy(), k, require_bits, mode)
            return torch.from_numpy(s), torch.from_numpy(r)
        t = t.to(torch.float32).reshape(-1, self.dim).contiguous().to(self.comm_device)
        words = _words_of(require_bits)
        self._header("query", t.shape[0], int(k), words=words)
        self._dist.broadcast(t, src=self._src, group=self.group)
        return self._do_query(t, int(k), words)

    def query_multimodal(self, img, txt, w, k: int, require_bits=None, mode: str = "auto"):
        self._require_front()
        img = np.ascontiguousarray(img, np.float32).reshape(-1, self.dim)
        txt = np.ascontiguousarray(txt, np.float32).reshape(-1, self.dim)
        ww = np.ascontiguousarray(np.broadcast_to(np.asarray(w, dtype=np.float64), (img.shape[0],)))
        words = _words_of(require_bits)
        self._header("query_mm", img.shape[0], int(k), words=words)
        return self._finish(*self._do_query_mm(img.shape[0], int(k), words, img, txt, ww))

    def remove(self, row: int) -> int:
        self._require_front()
        if not 0 <= row < self.n:
            raise ValueError(f"row {row} out of range [0,{self.n})")
        self._header("remove", int(row))
        return self._do_remove(int(row))

    def get_rows(self, first: int, n: int) -> np.ndarray:
        self._require_front()
        self._header("get_rows", int(first), int(n))
        return self._do_get_rows(int(first), int(n))

    def set_filter_bits(self, row: int, bits: Sequence[int]):
        self._require_front()
        words = _words_of(bits)
        self._header("bits_set", int(row), words=words)
        self._do_bits_set(int(row), words)

    def set_filter_bits_range(self, first: int, bits_lists):
        self._require_front()
        words = bits_to_words([list(b) for b in bits_lists])
        if words.shape[0]:
            self._header("bits_range", int(first), int(words.shape[0]))
            self._do_bits_range(int(first), int(words.shape[0]), words)

    def get_filter_bits(self, row: int):
        self._require_front()
        self._header("bits_get", int(row))
        return self._do_bits_get(int(row))

    def filter_words(self) -> int:
        return (self.n + 255) // 256 * 8

    def filter_sweep(self, prompts, tau: float) -> np.ndarray:
        """prompts [F, dim] -> uint32 bit mask [F, filter_words()] over GLOBAL rows (BASELINE config 4)."""
        self._require_front()
        p = np.ascontiguousarray(prompts, dtype=np.float32).reshape(-1, self.dim)
        self._header("sweep", p.shape[0], _f2i(tau))
        return self._do_sweep(p.shape[0], float(tau), p)

    def apply_filter_sweep(self, prompt, tau: float, bit: int) -> int:
        self._require_front()
        p = np.ascontiguousarray(prompt, dtype=np.float32).reshape(self.dim)
        self._header("sweep_apply", int(bit), _f2i(tau))
        return self._do_sweep_apply(int(bit), float(tau), p)

    def dedup(self, tau: float, row_lo: int = 0, row_hi: Optional[int] = None, capacity: int = 1 << 20):
        """All pairs (i<j) with cos >= tau over the whole sharded collection (BASELINE config 5)."""
        self._require_front()
        self._header("dedup", _f2i(tau), int(capacity))
        i, j, s = self._do_dedup(float(tau), int(capacity))
        hi = self.n if row_hi is None else row_hi
        keep = (i >= row_lo) & (i < hi)
        return i[keep], j[keep], s[keep]

    def clear(self):
        self._require_front()
        self._header("clear")
        self.local.clear()
        self.n = 0

    def close(self):
        if self.rank == 0 and not self._closed:
            self._header("close")
            self._closed = True
            self.local.close()
