"""GroupIndex -- a DeviceIndex-shaped facade over ``vs_group_t``: ONE process, every GPU of the box.

The reference is a single uvicorn process (``backend/run.py:10-14``) whose routes call
``collection.add / query / update / delete`` one request at a time (``backend/app/main.py:735-805``).
``Collection(index=GroupIndex(...))`` keeps exactly that shape on G GPUs: no torchrun, no pickled
messages, no second process.  Global row g lives on shard ``g % G`` at local row ``g // G`` (appends stay
balanced); every kernel reports true global rows, so ties order as in a single index.

* queries go through ``vs_group_query_host`` (csrc/group.cu): pinned host-mapped request area, one worker
  thread + ONE fused scan launch per GPU, candidates exchanged over NVLink inside the kernel, result and
  completion flag written straight into host-mapped memory -- no stream synchronise on the path;
* ingest / maintenance / filter bits go to the shards (``vs_group_shard``), GPU-parallel via threads
  (ctypes drops the GIL);
* the filter sweep (config 4) needs no exchange: every shard sweeps its rows and stores the outcome in
  its own filter bits on the device;
* all-pairs dedup (config 5) replicates the striped shards into one full index per GPU with NVLink peer
  reads (``vs_replicate_from``) and splits the i<j triangle into equal-work row ranges.
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from .index import DeviceIndex, _DTYPES, _MODES, _bits_array, bits_to_words
from .sharded import triangle_bounds


class GroupIndex:
    def __init__(self, dim: int, dtype: str = "bf16", devices: Optional[Sequence[int]] = None, capacity: int = 0,
                 b_max: int = 1024, k_max: int = 128):
        self._lib = N.load()
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        if devices is None:
            import torch
            devices = list(range(max(1, torch.cuda.device_count())))
        self.devices = [int(d) for d in devices]
        self.world = len(self.devices)
        self.dim, self.dtype = int(dim), ("bf16" if _DTYPES[dtype] else "f32")
        self.b_max, self.k_max = int(b_max), int(k_max)
        h = C.c_void_p()
        dev_arr = (C.c_int * self.world)(*self.devices)
        N.check(self._lib.vs_group_create(self.world, dev_arr, self.dim, _DTYPES[dtype], int(capacity), self.b_max,
                                          self.k_max, C.byref(h)))
        self._h = h
        self.shards: List[DeviceIndex] = [
            DeviceIndex(self.dim, self.dtype, _handle=self._lib.vs_group_shard(self._h, s)) for s in range(self.world)]
        self.device = self.devices[0]
        self._pool = ThreadPoolExecutor(max_workers=self.world) if self.world > 1 else None
        self._full: Optional[List[DeviceIndex]] = None     # per-GPU replicas for the all-pairs pass

    # ------------------------------------------------------------------ plumbing
    def _each(self, fn):
        """fn(shard_number, shard) on every shard, in parallel across GPUs."""
        if self._pool is None:
            return [fn(0, self.shards[0])]
        return list(self._pool.map(lambda s: fn(s, self.shards[s]), range(self.world)))

    def _owner(self, row: int) -> Tuple[int, int]:
        return row % self.world, row // self.world

    def _local_count(self, n: int, s: int) -> int:
        return (n - s + self.world - 1) // self.world if n > s else 0

    def close(self):
        if getattr(self, "_h", None):
            self._drop_replicas()
            if self._pool is not None:
                self._pool.shutdown(wait=True)
            for sh in self.shards:
                sh.close()                              # non-owning: just forgets the handle
            self._lib.vs_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self._lib.vs_group_count(self._h))

    count = __len__

    @property
    def last_query_path(self) -> str:
        return self.shards[0].last_query_path

    # ------------------------------------------------------------------ ingest / maintenance
    def add(self, rows) -> int:
        """Append rows ([n, dim] float32 numpy or CUDA tensor); row ``first + j`` goes to shard ``(first + j) % G``."""
        first = len(self)
        self._drop_replicas()
        if hasattr(rows, "is_cuda") and rows.is_cuda:
            import torch
            t = rows.detach().to(torch.float32)

            def put(s, sh):
                sel = t[(s - first) % self.world::self.world]
                if sel.shape[0]:
                    with torch.cuda.device(sh.device):
                        sh.add(sel.to(torch.device("cuda", sh.device)).contiguous())
                        torch.cuda.current_stream().synchronize()
            self._each(put)
            return first
        a = np.asarray(rows.detach().cpu().numpy() if hasattr(rows, "detach") else rows, dtype=np.float32)
        a = a.reshape(-1, self.dim)

        def put(s, sh):
            sel = a[(s - first) % self.world::self.world]
            if sel.shape[0]:
                sh.add(np.ascontiguousarray(sel))
        self._each(put)
        return first

    def add_raw(self, stored_rows: np.ndarray) -> int:
        """Append rows already in the storage dtype (persistence slab reload), striped over the shards."""
        first = len(self)
        self._drop_replicas()

        def put(s, sh):
            sel = stored_rows[(s - first) % self.world::self.world]
            if sel.shape[0]:
                sh.add_raw(np.ascontiguousarray(sel))
        self._each(put)
        return first

    def get_raw(self, first: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), dtype=self.shards[0].storage_dtype)

        def get(s, sh):
            j0 = (s - first) % self.world
            m = len(range(j0, n, self.world))
            if m:
                out[j0::self.world] = sh.get_raw((first + j0) // self.world, m)
        self._each(get)
        return out

    @property
    def storage_dtype(self):
        return self.shards[0].storage_dtype

    def reserve(self, capacity: int):
        self._each(lambda s, sh: sh.reserve((capacity + self.world - 1) // self.world))

    def remove(self, row: int) -> int:
        """Collection contract: the LAST global row moves into the hole.  Returns the moved row or -1."""
        n = len(self)
        if not 0 <= row < n:
            raise ValueError(f"row {row} out of range [0,{n})")
        self._drop_replicas()
        last = n - 1
        so, lo = self._owner(row)
        sl, ll = self._owner(last)
        if row != last:
            self.shards[so].copy_row_from(lo, self.shards[sl], ll)       # device to device (NVLink when so != sl)
        self.shards[sl].truncate(ll)
        return -1 if row == last else last

    def remove_rows(self, rows) -> Tuple[np.ndarray, np.ndarray]:
        """Bulk delete: global compaction plan on the host, at most G*G batched row-move kernels, one
        truncate per shard.  Returns (moved_src, moved_dst) in global rows."""
        n = len(self)
        d = np.unique(np.asarray(rows, dtype=np.int64).reshape(-1))
        if d.shape[0] and (d[0] < 0 or d[-1] >= n):
            raise ValueError(f"row out of range [0,{n})")
        self._drop_replicas()
        new_n = n - int(d.shape[0])
        holes = d[d < new_n]
        tail = np.arange(new_n, n, dtype=np.int64)
        fillers = np.setdiff1d(tail, d[d >= new_n], assume_unique=True)
        assert holes.shape == fillers.shape
        G = self.world
        if holes.shape[0]:
            jobs = {}
            ss, ds = fillers % G, holes % G
            for a in range(G):
                for b in range(G):
                    m = (ss == a) & (ds == b)
                    if m.any():
                        jobs.setdefault(b, []).append((a, fillers[m] // G, holes[m] // G))

            def run(s, sh):
                for a, src_l, dst_l in jobs.get(s, []):
                    sh.move_rows_from(self.shards[a], src_l, dst_l)
            self._each(run)       # every destination row is written by exactly one job; sources are never written
        self._each(lambda s, sh: sh.truncate(self._local_count(new_n, s)))
        return fillers, holes

    def clear(self):
        self._drop_replicas()
        self._each(lambda s, sh: sh.truncate(0))

    def get_rows(self, first: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), dtype=np.float32)

        def get(s, sh):
            j0 = (s - first) % self.world
            m = len(range(j0, n, self.world))
            if m:
                out[j0::self.world] = sh.get_rows((first + j0) // self.world, m)
        self._each(get)
        return out

    # ------------------------------------------------------------------ filter bits
    def set_filter_bits(self, row: int, bits: Sequence[int]):
        s, l = self._owner(row)
        self.shards[s].set_filter_bits(l, bits)

    def get_filter_bits(self, row: int):
        s, l = self._owner(row)
        return self.shards[s].get_filter_bits(l)

    def set_filter_bits_range(self, first: int, bits_lists):
        self.set_filter_words_range(first, bits_to_words(bits_lists))

    def set_filter_words_range(self, first: int, words: np.ndarray):
        w = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1, N.MASK_WORDS)

        def put(s, sh):
            j0 = (s - first) % self.world
            sel = w[j0::self.world]
            if sel.shape[0]:
                sh.set_filter_words_range((first + j0) // self.world, np.ascontiguousarray(sel))
        self._each(put)

    def get_filter_words_range(self, first: int, n: int) -> np.ndarray:
        out = np.zeros((n, N.MASK_WORDS), dtype=np.uint64)

        def get(s, sh):
            j0 = (s - first) % self.world
            m = len(range(j0, n, self.world))
            if m:
                out[j0::self.world] = sh.get_filter_words_range((first + j0) // self.world, m)
        self._each(get)
        return out

    # ------------------------------------------------------------------ query
    def _out(self, B: int, k: int):
        return np.empty((B, k), dtype=np.float32), np.empty((B, k), dtype=np.int64)

    def query(self, q, k: int, require_bits: Optional[Sequence[int]] = None, mode: str = "auto"):
        """HOST-buffer query through the group's request/response path (vs_group_query_host)."""
        a = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, self.dim)
        B = a.shape[0]
        s, r = self._out(B, k)
        N.check(self._lib.vs_group_query_host(self._h, a.ctypes.data, B, int(k), _bits_array(require_bits), _MODES[mode],
                                              s.ctypes.data, r.ctypes.data))
        return s, r

    def query_dev(self, q, k: int, out_scores=None, out_rows=None, require_bits=None, mode: str = "auto", stream=None):
        """Device-tensor flavour: the query has to reach every GPU anyway, so it travels through the pinned
        request area; returns CPU tensors."""
        import torch
        s, r = self.query(q.detach().cpu().numpy(), k, require_bits, mode)
        return torch.from_numpy(s), torch.from_numpy(r)

    def query_multimodal(self, img, txt, w, k: int, require_bits=None, mode: str = "auto"):
        """search_multimodal (backend/app/main.py:829-867): every GPU blends its own copy of the (img, txt, w)
        triples with the blend kernel, then the same fused query path."""
        a = np.ascontiguousarray(img, dtype=np.float32).reshape(-1, self.dim)
        t = np.ascontiguousarray(txt, dtype=np.float32).reshape(-1, self.dim)
        if a.shape != t.shape:
            raise ValueError("img/txt must both be [B, dim]")
        B = a.shape[0]
        ww = np.ascontiguousarray(np.broadcast_to(np.asarray(w, dtype=np.float64), (B,)))
        s, r = self._out(B, k)
        N.check(self._lib.vs_group_query_multimodal_host(self._h, a.ctypes.data, t.ctypes.data, ww.ctypes.data, B, int(k),
                                                         _bits_array(require_bits), _MODES[mode], s.ctypes.data,
                                                         r.ctypes.data))
        return s, r

    def exchange_error(self) -> int:
        return max(sh.exchange_error() for sh in self.shards)

    def last_timing_us(self):
        """Host-side timeline of the last query: (published, all launches enqueued, completion seen, returned) in us."""
        t = (C.c_double * 4)()
        N.check(self._lib.vs_group_last_timing(self._h, t))
        return tuple(t)

    # ------------------------------------------------------------------ filter sweep (config 4)
    def filter_words(self) -> int:
        return (len(self) + 255) // 256 * 8

    def filter_sweep(self, prompts, tau: float) -> np.ndarray:
        """prompts [F, dim] -> uint32 bit mask [F, filter_words()] over GLOBAL rows: every shard sweeps its
        own rows (no exchange), the per-shard bit rows are interleaved on the host."""
        p = np.ascontiguousarray(prompts, dtype=np.float32).reshape(-1, self.dim)
        n, F = len(self), p.shape[0]
        parts = self._each(lambda s, sh: sh.filter_sweep(p, tau) if len(sh) else None)
        glob = np.zeros((F, self.filter_words() * 32), dtype=np.uint8)
        for s, bits in enumerate(parts):
            if bits is None:
                continue
            ns = self._local_count(n, s)
            glob[:, s:n:self.world] = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :ns]
        return np.packbits(glob, axis=1, bitorder="little").view(np.uint32)

    def apply_filter_sweep(self, prompt, tau: float, bit: int) -> int:
        """One prompt -> filter bit ``bit`` of every row, on every GPU at once; returns the number of passing rows."""
        return int(sum(self._each(lambda s, sh: sh.apply_filter_sweep(prompt, tau, bit))))

    # ------------------------------------------------------------------ all-pairs dedup (config 5)
    def _drop_replicas(self):
        if self._full:
            for f in self._full:
                f.close()
        self._full = None

    def replicate(self) -> List[DeviceIndex]:
        """One FULL copy of the corpus per GPU (global row order), pulled from the shards over NVLink."""
        if self._full is None:
            n = len(self)

            def build(s, sh):
                full = DeviceIndex(self.dim, self.dtype, device=sh.device, capacity=n)
                for src_s, src in enumerate(self.shards):
                    full.replicate_from(src, src_s, self.world)
                assert len(full) == n
                return full
            self._full = self._each(build)
        return self._full

    def dedup(self, tau: float, row_lo: int = 0, row_hi: Optional[int] = None, capacity: int = 1 << 20):
        """All pairs (i<j) with cos >= tau, i in [row_lo, row_hi): (i, j, score) sorted by (i, j).  GPU s works
        on its equal-WORK slice of the triangle (``triangle_bounds``) against its full replica."""
        n = len(self)
        row_hi = n if row_hi is None else min(int(row_hi), n)
        if n < 2 or row_hi <= row_lo:
            return np.empty(0, np.int64), np.empty(0, np.int64), np.empty(0, np.float32)
        full = self.replicate()

        def run(s, _sh):
            lo, hi = triangle_bounds(n, self.world, s)
            lo, hi = max(lo, row_lo), min(hi, row_hi)
            if hi <= lo:
                return np.empty(0, np.int64), np.empty(0, np.int64), np.empty(0, np.float32)
            return full[s].dedup(tau, lo, hi, capacity=max(1024, capacity // self.world))
        parts = self._each(run)
        i = np.concatenate([p[0] for p in parts])
        j = np.concatenate([p[1] for p in parts])
        sc = np.concatenate([p[2] for p in parts])
        order = np.lexsort((j, i))
        return i[order], j[order], sc[order]
