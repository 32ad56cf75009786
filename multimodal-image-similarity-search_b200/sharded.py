"""Row-sharded search across the GPUs of one box: one process per GPU (torch.distributed).

SURVEY.md section 8(e): the corpus is split into contiguous row ranges, every rank scans its
shard for the (replicated) queries, the per-rank ``[B, k]`` candidates are exchanged and merged on
every rank, so all ranks end with the same global answer.  Two exchanges:

* ``exchange="nccl"``  -- one small all-gather pair (``B*k*12`` bytes per rank) + the K5 merge kernel;
* ``exchange="p2p"``   -- ONE fused kernel (``vs_exchange_merge_dev``): each rank stores its candidates
  straight into every peer's symmetric-memory buffer over NVLink/NVSwitch (P2P stores), signals
  per-query flags, waits for the peers' flags and merges.  No NCCL call on the query path.

The searcher is backend-agnostic on purpose: ``local_topk`` and ``merge`` are injected, so the
partitioning + exchange logic is testable with world_size-2 ``gloo`` on CPU (tests inject the
oracle there); the product wiring (``ShardedSearcher.for_index``) binds them to the CUDA kernels.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_bounds(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous split ``[g*ceil(N/G), (g+1)*ceil(N/G))`` clipped to N."""
    per = (n_rows + world_size - 1) // world_size
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def triangle_bounds(n_rows: int, world_size: int, rank: int, align: int = 128) -> Tuple[int, int]:
    """Row range ``[lo, hi)`` of rank ``rank`` for the all-pairs pass (pairs i<j): row i owns the
    ``n-1-i`` pairs to its right, so equal WORK means boundaries at ``n*(1-sqrt(1-r/G))`` (not equal
    row counts); rounded to ``align`` rows (the kernel's 128-row A blocks)."""
    def cut(r):
        if r <= 0:
            return 0
        if r >= world_size:
            return n_rows
        x = n_rows * (1.0 - (1.0 - r / world_size) ** 0.5)
        return min(n_rows, int(round(x / align)) * align)
    return cut(rank), cut(rank + 1)


def replicate_index(local_index, n_total: int, group=None, chunk_rows: int = 1 << 16):
    """All-gather the row shards into a NEW full DeviceIndex on every rank (the all-pairs pass needs
    every column on every GPU): each rank broadcasts its stored rows in chunks
    (device to device over NCCL) and every rank appends them in global row order."""
    import torch
    import torch.distributed as dist
    from .index import DeviceIndex
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = torch.device("cuda", local_index.device)
    full = DeviceIndex(local_index.dim, local_index.dtype, device=local_index.device, capacity=n_total)
    buf = torch.empty((chunk_rows, local_index.dim), dtype=torch.float32, device=dev)
    for src in range(world):
        lo, hi = shard_bounds(n_total, world, src)
        for c0 in range(lo, hi, chunk_rows):
            m = min(chunk_rows, hi - c0)
            view = buf[:m]
            if rank == src:
                local_index.get_rows_dev(c0 - lo, m, out=view)
            dist.broadcast(view, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
            full.add(view)
    assert len(full) == n_total
    return full


def find_duplicates_sharded(local_dedup: Callable, n_total: int, group=None):
    """All-pairs duplicate detection split over the ranks: rank r runs
    ``local_dedup(row_lo, row_hi) -> (i, j, score)`` (numpy, global rows, pairs i<j with i in the
    range) on its triangle slice (``triangle_bounds``); the pair lists are gathered on every rank and
    returned sorted by (i, j).  ``local_dedup`` is ``full_index.dedup(tau, lo, hi)`` in production."""
    import numpy as np
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = triangle_bounds(n_total, world, rank)
    i, j, s = local_dedup(lo, hi) if hi > lo else (np.empty(0, np.int64), np.empty(0, np.int64), np.empty(0, np.float32))
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (np.asarray(i), np.asarray(j), np.asarray(s)), group=group)
        i = np.concatenate([p[0] for p in parts])
        j = np.concatenate([p[1] for p in parts])
        s = np.concatenate([p[2] for p in parts])
    order = np.lexsort((j, i))
    return i[order], j[order], s[order]


class PeerExchange:
    """Exchange buffers of ``vs_query_topk_sharded_dev`` for one process-per-GPU group: every rank
    allocates its buffer in the library, the CUDA IPC handles are all-gathered once at set-up
    (host side), and each rank maps its peers' buffers.  Nothing but kernels runs per query."""

    def __init__(self, index, group=None, b_max: int = 1024, k_max: int = 32):
        import torch.distributed as dist
        self.index, self.b_max, self.k_max = index, int(b_max), int(k_max)
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world_size > 8:
            raise ValueError("the peer exchange supports at most 8 ranks (one NVSwitch domain)")
        # Set-up is failure-atomic across ranks: every rank always takes part in both object
        # all-gathers, so a rank that cannot create or map a buffer makes ALL ranks raise together
        # (the caller can then fall back to exchange="nccl" on every rank) instead of hanging its peers.
        mine, err = None, None
        try:
            mine = index.exchange_create(self.world_size, self.rank, self.b_max, self.k_max)
        except Exception as e:                        # noqa: BLE001
            err = f"{type(e).__name__}: {e}"
        handles = [None] * self.world_size
        dist.all_gather_object(handles, mine, group=group)       # also orders "every buffer is zeroed"
        if all(h is not None for h in handles):
            try:
                index.exchange_attach(ipc_handles=handles)
            except Exception as e:                    # noqa: BLE001
                err = f"{type(e).__name__}: {e}"
        errs = [None] * self.world_size
        dist.all_gather_object(errs, err, group=group)           # every rank has mapped every buffer (or reports why not)
        bad = {r: e for r, e in enumerate(errs) if e is not None}
        if bad:
            raise RuntimeError(f"peer exchange unavailable: {bad}")

    def search(self, q, k: int, mode: str = "auto", out_scores=None, out_rows=None, require_bits=None):
        return self.index.query_sharded_dev(q, k, out_scores=out_scores, out_rows=out_rows, mode=mode,
                                            require_bits=require_bits)

    def search_stream(self, queries, k: int, mode: str = "auto", out_scores=None, out_rows=None):
        """Throughput mode for a stream of independent queries (each ``[b_i, dim]``): every query
        kernel pushes its candidates into its own slots, ONE collect kernel merges them all."""
        self.index.exchange_begin()
        slot = 0
        for q in queries:
            self.index.query_push_dev(q, k, slot, mode=mode)
            slot += 1 if q.dim() == 1 else q.shape[0]
        return self.index.exchange_collect_dev(slot, k, out_scores=out_scores, out_rows=out_rows)

    def exchange_merge(self, s, r, out_scores=None, out_rows=None):
        return self.index.exchange_merge_dev(s, r, out_scores=out_scores, out_rows=out_rows)


class ShardedSearcher:
    """``local_topk(q, k) -> (scores [B,k] f32, rows [B,k] i64 GLOBAL rows, -1 = empty)`` and
    ``merge(cand_scores [G,B,k], cand_rows [G,B,k]) -> (scores [B,k], rows [B,k])`` operate on
    torch tensors living on the device the process group communicates on."""

    def __init__(self, local_topk: Callable, merge: Callable, group=None, peer_exchange: Optional[PeerExchange] = None,
                 mode: str = "auto"):
        import torch.distributed as dist
        self._dist = dist
        self.local_topk, self.merge, self.group = local_topk, merge, group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.peer_exchange, self.mode = peer_exchange, mode
        self._gs = self._gr = None

    @property
    def exchange(self) -> str:
        return "p2p" if self.peer_exchange is not None else "nccl"

    @classmethod
    def for_index(cls, index, group=None, mode: str = "auto", exchange: str = "nccl", b_max: int = 1024,
                  k_max: int = 32):
        """Product wiring: K1/K2 for the local scan, K5 (or the fused peer exchange) for the merge."""
        import torch.distributed as dist
        px = None
        if exchange == "p2p" and dist.is_initialized() and dist.get_world_size(group) > 1:
            px = PeerExchange(index, group, b_max, k_max)
        elif exchange not in ("nccl", "p2p"):
            raise ValueError("exchange must be 'nccl' or 'p2p'")
        return cls(lambda q, k, bits=None: index.query_dev(q, k, mode=mode, require_bits=bits),
                   lambda cs, cr: index.merge_dev(cs, cr), group, px, mode)

    def search(self, q, k: int, require_bits=None):
        """``require_bits``: filter-bit indices every returned row must carry ("pre" filter mode)."""
        import torch
        if self.peer_exchange is not None and self.world_size > 1 and k <= self.peer_exchange.k_max:
            return self.peer_exchange.search(q, k, self.mode, require_bits=require_bits)   # kernel + fused exchange
        s, r = self.local_topk(q, k) if require_bits is None else self.local_topk(q, k, require_bits)
        if self.world_size == 1:
            return s, r
        B = s.shape[0]
        if self._gs is None or self._gs.shape != (self.world_size, B, k) or self._gs.device != s.device:
            self._gs = torch.empty((self.world_size, B, k), dtype=torch.float32, device=s.device)
            self._gr = torch.empty((self.world_size, B, k), dtype=torch.int64, device=s.device)
        # [G*B, k] view: the concatenating form is accepted by both NCCL and gloo
        self._dist.all_gather_into_tensor(self._gs.view(-1, k), s.contiguous(), group=self.group)
        self._dist.all_gather_into_tensor(self._gr.view(-1, k), r.contiguous(), group=self.group)
        return self.merge(self._gs, self._gr)
