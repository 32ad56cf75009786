"""Row-sharded search across the GPUs of one box: one process per GPU (torch.distributed).

SURVEY.md section 8(e): the corpus is split into contiguous row ranges, every rank scans its
shard for the (replicated) queries, the per-rank ``[B, k]`` candidates are exchanged with ONE
small all-gather (``B*k*12`` bytes per rank over NVLink/NVSwitch with NCCL) and merged on every
rank by the K5 kernel, so all ranks end with the same global answer.

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


class ShardedSearcher:
    """``local_topk(q, k) -> (scores [B,k] f32, rows [B,k] i64 GLOBAL rows, -1 = empty)`` and
    ``merge(cand_scores [G,B,k], cand_rows [G,B,k]) -> (scores [B,k], rows [B,k])`` operate on
    torch tensors living on the device the process group communicates on."""

    def __init__(self, local_topk: Callable, merge: Callable, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.local_topk, self.merge, self.group = local_topk, merge, group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._gs = self._gr = None

    @classmethod
    def for_index(cls, index, group=None, mode: str = "auto"):
        """Product wiring: K1/K2 for the local scan, K5 for the merge (all CUDA)."""
        return cls(lambda q, k: index.query_dev(q, k, mode=mode),
                   lambda cs, cr: index.merge_dev(cs, cr), group)

    def search(self, q, k: int):
        import torch
        s, r = self.local_topk(q, k)
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
