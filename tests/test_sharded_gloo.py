"""CPU, world_size 2, gloo: the N>1 host logic (row partition, candidate all-gather, merge) gives
the same answer as a single shard.  The local search and the merge are injected (oracle here, the
CUDA kernels in production -- see ShardedSearcher.for_index)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmiss_b200.sharded import ShardedSearcher, shard_bounds
from oracle import cosine_oracle as O


def test_shard_bounds_cover_rows_exactly():
    for n, g in ((10_000_000, 8), (10, 3), (7, 8), (0, 2), (1_000_001, 4)):
        spans = [shard_bounds(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) == (n + g - 1) // g if n else True


def _worker(rank, world, port, n, d, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[n - 1] = X[3]                                   # an exact tie across shards
    Q = np.concatenate([rng.standard_normal((4, d)).astype(np.float32), X[3:4]])
    lo, hi = shard_bounds(n, world, rank)

    def local_topk(q, kk):
        s, r = O.cosine_topk(q.numpy(), X[lo:hi], kk)
        S = torch.full((q.shape[0], kk), float("-inf"))
        R = torch.full((q.shape[0], kk), -1, dtype=torch.int64)
        S[:, :s.shape[1]], R[:, :r.shape[1]] = torch.from_numpy(s), torch.from_numpy(r + lo)
        return S, R

    def merge(cs, cr):
        s, r = O.merge_topk(cs.numpy(), cr.numpy(), cs.shape[2])
        return torch.from_numpy(s), torch.from_numpy(r)

    searcher = ShardedSearcher(local_topk, merge)
    assert searcher.world_size == world and searcher.rank == rank
    s, r = searcher.search(torch.from_numpy(Q), k)
    full = O.cosine_scores(Q, X)
    for b in range(Q.shape[0]):
        ok, why = O.topk_matches(s[b].numpy(), r[b].numpy(), full[b], k, 1e-6)
        assert ok, why
    assert r[4][:2].tolist() == [3, n - 1]            # tie resolved by global row on every rank
    got = [torch.zeros_like(r) for _ in range(world)]
    dist.all_gather(got, r)
    assert all(torch.equal(got[0], g) for g in got)   # every rank ends with the same answer
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_matches_single_shard():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1001, 32, 10, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


def test_triangle_bounds_cover_and_balance():
    from mmiss_b200.sharded import triangle_bounds
    for n, g in ((2_000_000, 8), (200_000, 4), (1000, 3), (100, 8), (0, 2)):
        spans = [triangle_bounds(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        if n >= 100_000:                       # equal WORK (pairs to the right), not equal rows
            work = [sum(n - 1 - i for i in (lo, hi - 1)) / 2 * (hi - lo) for lo, hi in spans]
            assert max(work) / (sum(work) / g) < 1.02
            assert spans[0][1] - spans[0][0] < spans[-1][1] - spans[-1][0]


def _dedup_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmiss_b200.sharded import find_duplicates_sharded
    rng = np.random.default_rng(5)
    n, d = 700, 48
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[600:650] = X[10:60] + 0.01 * rng.standard_normal((50, d)).astype(np.float32)
    wi, wj, ws = O.dedup_pairs(X, 0.95)

    def local(lo, hi):                                  # the oracle restricted to rows [lo, hi)
        m = (wi >= lo) & (wi < hi)
        return wi[m], wj[m], ws[m]

    i, j, s = find_duplicates_sharded(local, n)
    assert np.array_equal(i, wi) and np.array_equal(j, wj) and np.array_equal(s, ws) and len(i) >= 50
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_sharded_dedup():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_dedup_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


def _service_worker(rank, world, port, out):
    """rank 0 = the reference's single server process with an ordinary Collection on a
    ShardedIndex; the other ranks serve.  Every Collection operation the reference uses must give the
    same answers as a single-shard Collection (backend/app/main.py:735-740, 761-765, 503-510, 1069)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmiss_b200 import collection as C
    from mmiss_b200.sharded_index import ShardedIndex
    from tests.fake_index import FakeIndex

    def searcher_factory(ix):
        def local_topk(q, kk, bits=None):
            s, r = ix.query(q.numpy(), kk, require_bits=bits)
            return torch.from_numpy(s), torch.from_numpy(r)

        def merge(cs, cr):
            s, r = O.merge_topk(cs.numpy(), cr.numpy(), cs.shape[2])
            return torch.from_numpy(s), torch.from_numpy(r)
        return ShardedSearcher(local_topk, merge)

    sh = ShardedIndex(24, "f32", index_factory=FakeIndex, searcher_factory=searcher_factory)
    if rank != 0:
        sh.serve()
        dist.destroy_process_group()
        return
    C.DeviceIndex = FakeIndex
    one = C.Collection("one", {"hnsw:space": "cosine"})
    many = C.Collection("many", {"hnsw:space": "cosine"}, index=sh)
    rng = np.random.default_rng(3)
    X = rng.standard_normal((61, 24)).astype(np.float32)
    X[50] = X[3]                                   # an exact tie between rows on different shards (3 % 3 != 50 % 3)
    ids = [f"img_{i:02d}" for i in range(61)]
    metas = [{"filename": f"{i}.jpg", "filter_results_json": '{"is it red?": "%s"}' % ("yes" if i % 3 == 0 else "no")}
             for i in range(61)]
    for col in (one, many):
        col.add(ids=ids[:40], embeddings=X[:40].tolist(), metadatas=metas[:40])
        col.add(ids=ids[40:], embeddings=X[40:], metadatas=metas[40:])        # second batch: striping continues
        col.add(ids=ids[:2], embeddings=X[:2])                                 # existing ids are skipped
    assert many.count() == one.count() == 61 and len(sh) == 61
    Q = np.concatenate([rng.standard_normal((4, 24)).astype(np.float32), X[3:4]])
    assert many.query(query_embeddings=Q[4:5].tolist(), n_results=2)["ids"] == [[ids[3], ids[50]]]   # tie: lower global row first

    def same(**kw):
        a, b = one.query(query_embeddings=Q.tolist(), **kw), many.query(query_embeddings=Q.tolist(), **kw)
        assert a["ids"] == b["ids"], (kw, a["ids"], b["ids"])
        for da, db in zip(a["distances"], b["distances"]):                      # ragged after a post-filter
            np.testing.assert_allclose(da, db, atol=1e-6)
        assert a["metadatas"] == b["metadatas"]
    same(n_results=10, include=["metadatas", "distances"])
    same(n_results=100, include=["metadatas", "distances"])                    # clamped to count()
    same(n_results=7, include=["metadatas", "distances"], where_filters=["is it red?"], filter_mode="post")
    same(n_results=7, include=["metadatas", "distances"], where_filters=["is it red?"], filter_mode="pre")
    for col in (one, many):                                                    # delete: last row moves across shards
        col.delete(ids=[ids[5], ids[60], ids[17]])
        col.update(ids=[ids[6]], metadatas=[{"filter_results_json": '{"is it red?": "no"}'}])
    assert many.count() == 58
    same(n_results=10, include=["metadatas", "distances"])
    same(n_results=9, include=["metadatas", "distances"], where_filters=["is it red?"], filter_mode="pre")
    a = one.get(ids=[ids[30]], include=["embeddings", "metadatas"])
    b = many.get(ids=[ids[30]], include=["embeddings", "metadatas"])
    assert a["ids"] == b["ids"] and a["metadatas"] == b["metadatas"]
    np.testing.assert_allclose(a["embeddings"][0], b["embeddings"][0])
    mm = many.query_multimodal(Q[:2], Q[2:4], 0.3, n_results=4)
    mo = one.query_multimodal(Q[:2], Q[2:4], 0.3, n_results=4)
    assert mm["ids"] == mo["ids"]
    # config 4 / 5 on the sharded collection: shard-local sweep gathered + interleaved, triangle-split all-pairs pass
    P = rng.standard_normal((3, 24)).astype(np.float32)
    np.testing.assert_array_equal(many.filter_sweep(P, 0.1), one.filter_sweep(P, 0.1))
    for col in (one, many):
        assert col.apply_filter_sweep("like P1", P[1], 0.1) == int(one.filter_sweep(P[1:2], 0.1).sum())
        assert col.filter_progress["like P1"]["status"] == "completed"
    same(n_results=20, include=["metadatas", "distances"], where_filters=["like P1"], filter_mode="pre")
    same(n_results=20, include=["metadatas", "distances"], where_filters=["like P1", "is it red?"], filter_mode="post")
    assert many.find_duplicates(0.999) == one.find_duplicates(0.999) and len(one.find_duplicates(0.999)) == 1
    many.close()                                                               # releases the workers
    out.put("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_3_sharded_collection_service():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_service_worker, args=(r, 3, port, out)) for r in range(3)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


def _px_fail_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmiss_b200.sharded import PeerExchange

    class Stub:                                            # rank 1 cannot create its buffer
        device = 0

        def exchange_create(self, g, r, b, k):
            if r == 1:
                raise MemoryError("no memory for the exchange buffer")
            return b"h" * 64

        def exchange_attach(self, ipc_handles=None, peer_ptrs=None):
            raise AssertionError("must not attach when a peer has no buffer")
    try:
        PeerExchange(Stub(), None, 8, 8)
        out.put(f"rank {rank}: no error")
    except RuntimeError as e:
        assert "MemoryError" in str(e)
        out.put("raised")
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_peer_exchange_setup_fails_on_every_rank_together():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_px_fail_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    assert [out.get(timeout=5), out.get(timeout=5)] == ["raised", "raised"]
