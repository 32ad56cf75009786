"""-m gpu: the single-process multi-GPU collection (vs_group_t / GroupIndex) against the CPU oracle and
against a single-GPU index holding the same rows.

On a 1-GPU box the group's shards all live on cuda:0 (``devices=[0, 0, 0]``): same worker threads, same
fused exchange protocol, same host-mapped request/response path as across NVLink.  With >= 2 GPUs the
same tests also run on distinct devices.
"""
import numpy as np
import pytest

from oracle import cosine_oracle as O

pytestmark = pytest.mark.gpu
TOL = {"f32": 1e-5, "bf16": 2e-3}


def _device_sets():
    import torch
    sets = [[0], [0, 0, 0]]
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n >= 2:
        sets.append(list(range(min(n, 8))))
    return sets


def _check(s, r, full, k, tol):
    for b in range(s.shape[0]):
        ok, why = O.topk_matches(s[b], r[b], full[b], k, tol)
        assert ok, f"query {b}: {why}"


@pytest.mark.parametrize("devices", _device_sets(), ids=lambda d: "dev" + "".join(map(str, d)))
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_group_query_paths_match_oracle(gpu, devices, dtype):
    rng = np.random.default_rng(len(devices))
    n, d = 20_011, 512
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[n - 1] = X[7]                                   # exact tie that lands on two different shards
    Q = np.concatenate([rng.standard_normal((99, d)).astype(np.float32), X[7:8]])
    gx = gpu.GroupIndex(d, dtype, devices=devices, b_max=64, k_max=128)
    assert gx.add(X[:12_000]) == 0 and gx.add(X[12_000:]) == 12_000 and len(gx) == n
    full = O.cosine_scores(Q, X, corpus_dtype=dtype)
    # request/response path: one query at a time (the reference's service shape), fused scan + exchange
    for b in (0, 1, 99):
        for rep in range(3):
            s, r = gx.query(Q[b:b + 1], 10, mode="scan")
            _check(s, r, full[b:b + 1], 10, TOL[dtype])
    assert r[0][:2].tolist() == [7, n - 1]            # tie ordered by global row
    # batches inside one fused launch, across launches (B > 64 -> b_max chunks), k up to the exchange's k_max
    for B, k in ((7, 10), (64, 32), (100, 10), (5, 100)):
        s, r = gx.query(Q[:B], k, mode="scan")
        _check(s, r, full[:B], k, TOL[dtype])
    # k above k_max (the UI's "All" = 1000): candidates gathered onto GPU 0 and merged there
    s, r = gx.query(Q[:3], 1000, mode="scan")
    _check(s, r, full[:3], 1000, TOL[dtype])
    if dtype == "bf16":
        fullr = O.cosine_scores(Q, X, corpus_dtype="bf16", round_queries=True)
        s, r = gx.query(Q, 10, mode="tensor")         # tcgen05 path + exchange kernel
        _check(s, r, fullr, 10, TOL[dtype])
        s, r = gx.query(Q[:40], 10, mode="auto")
        _check(s, r, fullr[:40], 10, TOL[dtype])
    assert gx.exchange_error() == 0
    gx.close()


@pytest.mark.parametrize("devices", _device_sets()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
def test_group_matches_single_index_through_collection(gpu, devices):
    """Collection(index=GroupIndex) == Collection on one GPU: ids, distances, filters, deletes, multimodal."""
    rng = np.random.default_rng(5)
    n, d = 9_001, 256
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[n - 1] = X[11]
    Q = np.concatenate([rng.standard_normal((5, d)).astype(np.float32), X[11:12]])
    ids = [f"img_{i:05x}" for i in range(n)]
    metas = [{"filename": f"{i}.jpg", "filter_results_json": '{"is it red?": "%s"}' % ("yes" if i % 4 == 0 else "no")}
             for i in range(n)]
    many = gpu.Collection("many", {"hnsw:space": "cosine"}, index=gpu.GroupIndex(d, "bf16", devices=devices, b_max=64))
    one = gpu.Collection("one", {"hnsw:space": "cosine"}, dtype="bf16", device=0)
    for col in (one, many):
        col.add(ids=ids[:5000], embeddings=X[:5000], metadatas=metas[:5000])
        col.add(ids=ids[5000:], embeddings=X[5000:], metadatas=metas[5000:])

    def same(**kw):
        a = one.query(query_embeddings=Q.tolist(), **kw)
        b = many.query(query_embeddings=Q.tolist(), **kw)
        assert a["ids"] == b["ids"], kw
        for da, db in zip(a["distances"], b["distances"]):
            np.testing.assert_allclose(da, db, atol=1e-6)
    same(n_results=10, include=["metadatas", "distances"])
    same(n_results=10, include=["distances"], where_filters=["is it red?"], filter_mode="pre")
    same(n_results=10, include=["distances"], where_filters=["is it red?"], filter_mode="post")
    same(n_results=1000, include=["distances"])
    a = one.query_multimodal(Q[:3], Q[3:6], [0.25, 0.5, 0.9], n_results=10, include=["distances"])
    b = many.query_multimodal(Q[:3], Q[3:6], [0.25, 0.5, 0.9], n_results=10, include=["distances"])
    assert a["ids"] == b["ids"]
    # single delete (last global row moves across shards), bulk delete (compaction plan over the shards)
    for col in (one, many):
        col.delete(ids=[ids[3]])
        col.delete(ids=[ids[i] for i in range(100, 3000, 7)] + [ids[n - 1], ids[n - 2]])
        col.update(ids=[ids[4]], metadatas=[{"filter_results_json": '{"is it red?": "no"}'}])
    assert many.count() == one.count()
    assert many.get(include=[])["ids"] == one.get(include=[])["ids"]       # same compaction plan on both
    same(n_results=10, include=["metadatas", "distances"])
    same(n_results=10, include=["distances"], where_filters=["is it red?"], filter_mode="pre")
    np.testing.assert_array_equal(many.index.get_rows(0, 50), one.index.get_rows(0, 50))
    # filter sweep + dedup on the group == on one GPU
    P = rng.standard_normal((3, d)).astype(np.float32)
    np.testing.assert_array_equal(many.filter_sweep(P, 0.05), one.filter_sweep(P, 0.05))
    for col in (one, many):
        assert col.apply_filter_sweep("looks like P0", P[0], 0.05) == int(one.filter_sweep(P[:1], 0.05).sum())
    same(n_results=10, include=["metadatas", "distances"], where_filters=["looks like P0"], filter_mode="pre")
    same(n_results=50, include=["metadatas", "distances"], where_filters=["looks like P0"], filter_mode="post")
    assert set(many.find_duplicates(0.95)) == set(one.find_duplicates(0.95))
    # reset_system: delete everything in one call
    for col in (one, many):
        col.delete(ids=col.get(include=[])["ids"])
        assert col.count() == 0
        col.add(ids=["again"], embeddings=X[:1])
        assert col.query(query_embeddings=X[:1], n_results=5)["ids"] == [["again"]]
    many.close()
    one.close()


def test_exchange_timeout_returns_empty_and_reports(gpu):
    """A peer that never arrives: the wait times out (~3 s), the result is EMPTY (never a merge of stale
    lists), the sticky error word is raised and the host entry point returns VS_ERR_EXCHANGE."""
    import torch
    rng = np.random.default_rng(9)
    X = rng.standard_normal((4000, 64)).astype(np.float32)
    sh = []
    for g in range(2):
        ix = gpu.DeviceIndex(64, "f32", device=0, row_base=2000 * g)
        ix.add(X[2000 * g:2000 * (g + 1)])
        ix.exchange_create(2, g, 16, 32)
        sh.append(ix)
    ptrs = [ix.exchange_local_ptr() for ix in sh]
    for ix in sh:
        ix.exchange_attach(peer_ptrs=ptrs)
    q = torch.from_numpy(X[:1]).cuda()
    s, r = sh[0].query_sharded_dev(q, 5, mode="scan")          # rank 1 never issues its query
    torch.cuda.synchronize()
    assert (r.cpu().numpy() == -1).all() and np.isneginf(s.cpu().numpy()).all()
    assert sh[0].exchange_error() == 1
    with pytest.raises(gpu.VecSearchError) as ei:                # poisoned until acknowledged
        sh[0].query_sharded_dev(q, 5, mode="scan")
    assert ei.value.code == -6
    sh[0].exchange_clear_error()
    assert sh[0].exchange_error() == 0
    for ix in sh:
        ix.close()


def test_remove_rows_is_one_compaction_and_keeps_answers(gpu):
    rng = np.random.default_rng(13)
    n, d = 30_000, 128
    X = rng.standard_normal((n, d)).astype(np.float32)
    ix = gpu.DeviceIndex(d, "bf16")
    ix.add(X)
    ix.set_filter_bits_range(0, [[r % 3] for r in range(n)])
    gone = np.unique(rng.integers(0, n, 9000))
    launches = gpu.launch_count()
    src, dst = ix.remove_rows(gone)
    assert gpu.launch_count() - launches <= 2              # compaction kernel + group-bound refresh
    assert len(ix) == n - len(gone)
    perm = np.arange(n)
    perm[dst] = perm[src]
    perm = perm[:len(ix)]                                     # perm[new row] = original row
    assert len(set(perm.tolist()) & set(gone.tolist())) == 0 and len(set(perm.tolist())) == len(perm)
    np.testing.assert_array_equal(ix.get_rows(0, len(ix)), O.bf16_round(X[perm]))
    assert ix.get_filter_bits(int(dst[0])) == [int(src[0]) % 3]
    Q = rng.standard_normal((40, d)).astype(np.float32)
    for mode, rq in (("scan", False), ("tensor", True)):       # the tensor path reads the refreshed group bounds
        s, r = ix.query(Q, 10, mode=mode)
        full = O.cosine_scores(Q, X[perm], corpus_dtype="bf16", round_queries=rq)
        _check(s, r, full, 10, TOL["bf16"])
    ix.truncate(0)
    assert len(ix) == 0
    ix.close()


def test_raw_slab_roundtrip_is_bit_exact(gpu):
    rng = np.random.default_rng(17)
    X = rng.standard_normal((5000, 200)).astype(np.float32) * 3.0     # pitch 208 > dim: padded rows
    Q = rng.standard_normal((20, 200)).astype(np.float32)
    for dtype in ("bf16", "f32"):
        a = gpu.DeviceIndex(200, dtype)
        a.add(X)
        slab = a.get_raw(0, len(a))
        b = gpu.DeviceIndex(200, dtype)
        b.add_raw(slab[:3000])
        b.add_raw(slab[3000:])
        np.testing.assert_array_equal(b.get_raw(0, 5000), slab)
        sa, ra = a.query(Q, 10, mode="scan")
        sb, rb = b.query(Q, 10, mode="scan")
        np.testing.assert_array_equal(ra, rb)
        np.testing.assert_array_equal(sa, sb)                          # same stored bits, same inverse norms
        a.close()
        b.close()


def test_group_is_thread_safe_under_concurrent_queries_and_adds(gpu):
    """The filter task runs on a worker thread while routes keep querying (backend/app/main.py:410): concurrent callers on
    ONE group serialise on the front mutex / the shard mutexes and every answer stays exact for the rows present."""
    import threading
    rng = np.random.default_rng(23)
    n0, d = 6000, 128
    X = rng.standard_normal((n0 + 2000, d)).astype(np.float32)
    # (capacity reserved up front: with both shards on ONE GPU a slab re-allocation would synchronise the device while
    #  the peer shard's kernel waits for this shard's push -- on distinct GPUs that is only a delay)
    gx = gpu.GroupIndex(d, "f32", devices=[0, 0], capacity=n0 + 2000, b_max=64, k_max=32)
    gx.add(X[:n0])
    errors = []

    def querier(seed):
        r = np.random.default_rng(seed)
        try:
            for _ in range(60):
                i = int(r.integers(0, n0))
                s, rows = gx.query(X[i:i + 1], 5, mode="scan")
                if rows[0][0] != i or abs(s[0][0] - 1.0) > 1e-5:          # a stored row is its own nearest neighbour
                    errors.append(("wrong", i, rows[0].tolist(), s[0].tolist()))
        except Exception as e:                                             # noqa: BLE001
            errors.append(("exc", repr(e)))

    def adder():
        try:
            for c0 in range(n0, n0 + 2000, 250):
                gx.add(X[c0:c0 + 250])
        except Exception as e:                                             # noqa: BLE001
            errors.append(("exc", repr(e)))

    threads = [threading.Thread(target=querier, args=(s,)) for s in range(4)] + [threading.Thread(target=adder)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    assert len(gx) == n0 + 2000 and gx.exchange_error() == 0
    s, r = gx.query(X[n0 + 1999:n0 + 2000], 3, mode="scan")
    assert r[0][0] == n0 + 1999
    gx.close()
