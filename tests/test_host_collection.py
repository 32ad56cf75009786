"""CPU: the host-side mirror of the chromadb surface (Collection / PersistentClient / SearchService)
with the device index replaced by tests/fake_index.py.  Checks the contract the reference relies on
(call sites in /root/reference/backend/app/main.py cited per test) and replays the golden vectors
recorded from the real reference code (tests/golden/reference_golden.json)."""
import json
import os
import threading

import numpy as np
import pytest

import mmiss_b200
from mmiss_b200 import collection as C
from oracle import cosine_oracle as O
from tests.fake_index import FakeIndex

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_golden.json")) as f:
    REF = json.load(f)


@pytest.fixture(autouse=True)
def fake_device(monkeypatch):
    monkeypatch.setattr(C, "DeviceIndex", FakeIndex)


def _coll(**kw):
    return C.Collection("image-match", {"hnsw:space": "cosine"}, **kw)


def _vecs(n, d=16, seed=0):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


def test_init_chromadb_sequence(tmp_path):
    """backend/app/utils.py:113-130: PersistentClient -> list_collections (names) -> create / get."""
    client = mmiss_b200.PersistentClient(path=str(tmp_path))
    assert client.list_collections() == []
    col = client.create_collection(name="image-match", metadata={"hnsw:space": "cosine"})
    assert "image-match" in client.list_collections()
    assert client.get_collection(name="image-match") is col
    with pytest.raises(ValueError):
        client.create_collection(name="image-match")
    with pytest.raises(ValueError):
        client.get_collection("nope")
    with pytest.raises(ValueError):
        client.create_collection("l2", metadata={"hnsw:space": "l2"})


def test_add_query_shapes_and_distance_semantics():
    """main.py:735-740 (add one row with python lists) and :761-777 (query -> lists per query)."""
    col = _coll()
    X = _vecs(20)
    for i in range(20):
        col.add(ids=[f"img_{i}"], embeddings=[X[i].tolist()], metadatas=[{"id": f"img_{i}", "filename": f"{i}.jpg"}],
                documents=[f"desc {i}"])
    assert col.count() == 20
    res = col.query(query_embeddings=[X[3].tolist()], n_results=5, include=["metadatas", "distances"])
    assert set(res) >= {"ids", "distances", "metadatas", "documents", "embeddings"}
    assert len(res["ids"]) == 1 and len(res["ids"][0]) == 5
    assert res["documents"] is None and res["embeddings"] is None
    assert res["ids"][0][0] == "img_3" and abs(res["distances"][0][0]) < 1e-6       # distance = 1 - cos
    assert res["distances"][0] == sorted(res["distances"][0])                        # ascending distance
    assert res["metadatas"][0][0] == {"id": "img_3", "filename": "3.jpg"}
    full = O.cosine_scores(X[3], X)[0]
    want = np.argsort(-full, kind="stable")[:5]
    assert res["ids"][0] == [f"img_{i}" for i in want]
    np.testing.assert_allclose(res["distances"][0], 1.0 - full[want], atol=1e-6)
    # batched form
    res = col.query(query_embeddings=X[:3], n_results=2)
    assert [r[0] for r in res["ids"]] == ["img_0", "img_1", "img_2"]


def test_n_results_clamped_and_empty_collection():
    col = _coll()
    res = col.query(query_embeddings=[[0.0] * 8], n_results=5, include=["metadatas", "distances"])
    assert res["ids"] == [[]] and res["distances"] == [[]] and res["metadatas"] == [[]]
    col.add(ids=["a", "b"], embeddings=_vecs(2, 8))
    res = col.query(query_embeddings=_vecs(1, 8, 1), n_results=1000)          # the UI's "All" (main.py:757)
    assert len(res["ids"][0]) == 2


def test_query_texts_raises_for_legacy_fallback():
    """app.py:343-372 tries query_texts first and relies on the exception to fall back to CLIP."""
    col = _coll()
    col.add(ids=["a"], embeddings=_vecs(1, 8))
    with pytest.raises(ValueError):
        col.query(query_texts=["red drill"], n_results=3)


def test_get_contract_used_by_duplicate_check_and_startup():
    """main.py:533 (include=[] -> ids), :563-566, :631-640 (unknown id -> empty lists)."""
    col = _coll()
    col.add(ids=["img_a", "img_b"], embeddings=_vecs(2, 8), metadatas=[{"id": "img_a"}, {"id": "img_b"}])
    assert col.get(include=[])["ids"] == ["img_a", "img_b"]
    hit = col.get(ids=["img_b"], include=["metadatas"])
    assert hit["ids"] == ["img_b"] and hit["metadatas"] == [{"id": "img_b"}]
    assert hit["ids"].index("img_b") == 0
    miss = col.get(ids=["img_zzz"], include=["metadatas"])
    assert miss["ids"] == [] and miss["metadatas"] == []
    both = col.get(ids=["img_zzz", "img_a"], include=["metadatas"])
    assert both["ids"] == ["img_a"]


def test_add_existing_id_is_skipped_and_dimension_checked():
    col = _coll()
    col.add(ids=["a"], embeddings=_vecs(1, 8), metadatas=[{"v": "1"}])
    col.add(ids=["a", "b"], embeddings=_vecs(2, 8, 5), metadatas=[{"v": "2"}, {"v": "3"}])
    assert col.count() == 2 and col.get(ids=["a"])["metadatas"] == [{"v": "1"}]
    with pytest.raises(ValueError):
        col.add(ids=["c"], embeddings=_vecs(1, 9))
    with pytest.raises(ValueError):
        col.add(ids=["c"], embeddings=_vecs(1, 8), metadatas=[{"nested": {"x": 1}}])


def test_update_merges_metadata():
    """main.py:503-510 passes a PARTIAL dict; chromadb merges keys."""
    col = _coll()
    col.add(ids=["a"], embeddings=_vecs(1, 8), metadatas=[{"id": "a", "filename": "a.jpg", "description": "old"}],
            documents=["old"])
    col.update(ids=["a"], metadatas=[{"description": "new", "custom_metadata": "x"}], documents=["new"])
    got = col.get(ids=["a"], include=["metadatas", "documents"])
    assert got["metadatas"] == [{"id": "a", "filename": "a.jpg", "description": "new", "custom_metadata": "x"}]
    assert got["documents"] == ["new"]
    col.update(ids=["ghost"], metadatas=[{"x": "y"}])            # unknown id: ignored
    assert col.count() == 1


def test_delete_keeps_host_and_device_rows_aligned():
    """main.py:1069 bulk delete; the engine moves the last row into the hole."""
    col = _coll()
    X = _vecs(10, 8)
    col.add(ids=[f"i{i}" for i in range(10)], embeddings=X, metadatas=[{"n": i} for i in range(10)])
    col.delete(ids=["i3", "i9", "i0"])
    assert col.count() == 7 and sorted(col.get(include=[])["ids"]) == sorted(f"i{i}" for i in (1, 2, 4, 5, 6, 7, 8))
    for i in (1, 2, 4, 5, 6, 7, 8):
        res = col.query(query_embeddings=[X[i]], n_results=1, include=["metadatas", "distances"])
        assert res["ids"][0] == [f"i{i}"] and res["metadatas"][0] == [{"n": i}] and abs(res["distances"][0][0]) < 1e-6
    col.delete(ids=col.get(include=[])["ids"])
    assert col.count() == 0


def test_filter_pass_post_and_pre_modes():
    """'post' = the reference's order of operations (main.py:201-222: top-k, then the predicate);
    'pre' = predicate inside the search (more useful: always n_results rows when enough match)."""
    case = REF["filter_pass"][2]                      # filters: is red? + has cord?
    metas, dists = case["input_metadatas"], case["distances"]
    col = _coll()
    rng = np.random.default_rng(0)
    q = rng.standard_normal(8).astype(np.float32)
    q /= np.linalg.norm(q)
    # build rows whose cosine to q reproduces the recorded distances, in order
    X = []
    for d in dists:
        c = 1.0 - d
        o = rng.standard_normal(8).astype(np.float32)
        o -= o.dot(q) * q
        o /= np.linalg.norm(o)
        X.append(c * q + np.sqrt(max(0.0, 1 - c * c)) * o)
    col.add(ids=[m["id"] for m in metas], embeddings=np.stack(X), metadatas=metas)
    assert col.filter_names() == ["is red?", "has cord?"]
    for fc in REF["filter_pass"]:
        res = col.query(query_embeddings=[q], n_results=10, include=["metadatas"], where_filters=fc["filters"],
                        filter_mode="post")
        assert res["ids"][0] == fc["kept_ids"], fc["filters"]
        res = col.query(query_embeddings=[q], n_results=10, include=["metadatas"], where_filters=fc["filters"],
                        filter_mode="pre")
        assert res["ids"][0] == fc["kept_ids"], fc["filters"]
    # post mode truncates BEFORE filtering, pre mode after
    post = col.query(query_embeddings=[q], n_results=1, where_filters=["has cord?"], filter_mode="post")
    pre = col.query(query_embeddings=[q], n_results=1, where_filters=["has cord?"], filter_mode="pre")
    assert post["ids"][0] == [] and pre["ids"][0] == [metas[1]["id"]]


def test_search_service_matches_reference_golden():
    """search_similar / apply_filters reproduce the outputs recorded from the reference module."""
    for case in REF["search_similar"]:
        class Canned:
            def query(self, **kw):
                assert sorted(kw) == case["query_kwargs"] and kw["n_results"] == case["n_results"]
                assert kw["include"] == case["include"] and isinstance(kw["query_embeddings"][0], list)
                return {"ids": [case["ids"]], "metadatas": [case["metadatas"]], "distances": [case["distances"]]}
        svc = mmiss_b200.SearchService(Canned())
        assert svc.search_similar(np.zeros(REF["dim"], np.float32), limit=case["limit"]) == case["results"]
    for fc in REF["filter_pass"]:
        ranked = [dict(m, similarity_score=1 - d / 2) for m, d in zip(fc["input_metadatas"], fc["distances"])]
        assert [r["id"] for r in mmiss_b200.apply_filters(ranked, fc["filters"])] == fc["kept_ids"]

    class Boom:
        def query(self, **kw):
            raise RuntimeError("backend down")
    assert mmiss_b200.SearchService(Boom()).search_similar(np.zeros(4, np.float32)) == []      # main.py:803-805
    assert mmiss_b200.similarity_from_distance([0.5], legacy=True) == [0.5]                        # app.py:326


def test_service_routes_and_duplicate_check_end_to_end():
    d = 32
    rng = np.random.default_rng(3)
    table = {}

    def encoder(image=None, text=None):            # deterministic stub of generate_clip_embedding
        key = ("i", image) if image is not None else ("t", text)
        if key not in table:
            v = rng.standard_normal(d).astype(np.float32)
            table[key] = v / np.linalg.norm(v)
        return {"image" if image is not None else "text": table[key][None]}

    col = _coll()
    svc = mmiss_b200.SearchService(col, encoder)
    for i in range(12):
        meta = {"id": f"img_{i:016x}", "filename": f"{i}.jpg", "url": f"/static/uploads/{i}.jpg",
                "filter_results_json": json.dumps({"is red?": "yes" if i % 2 == 0 else "no"})}
        got, is_new = svc.add_embedding(meta["id"], encoder(image=f"img{i}")["image"][0], meta, f"drill {i}")
        assert is_new and got == meta
    got, is_new = svc.add_embedding("img_%016x" % 3, encoder(image="other")["image"][0], {"id": "x"}, "dup")
    assert not is_new and got["filename"] == "3.jpg"                      # main.py:627-640 -> HTTP 409
    out = svc.route_search_image("img5", filters=None, limit=5)
    assert [r["id"] for r in out["results"]][0] == "img_%016x" % 5
    assert abs(out["results"][0]["similarity_score"] - 1.0) < 1e-6       # 1 - d/2 with d = 0
    assert out["results"][0]["thumbnail_url"] == "/static/processed/img_%016x.png" % 5
    out = svc.route_search_text("red drill", filters=["is red?"], limit=0)          # limit 0 -> "All"
    assert len(out["results"]) == 6 and all(int(r["id"][4:], 16) % 2 == 0 for r in out["results"])
    scores = [r["similarity_score"] for r in out["results"]]
    assert scores == sorted(scores, reverse=True)
    # multimodal: blend done by the index (device kernel in production), equals the reference blend
    out = svc.route_search_multimodal("img5", "red drill", weight_image=0.7, limit=3)
    c = O.blend(table[("i", "img5")], table[("t", "red drill")], 0.7)
    want = col.query(query_embeddings=[c], n_results=3, include=["metadatas", "distances"])
    assert [r["id"] for r in out["results"]] == want["ids"][0]


def test_persistence_replays_log(tmp_path):
    client = mmiss_b200.PersistentClient(path=str(tmp_path), dtype="f32")
    col = client.create_collection("image-match", metadata={"hnsw:space": "cosine"})
    X = _vecs(6, 8)
    col.add(ids=[f"i{i}" for i in range(6)], embeddings=X, metadatas=[{"n": i} for i in range(6)],
            documents=[f"d{i}" for i in range(6)])
    col.update(ids=["i2"], metadatas=[{"filter_results_json": json.dumps({"f": "yes"})}])
    col.delete(ids=["i4"])
    col.close()
    client2 = mmiss_b200.PersistentClient(path=str(tmp_path))
    assert client2.list_collections() == ["image-match"]
    col2 = client2.get_collection("image-match")
    assert col2.count() == 5 and set(col2.get(include=[])["ids"]) == {"i0", "i1", "i2", "i3", "i5"}
    assert col2.get(ids=["i2"])["metadatas"][0] == {"n": 2, "filter_results_json": json.dumps({"f": "yes"})}
    res = col2.query(query_embeddings=[X[5]], n_results=1, include=["documents", "distances"])
    assert res["ids"][0] == ["i5"] and res["documents"][0] == ["d5"]
    assert col2.query(query_embeddings=[X[0]], n_results=5, where_filters=["f"], filter_mode="pre")["ids"][0] == ["i2"]


def test_update_from_worker_thread_while_querying():
    """main.py:410: process_filter_on_all_images runs on a threadpool worker and calls
    collection.update while the event-loop thread may be inside collection.query."""
    col = _coll()
    X = _vecs(200, 8)
    col.add(ids=[f"i{i}" for i in range(200)], embeddings=X, metadatas=[{"n": i} for i in range(200)])
    errors = []

    def worker():
        try:
            for i in range(200):
                col.update(ids=[f"i{i}"], metadatas=[{"filter_results_json": json.dumps({"f": "yes"})}])
        except Exception as e:                      # pragma: no cover
            errors.append(e)

    t = threading.Thread(target=worker)
    t.start()
    for i in range(100):
        res = col.query(query_embeddings=[X[i]], n_results=3, include=["metadatas", "distances"])
        assert res["ids"][0][0] == f"i{i}"
    t.join()
    assert not errors
    assert len(col.query(query_embeddings=[X[0]], n_results=200, where_filters=["f"], filter_mode="pre")["ids"][0]) == 200


def test_filter_sweep_and_duplicates_api():
    col = _coll(dtype="bf16")
    g = np.load(os.path.join(HERE, "golden", "oracle_golden.npz"))
    Xd = g["Xd"]
    col.add(ids=[f"r{i}" for i in range(Xd.shape[0])], embeddings=Xd)
    mask = col.filter_sweep(g["F"], float(g["filter_tau"]))
    np.testing.assert_array_equal(mask, g["filter_mask"])
    n_yes = col.apply_filter_sweep("is drill?", g["F"][0], float(g["filter_tau"]))
    assert n_yes == int(g["filter_mask"][0].sum())
    kept = col.query(query_embeddings=[Xd[0]], n_results=300, where_filters=["is drill?"], filter_mode="post")
    assert len(kept["ids"][0]) == n_yes
    dups = col.find_duplicates(float(g["dedup_tau"]))
    assert [(a, b) for a, b, _ in dups] == [(f"r{i}", f"r{j}") for i, j in zip(g["dedup_i"], g["dedup_j"])]


def test_micro_batcher_stacks_concurrent_queries():
    """concurrent single queries are served by ONE batched collection.query; every
    caller gets what its own query would have returned (main.py:761-765 shape)."""
    col = _coll()
    X = _vecs(200, 16, seed=3)
    col.add(ids=[f"img_{i}" for i in range(200)], embeddings=X, metadatas=[{"filename": f"{i}.jpg"} for i in range(200)])
    mb = mmiss_b200.MicroBatcher(col, max_batch=16, max_wait_ms=200.0)
    Q = _vecs(12, 16, seed=4)
    want = [col.query(query_embeddings=[q.tolist()], n_results=3 + (i % 4), include=["metadatas", "distances"])
            for i, q in enumerate(Q)]
    got = [None] * len(Q)

    def worker(i):
        got[i] = mb.query(Q[i], n_results=3 + (i % 4))
    ts = [threading.Thread(target=worker, args=(i,)) for i in range(len(Q))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(30)
    mb.close()
    for i in range(len(Q)):
        assert got[i]["ids"] == want[i]["ids"] and len(got[i]["ids"][0]) == 3 + (i % 4)
        np.testing.assert_allclose(got[i]["distances"][0], want[i]["distances"][0], atol=1e-6)
        assert got[i]["metadatas"] == want[i]["metadatas"]
    assert sum(mb.batches) == len(Q) and max(mb.batches) > 1      # requests really were stacked


def test_micro_batcher_propagates_errors():
    col = _coll()
    col.add(ids=["a"], embeddings=_vecs(1, 16))
    mb = mmiss_b200.MicroBatcher(col, max_wait_ms=1.0)
    with pytest.raises(ValueError):
        mb.query(np.zeros(7, np.float32), 1)          # wrong dimension: the collection's error reaches the caller
    mb.close()
    with pytest.raises(RuntimeError):
        mb.query(np.zeros(16, np.float32), 1)


def test_snapshot_plus_log_format_and_compaction(tmp_path):
    """raw row slab + ids + metadata JSONL snapshot, operation log since, compaction on checkpoint
    (the store the reference reopens at backend/app/utils.py:109-123 and walks at main.py:522-579)."""
    client = mmiss_b200.PersistentClient(path=str(tmp_path), dtype="f32")
    col = client.create_collection("c", metadata={"hnsw:space": "cosine"})
    X = _vecs(50, 8)
    ids = [f"i{i}" for i in range(50)]
    col.add(ids=ids, embeddings=X, metadatas=[{"n": i} for i in range(50)], documents=[f"d{i}" for i in range(50)])
    col.delete(ids=ids[10:30])
    d = os.path.join(str(tmp_path), "c")
    assert os.path.getsize(os.path.join(d, "oplog.0.vec")) == 50 * 8 * 4          # the log still holds the deleted rows
    col.persist()                                                                  # checkpoint: generation 1, compacted
    info = json.load(open(os.path.join(d, "collection.json")))
    assert info["gen"] == 1 and info["count"] == 30 and info["dim"] == 8 and info["format"] == C.FORMAT_VERSION
    assert os.path.getsize(os.path.join(d, "rows.1.bin")) == 30 * 8 * 4           # deleted rows are gone from disk
    assert not os.path.exists(os.path.join(d, "oplog.0.vec")) and os.path.getsize(os.path.join(d, "oplog.1.jsonl")) == 0
    assert open(os.path.join(d, "ids.1.txt")).read().split("\n")[:30] == col.get(include=[])["ids"]
    col.add(ids=["late"], embeddings=X[:1] * 2.0, metadatas=[{"n": -1}])           # lands in the new log
    col.update(ids=["i3"], metadatas=[{"k": "v"}])
    want = col.query(query_embeddings=X[:4], n_results=5, include=["metadatas", "documents", "distances"])
    col._log.close(); col._vec.close(); col._log = None                            # crash: no close(), no checkpoint
    col2 = mmiss_b200.PersistentClient(path=str(tmp_path)).get_collection("c")     # snapshot upload + log replay
    assert col2.count() == 31
    assert isinstance(col2._metas[0], bytes)                                       # metadata lines stay unparsed until read
    got = col2.query(query_embeddings=X[:4], n_results=5, include=["metadatas", "documents", "distances"])
    assert got["ids"] == want["ids"] and got["metadatas"] == want["metadatas"] and got["documents"] == want["documents"]
    np.testing.assert_allclose(got["distances"], want["distances"], atol=1e-7)
    assert col2.get(ids=["i3"])["metadatas"][0] == {"n": 3, "k": "v"}
    col2.close()                                                                   # folds the log: generation 2
    assert json.load(open(os.path.join(d, "collection.json")))["gen"] == 2


def test_torn_log_tail_and_orphan_vectors_are_ignored(tmp_path):
    """a crash between the vector write and its log line must not shift later embeddings."""
    client = mmiss_b200.PersistentClient(path=str(tmp_path), dtype="f32")
    col = client.create_collection("c")
    X = _vecs(5, 8)
    col.add(ids=["a", "b"], embeddings=X[:2])
    col.add(ids=["c"], embeddings=X[2:3])
    col._log.flush(); col._vec.flush()
    d = os.path.join(str(tmp_path), "c")
    with open(os.path.join(d, "oplog.0.vec"), "ab") as f:                          # orphan vector: its record never made it
        f.write(X[3].tobytes())
    with open(os.path.join(d, "oplog.0.jsonl"), "ab") as f:                        # torn record
        f.write(b'{"op":"add","ids":["d"],"off":96,"nby')
    col._log.close(); col._vec.close(); col._log = None
    col2 = mmiss_b200.PersistentClient(path=str(tmp_path)).get_collection("c")
    assert col2.get(include=[])["ids"] == ["a", "b", "c"]
    assert os.path.getsize(os.path.join(d, "oplog.0.vec")) == 3 * 8 * 4            # truncated back to a clean boundary
    col2.add(ids=["e"], embeddings=X[4:5])                                         # appends after the clean boundary
    col2._log.close(); col2._vec.close(); col2._log = None
    col3 = mmiss_b200.PersistentClient(path=str(tmp_path)).get_collection("c")
    assert col3.get(include=[])["ids"] == ["a", "b", "c", "e"]
    np.testing.assert_allclose(col3.get(ids=["e"], include=["embeddings"])["embeddings"][0], X[4])


def test_bulk_delete_and_reset_match_sequential_semantics():
    """collection.delete(ids=all_ids) (backend/app/main.py:1065-1069) and arbitrary bulk deletes keep ids,
    metadata and device rows aligned."""
    col = _coll()
    X = _vecs(200, 8)
    ids = [f"i{i}" for i in range(200)]
    col.add(ids=ids, embeddings=X, metadatas=[{"n": i} for i in range(200)])
    gone = [ids[i] for i in (0, 5, 6, 7, 150, 198, 199, 42)] + ["nope"]
    col.delete(ids=gone)
    assert col.count() == 192
    for i in (1, 8, 100, 197):
        res = col.query(query_embeddings=[X[i]], n_results=1, include=["metadatas", "distances"])
        assert res["ids"][0] == [ids[i]] and res["metadatas"][0][0] == {"n": i} and abs(res["distances"][0][0]) < 1e-6
    assert col.get(ids=["i5"])["ids"] == []
    svc = mmiss_b200.SearchService(col)
    assert svc.reset_system() and col.count() == 0
    col.add(ids=["x"], embeddings=X[:1])
    assert col.query(query_embeddings=X[:1], n_results=3)["ids"] == [["x"]]


def test_filter_sweep_answers_are_lazy_and_follow_the_progress_contract(tmp_path):
    """process_filter_on_all_images (main.py:939-1056) on CLIP embeddings: answers land in the
    rows' filter bits, `filter_results_json` is materialised on read, progress dict has the reference's shape."""
    client = mmiss_b200.PersistentClient(path=str(tmp_path), dtype="f32")
    col = client.create_collection("c")
    X = _vecs(40, 8)
    ids = [f"i{i}" for i in range(40)]
    col.add(ids=ids, embeddings=X, metadatas=[{"n": i, "filter_results_json": json.dumps({"old": "yes" if i % 2 else "no"})}
                                               for i in range(40)])
    svc = mmiss_b200.SearchService(col, encoder=lambda text=None, image=None: {"text": X[3:4]})
    assert svc.get_filter_progress("is it X3?") == {"status": "not_found"}
    svc.process_filter_on_all_images("is it X3?", tau=0.5)
    prog = svc.get_filter_progress("is it X3?")
    assert prog["status"] == "completed" and prog["progress"] == 100 and prog["processed"] == prog["total"] == 40
    mask = col.filter_sweep(X[3:4], 0.5)[0]
    assert prog["matched"] == int(mask.sum()) >= 1
    raw = col._metas[3]
    assert "is it X3?" not in raw["filter_results_json"]                            # nothing was rewritten per row
    seen = json.loads(col.get(ids=["i3"])["metadatas"][0]["filter_results_json"])
    assert seen == {"old": "yes", "is it X3?": "yes"}
    res = svc.route_search_text("whatever", filters=["is it X3?"], limit=0)         # the reference's post-filter sees them
    assert sorted(r["n"] for r in res["results"]) == np.flatnonzero(mask).tolist()
    pre = col.query(query_embeddings=X[:1], n_results=40, where_filters=["is it X3?", "old"], filter_mode="pre")
    assert sorted(pre["ids"][0]) == sorted(ids[i] for i in np.flatnonzero(mask) if i % 2)
    col.update(ids=["i3"], metadatas=[{"filter_results_json": json.dumps({"old": "no"})}])   # swept bit survives an update
    assert json.loads(col.get(ids=["i3"])["metadatas"][0]["filter_results_json"]) == {"old": "no", "is it X3?": "yes"}
    col.delete(ids=["i0", "i1"])                                                    # answers move with their rows
    assert json.loads(col.get(ids=["i3"])["metadatas"][0]["filter_results_json"])["is it X3?"] == "yes"
    col._log.close(); col._vec.close(); col._log = None                            # crash + replay keeps the sweep
    col2 = mmiss_b200.PersistentClient(path=str(tmp_path)).get_collection("c")
    assert json.loads(col2.get(ids=["i3"])["metadatas"][0]["filter_results_json"]) == {"old": "no", "is it X3?": "yes"}
    col2.close()
    col3 = mmiss_b200.PersistentClient(path=str(tmp_path)).get_collection("c")     # ... and so does the snapshot
    assert json.loads(col3.get(ids=["i3"])["metadatas"][0]["filter_results_json"]) == {"old": "no", "is it X3?": "yes"}
    bad = mmiss_b200.SearchService(col3)                                            # no encoder: the reference's error shape
    bad.process_filter_on_all_images("other")
    assert bad.get_filter_progress("other") == {"status": "error", "message": "Model not available", "progress": 0}
