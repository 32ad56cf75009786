"""-m gpu: the chromadb-shaped Collection / SearchService on the real CUDA engine, end to end through
ctypes -> C ABI -> kernels, checked against the oracle and against the golden vectors recorded from
the reference's own code (tests/golden/reference_golden.json)."""
import json
import os
import threading

import numpy as np
import pytest

from oracle import cosine_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_golden.json")) as f:
    REF = json.load(f)


def _f32(bits):
    return np.asarray(bits, dtype=np.uint32).view(np.float32)


def test_config1_shape_six_images_top5(gpu):
    """BASELINE config 1 in shape: 6 rows of 768-d (LongCLIP ViT-L/14 width), text query, top-5, the
    exact-scan regime the reference itself is in below chromadb's 100-row HNSW batch."""
    rng = np.random.default_rng(1)
    X = O.normalize_rows(rng.standard_normal((6, 768)).astype(np.float32))
    q = O.normalize_rows(rng.standard_normal(768).astype(np.float32))
    col = gpu.Collection("image-match", {"hnsw:space": "cosine"}, dtype="f32")
    for i in range(6):
        col.add(ids=[f"img_{i:016x}"], embeddings=[X[i].tolist()], metadatas=[{"id": f"img_{i:016x}", "filename": f"drill{i}.jpg"}],
                documents=[f"a drill {i}"])
    svc = gpu.SearchService(col, encoder=lambda image=None, text=None: {"text": q[None]})
    res = svc.search_by_text("red drill", limit=5)
    full = O.cosine_scores(q, X)[0]
    order = np.argsort(-full, kind="stable")[:5]
    assert [r["id"] for r in res] == [f"img_{i:016x}" for i in order]
    np.testing.assert_allclose([r["similarity_score"] for r in res], (1 + full[order]) / 2, atol=1e-6)
    assert res[0]["url"] == f"/static/processed/img_{order[0]:016x}.png"
    col.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_collection_query_matches_oracle(gpu, dtype):
    rng = np.random.default_rng(2)
    n, d = 3000, 512
    X = rng.standard_normal((n, d)).astype(np.float32)
    col = gpu.Collection("c", dtype=dtype)
    ids = [f"id{i}" for i in range(n)]
    col.add(ids=ids, embeddings=X, metadatas=[{"n": i} for i in range(n)])
    Q = rng.standard_normal((3, d)).astype(np.float32)
    res = col.query(query_embeddings=Q.tolist(), n_results=10, include=["metadatas", "distances"])
    full = O.cosine_scores(Q, X, corpus_dtype=dtype)
    for b in range(3):
        rows = np.array([m["n"] for m in res["metadatas"][b]])
        assert res["ids"][b] == [ids[r] for r in rows]
        ok, why = O.topk_matches(1.0 - np.array(res["distances"][b], np.float32), rows, full[b], 10,
                                 1e-5 if dtype == "f32" else 2e-3)
        assert ok, why
    # batched through the tensor path (bf16, B >= 16) returns the same shapes
    if dtype == "bf16":
        res = col.query(query_embeddings=np.tile(Q, (8, 1)), n_results=10, include=["distances"])
        assert col.index.last_query_path == "tensor" and len(res["ids"]) == 24 and len(res["ids"][0]) == 10
    col.close()


@pytest.mark.parametrize("case", REF["blend"], ids=lambda c: f"w={c['weight_image']}")
def test_device_blend_matches_reference_vector(gpu, case):
    """The vector the reference hands to collection.query after its numpy blend (main.py:850-860),
    vs our device blend: same ranking and distances on a corpus (scores within 1e-5)."""
    img, txt, w = _f32(case["image_bits"]), _f32(case["text_bits"]), case["weight_image"]
    sent = _f32(case["sent_bits"])
    rng = np.random.default_rng(5)
    X = rng.standard_normal((2000, REF["dim"])).astype(np.float32)
    col = gpu.Collection("c", dtype="f32")
    col.add(ids=[str(i) for i in range(2000)], embeddings=X)
    a = col.query_multimodal([img.tolist()], [txt.tolist()], weight_image=w, n_results=10, include=["distances"])
    b = col.query(query_embeddings=[sent.tolist()], n_results=10, include=["distances"])
    np.testing.assert_allclose(a["distances"][0], b["distances"][0], atol=1e-5)
    full = O.cosine_scores(sent, X)[0]
    ok, why = O.topk_matches(1.0 - np.array(a["distances"][0], np.float32), np.array(a["ids"][0], np.int64), full, 10, 1e-5)
    assert ok, why
    col.close()


def test_filter_pass_golden_on_device(gpu):
    """Reference filter pass (main.py:201-222) replayed: post mode == reference, pre mode == in-kernel."""
    col = gpu.Collection("c", dtype="f32")
    rng = np.random.default_rng(0)
    q = O.normalize_rows(rng.standard_normal(64).astype(np.float32))
    fc0 = REF["filter_pass"][0]
    X = []
    for d in fc0["distances"]:
        c = 1.0 - d
        o = rng.standard_normal(64).astype(np.float32)
        o -= o.dot(q) * q
        o /= np.linalg.norm(o)
        X.append(c * q + np.sqrt(max(0.0, 1 - c * c)) * o)
    col.add(ids=[m["id"] for m in fc0["input_metadatas"]], embeddings=np.stack(X), metadatas=fc0["input_metadatas"])
    svc = gpu.SearchService(col, encoder=lambda image=None, text=None: {"text": q[None]})
    for fc in REF["filter_pass"]:
        out = svc.route_search_text("drill", filters=fc["filters"], limit=10)
        assert [r["id"] for r in out["results"]] == fc["kept_ids"]
        np.testing.assert_allclose([r["similarity_score"] for r in out["results"]],
                                   [r["similarity_score"] for r in fc["results"]], atol=1e-6)
        pre = col.query(query_embeddings=[q], n_results=10, where_filters=fc["filters"], filter_mode="pre")
        assert pre["ids"][0] == fc["kept_ids"]
    col.close()


def test_duplicate_check_update_delete_and_persistence(gpu, tmp_path):
    client = gpu.PersistentClient(path=str(tmp_path), dtype="bf16")
    names = client.list_collections()
    col = client.get_collection("image-match") if "image-match" in names else \
        client.create_collection(name="image-match", metadata={"hnsw:space": "cosine"})
    svc = gpu.SearchService(col)
    rng = np.random.default_rng(7)
    X = rng.standard_normal((50, 768)).astype(np.float32)
    for i in range(50):
        meta, is_new = svc.add_embedding(f"img_{i:016x}", X[i], {"id": f"img_{i:016x}", "filename": f"{i}.png"}, f"d{i}")
        assert is_new
    meta, is_new = svc.add_embedding(f"img_{7:016x}", X[8], {"id": "other"}, "dup")
    assert not is_new and meta["filename"] == "7.png" and col.count() == 50
    col.update(ids=[f"img_{7:016x}"], metadatas=[{"description": "edited"}])
    col.delete(ids=[f"img_{i:016x}" for i in range(10, 20)])
    assert col.count() == 40
    res = col.query(query_embeddings=[X[49].tolist()], n_results=1, include=["metadatas", "distances"])
    assert res["ids"][0] == [f"img_{49:016x}"] and res["distances"][0][0] < 1e-3
    col.close()
    col2 = gpu.PersistentClient(path=str(tmp_path)).get_collection("image-match")
    assert col2.count() == 40 and col2.get(ids=[f"img_{7:016x}"])["metadatas"][0]["description"] == "edited"
    res = col2.query(query_embeddings=[X[25].tolist()], n_results=3, include=["distances"])
    assert res["ids"][0][0] == f"img_{25:016x}"
    col2.close()


def test_concurrent_update_and_query(gpu):
    col = gpu.Collection("c", dtype="f32")
    X = np.random.default_rng(9).standard_normal((300, 128)).astype(np.float32)
    col.add(ids=[f"i{i}" for i in range(300)], embeddings=X, metadatas=[{"n": i} for i in range(300)])
    errors = []

    def worker():
        try:
            for i in range(300):
                col.update(ids=[f"i{i}"], metadatas=[{"filter_results_json": json.dumps({"f": "yes" if i % 3 == 0 else "no"})}])
        except Exception as e:                       # pragma: no cover
            errors.append(e)
    t = threading.Thread(target=worker)
    t.start()
    for i in range(150):
        assert col.query(query_embeddings=[X[i]], n_results=2)["ids"][0][0] == f"i{i}"
    t.join()
    assert not errors
    got = col.query(query_embeddings=[X[0]], n_results=300, where_filters=["f"], filter_mode="pre")["ids"][0]
    assert sorted(got) == sorted(f"i{i}" for i in range(0, 300, 3))
    col.close()


def test_filter_sweep_and_find_duplicates_on_collection(gpu):
    g = np.load(os.path.join(HERE, "golden", "oracle_golden.npz"))
    Xd = g["Xd"]
    col = gpu.Collection("c", dtype="bf16")
    col.add(ids=[f"r{i}" for i in range(Xd.shape[0])], embeddings=Xd)
    mask = col.filter_sweep(g["F"], float(g["filter_tau"]))
    scores = O.cosine_scores(g["F"], Xd, "bf16", True)
    bad = mask != g["filter_mask"]
    assert np.all(np.abs(scores[bad] - float(g["filter_tau"])) < 1e-4)
    n_yes = col.apply_filter_sweep("is drill?", g["F"][0], float(g["filter_tau"]))
    kept = col.query(query_embeddings=[Xd[0]], n_results=300, where_filters=["is drill?"], filter_mode="pre")
    assert len(kept["ids"][0]) == n_yes
    dups = col.find_duplicates(float(g["dedup_tau"]))
    assert {(a, b) for a, b, _ in dups} == {(f"r{i}", f"r{j}") for i, j in zip(g["dedup_i"], g["dedup_j"])}
    col.close()


def test_device_resident_add_and_query(gpu, tmp_path):
    """SURVEY 8(f2/f4): embeddings produced on the GPU (CLIP output) are ingested and queried without
    a host round trip; results equal the host-list path (main.py:735-740, 761-765), and the
    persistent collection still logs the rows."""
    import torch
    rng = np.random.default_rng(12)
    X = rng.standard_normal((500, 512)).astype(np.float32)
    Q = rng.standard_normal((20, 512)).astype(np.float32)
    ids = [f"img_{i:04x}" for i in range(500)]
    host = gpu.Collection("h", {"hnsw:space": "cosine"}, dtype="bf16")
    host.add(ids=ids, embeddings=X.tolist(), metadatas=[{"filename": f"{i}.jpg"} for i in range(500)])
    client = gpu.PersistentClient(path=str(tmp_path), dtype="bf16")
    dev = client.create_collection("d", metadata={"hnsw:space": "cosine"})
    dev.add(ids=ids, embeddings=torch.from_numpy(X).cuda(), metadatas=[{"filename": f"{i}.jpg"} for i in range(500)])
    dev.add(ids=ids[:3], embeddings=torch.from_numpy(X[:3]).cuda())       # existing ids are skipped
    assert dev.count() == 500
    a = host.query(query_embeddings=Q.tolist(), n_results=10, include=["metadatas", "distances"])
    b = dev.query(query_embeddings=torch.from_numpy(Q).cuda(), n_results=10, include=["metadatas", "distances"])
    assert a["ids"] == b["ids"] and a["metadatas"] == b["metadatas"]
    np.testing.assert_allclose(np.array(a["distances"]), np.array(b["distances"]), atol=2e-3)
    mb = gpu.MicroBatcher(dev, max_batch=32, max_wait_ms=100.0)
    got = [None] * 20
    ts = [threading.Thread(target=lambda i=i: got.__setitem__(i, mb.query(Q[i], 5))) for i in range(20)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    mb.close()
    for i in range(20):
        assert got[i]["ids"][0] == a["ids"][i][:5]
    dev.close()
    again = gpu.PersistentClient(path=str(tmp_path), dtype="bf16").get_collection("d")   # rows were logged
    assert again.count() == 500
    c = again.query(query_embeddings=Q[:2].tolist(), n_results=10)
    assert c["ids"] == a["ids"][:2]
    again.close()
    host.close()
