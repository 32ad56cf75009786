"""Randomised tests (hypothesis).

CPU: a random sequence of the Collection operations the reference uses (add / update / delete / get /
query, backend/app/main.py:735-740, 503-510, 1069, 631-634, 761-765) against a brute-force model --
exercises the move-last-row bookkeeping, id skipping, metadata merge and n_results clamping.
GPU (-m gpu): random shapes of the scan kernel (n, dim, k, norms, dtype) against the oracle.
"""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from mmiss_b200 import collection as C
from oracle import cosine_oracle as O
from tests.fake_index import FakeIndex

DIM = 12
_vec = st.lists(st.floats(-3, 3, allow_nan=False, width=32), min_size=DIM, max_size=DIM).filter(
    lambda v: sum(abs(x) for x in v) > 1e-3)
_id = st.integers(0, 25).map(lambda i: f"img_{i:02x}")
_op = st.one_of(
    st.tuples(st.just("add"), _id, _vec, st.sampled_from(["a.jpg", "b.jpg", "c.jpg"])),
    st.tuples(st.just("delete"), _id),
    st.tuples(st.just("update"), _id, st.sampled_from(["x", "y"])),
    st.tuples(st.just("query"), _vec, st.integers(1, 40)),
    st.tuples(st.just("get"), _id),
)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(st.lists(_op, min_size=1, max_size=40))
def test_collection_matches_bruteforce_model(monkeypatch, ops):
    monkeypatch.setattr(C, "DeviceIndex", FakeIndex)
    col = C.Collection("c", {"hnsw:space": "cosine"})
    model = {}                                              # id -> (vec, meta); insertion order irrelevant
    for op in ops:
        if op[0] == "add":
            _, id_, v, fn = op
            col.add(ids=[id_], embeddings=[v], metadatas=[{"filename": fn}], documents=["d"])
            model.setdefault(id_, (np.asarray(v, np.float32), {"filename": fn}))      # existing id: skipped
        elif op[0] == "delete":
            col.delete(ids=[op[1]])
            model.pop(op[1], None)
        elif op[0] == "update":
            col.update(ids=[op[1]], metadatas=[{"description": op[2]}])
            if op[1] in model:
                model[op[1]][1]["description"] = op[2]                                  # merged, not replaced
        elif op[0] == "get":
            got = col.get(ids=[op[1]], include=["metadatas"])
            if op[1] in model:
                assert got["ids"] == [op[1]] and got["metadatas"] == [model[op[1]][1]]
            else:
                assert got["ids"] == [] and got["metadatas"] == []
        else:
            _, q, n = op
            res = col.query(query_embeddings=[q], n_results=n, include=["metadatas", "distances"])
            assert len(res["ids"]) == 1 and len(res["ids"][0]) == min(n, len(model))
            if model:
                ids = list(model)
                X = np.stack([model[i][0] for i in ids])
                d = 1.0 - O.cosine_scores(np.asarray([q], np.float32), X)[0]
                want = sorted(d.tolist())[:len(res["ids"][0])]
                np.testing.assert_allclose(res["distances"][0], want, atol=2e-6)
                for id_, dist_, meta in zip(res["ids"][0], res["distances"][0], res["metadatas"][0]):
                    assert abs(d[ids.index(id_)] - dist_) < 2e-6 and meta == model[id_][1]
        assert col.count() == len(model)
    assert sorted(col.get(include=[])["ids"]) == sorted(model)


@pytest.mark.gpu
@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(1, 4000), d8=st.integers(1, 96), k=st.integers(1, 140), dtype=st.sampled_from(["f32", "bf16"]),
       B=st.integers(1, 5), seed=st.integers(0, 2**31 - 1))
def test_scan_random_shapes_match_oracle(gpu, n, d8, k, dtype, B, seed):
    d = d8 * 8 if seed % 3 else max(1, d8 * 8 - seed % 7)     # mostly multiples of 8, sometimes ragged
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal((n, d)) * rng.uniform(0.01, 50.0, (n, 1))).astype(np.float32)
    Q = (rng.standard_normal((B, d)) * rng.uniform(0.1, 10.0, (B, 1))).astype(np.float32)
    ix = gpu.DeviceIndex(d, dtype)
    ix.add(X)
    s, r = ix.query(Q, k, mode="scan")
    full = O.cosine_scores(Q, X, corpus_dtype=dtype)
    kk = min(k, n)
    for b in range(B):
        ok, why = O.topk_matches(s[b][:kk], r[b][:kk], full[b], kk, 1e-5 if dtype == "f32" else 2e-3)
        assert ok, why
        assert (r[b][kk:] == -1).all()
    ix.close()
