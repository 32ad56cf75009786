"""CPU: the hnswlib restatement (oracle/hnsw_oracle.c) behaves like an HNSW index should -- exact on
small sets when the beam covers them, high recall on clustered data at chromadb's defaults, monotone
in ef -- so the recall-vs-exact number reported by bench.py's reference arm means something."""
import numpy as np

from oracle import cosine_oracle as O
from oracle.hnsw_oracle import HnswIndex, recall_at_k


def _exact(Q, X, k):
    s, r = O.cosine_topk(Q, X, k)
    return r


def test_exact_when_beam_covers_everything():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((300, 32)).astype(np.float32) * rng.uniform(0.1, 5, (300, 1)).astype(np.float32)
    Q = rng.standard_normal((20, 32)).astype(np.float32)
    ix = HnswIndex(32, 300)
    ix.add(X)
    assert len(ix) == 300
    ids, dist = ix.search(Q, 10, ef=300)
    ex = _exact(Q, X, 10)
    assert recall_at_k(ids, ex) == 1.0
    full = O.cosine_scores(Q, X)
    for b in range(20):                                  # distances are 1 - cos, ascending
        np.testing.assert_allclose(dist[b], 1.0 - full[b][ids[b]], atol=2e-6)
        assert (np.diff(dist[b]) >= -1e-7).all()


def test_recall_on_clustered_data_at_chroma_defaults():
    rng = np.random.default_rng(1)
    centres = rng.standard_normal((40, 64)).astype(np.float32)
    X = (centres[rng.integers(0, 40, 6000)] + 0.35 * rng.standard_normal((6000, 64))).astype(np.float32)
    Q = (centres[rng.integers(0, 40, 100)] + 0.35 * rng.standard_normal((100, 64))).astype(np.float32)
    ix = HnswIndex(64, 6000, M=16, ef_construction=100)
    ix.add(X)
    ex = _exact(Q, X, 10)
    r10 = recall_at_k(ix.search(Q, 10, ef=10)[0], ex)
    r100 = recall_at_k(ix.search(Q, 10, ef=100)[0], ex)
    assert r100 >= r10 - 1e-9 and r100 > 0.9 and r10 > 0.5, (r10, r100)


def test_fewer_points_than_k_and_full_index():
    X = np.eye(4, 8, dtype=np.float32)
    ix = HnswIndex(8, 4)
    ix.add(X)
    ids, dist = ix.search(X[:1], 10, ef=10)
    assert ids[0][:4].tolist()[0] == 0 and (ids[0][4:] == -1).all() and np.isinf(dist[0][4:]).all()
    try:
        ix.add(X[:1])
        raise AssertionError("add to a full index must fail")
    except ValueError:
        pass


def test_measure_reports_recall_and_rates():
    from oracle import hnsw_oracle
    out = hnsw_oracle.measure(1500, 32, 40, 10, clustered=True)
    assert 0.5 < out["recall@10_ef10"] <= out["recall@10_ef100"] + 1e-9 <= 1.0 + 1e-9
    assert out["qps_ef10"] > 0 and out["build_rows_per_s"] > 0
