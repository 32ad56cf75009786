"""Recall@k of the reference's approximate index (chromadb = hnswlib, cosine space, restated in
oracle/hnsw_oracle.c) against the EXACT result -- the result the CUDA engine returns -- on synthetic
unit-norm CLIP-dimension vectors, plus the index's CPU build and query rates.  CPU only.

    python tools/hnsw_recall.py [--rows 100000] [--dim 512] [--queries 200] [--k 10] [--clustered]

BASELINE.json north_star: "Recall@k of the reference's ChromaDB HNSW index against the exact result
is also reported".  chromadb itself is not installable offline; see the C file's header for what is
restated.  Uniform random vectors in 512-d are the worst case for any graph index (all distances
concentrate); --clustered draws CLIP-like data (a mixture of directions) as a second point.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cosine_oracle as O  # noqa: E402
from oracle.hnsw_oracle import HnswIndex, recall_at_k  # noqa: E402


def synth(rows, dim, queries, clustered, seed=0):
    rng = np.random.default_rng(seed)
    if not clustered:
        X = rng.standard_normal((rows, dim), dtype=np.float32)
        Q = rng.standard_normal((queries, dim), dtype=np.float32)
    else:
        nc = max(16, rows // 500)
        c = rng.standard_normal((nc, dim), dtype=np.float32)
        X = c[rng.integers(0, nc, rows)] + 0.6 * rng.standard_normal((rows, dim), dtype=np.float32)
        Q = c[rng.integers(0, nc, queries)] + 0.6 * rng.standard_normal((queries, dim), dtype=np.float32)
    return O.normalize_rows(X), O.normalize_rows(Q)


def measure(rows, dim, queries, k, clustered, efs=(10, 100), M=16, efc=100):
    X, Q = synth(rows, dim, queries, clustered)
    t0 = time.perf_counter()
    ix = HnswIndex(dim, rows, M=M, ef_construction=efc)
    ix.add(X)
    t_build = time.perf_counter() - t0
    exact = np.stack([np.argsort(-(X @ q), kind="stable")[:k] for q in Q])
    out = {"index": f"HNSW M={M} ef_construction={efc} (chromadb defaults), cosine, single thread",
           "rows": rows, "dim": dim, "queries": queries, "k": k,
           "data": "clustered unit-norm" if clustered else "uniform random unit-norm",
           "build_s": round(t_build, 2), "build_rows_per_s": round(rows / t_build, 1)}
    for ef in efs:
        t0 = time.perf_counter()
        ids, _ = ix.search(Q, k, ef=ef)
        dt = time.perf_counter() - t0
        out[f"recall@{k}_ef{ef}"] = round(recall_at_k(ids, exact), 4)
        out[f"qps_ef{ef}"] = round(queries / dt, 1)
    ix.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--queries", type=int, default=200)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--clustered", action="store_true")
    a = ap.parse_args()
    print(json.dumps(measure(a.rows, a.dim, a.queries, a.k, a.clustered)), flush=True)


if __name__ == "__main__":
    main()
