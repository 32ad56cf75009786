"""ctypes driver of oracle/hnsw_oracle.c -- TEST INFRASTRUCTURE (see the C file's header): a CPU
restatement of the hnswlib cosine index behind the reference's chromadb collection
(backend/app/utils.py:127-130), used to report recall@k of that approximate index against the exact
result.  Imported only by tests/, tools/ and bench.py's reference arm."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "hnsw_oracle.c")
LIB = os.path.join(_HERE, "_build", "libhnsw_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O3 (AVX2/FMA: any x86 server of the last decade) -> oracle/_build/libhnsw_oracle.so."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["gcc", "-O3", "-mavx2", "-mfma", "-ffast-math", "-shared", "-fPIC", "-o", LIB + ".tmp", SRC, "-lm"],
                   check=True)
    os.replace(LIB + ".tmp", LIB)
    return LIB


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.hnsw_create.restype = C.c_void_p
        lib.hnsw_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]
        lib.hnsw_destroy.argtypes = [C.c_void_p]
        lib.hnsw_count.argtypes = [C.c_void_p]
        lib.hnsw_add_many.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.hnsw_search_many.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


class HnswIndex:
    """chromadb defaults: M = 16 (``hnsw:M``), ef_construction = 100 (``hnsw:construction_ef``),
    search ef = 10 (``hnsw:search_ef``; newer releases default to 100)."""

    def __init__(self, dim: int, capacity: int, M: int = 16, ef_construction: int = 100, seed: int = 100):
        if not 2 <= M <= 64:
            raise ValueError("M must be in [2, 64]")
        self._lib = _load()
        self.dim, self.capacity = int(dim), int(capacity)
        self._h = self._lib.hnsw_create(self.dim, self.capacity, int(M), int(ef_construction), int(seed))
        if not self._h:
            raise MemoryError("hnsw_create failed")

    def __len__(self):
        return int(self._lib.hnsw_count(self._h))

    def add(self, X):
        X = np.ascontiguousarray(X, dtype=np.float32).reshape(-1, self.dim)
        done = self._lib.hnsw_add_many(self._h, X.ctypes.data, X.shape[0])
        if done != X.shape[0]:
            raise ValueError("index is full")

    def search(self, Q, k: int, ef: int = 10):
        """-> (ids [B,k] int32, cosine distances [B,k] float32), ascending distance; -1 / inf pad."""
        Q = np.ascontiguousarray(Q, dtype=np.float32).reshape(-1, self.dim)
        ids = np.empty((Q.shape[0], k), dtype=np.int32)
        dist = np.empty((Q.shape[0], k), dtype=np.float32)
        self._lib.hnsw_search_many(self._h, Q.ctypes.data, Q.shape[0], int(k), int(ef), ids.ctypes.data, dist.ctypes.data)
        return ids, dist

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hnsw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def recall_at_k(approx_ids: np.ndarray, exact_ids: np.ndarray) -> float:
    """mean over queries of |approx ∩ exact| / k."""
    k = exact_ids.shape[1]
    hit = sum(len(set(a.tolist()) & set(e.tolist())) for a, e in zip(approx_ids, exact_ids))
    return hit / (k * exact_ids.shape[0])


# ------------------------------------------------------------------------------------------------
# the number the spec asks for: recall@k of this index against the exact result, with its CPU rates
# (called by tests/ and by `bench.py --impl reference`)
# ------------------------------------------------------------------------------------------------
import time  # noqa: E402

from . import cosine_oracle as O  # noqa: E402


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
