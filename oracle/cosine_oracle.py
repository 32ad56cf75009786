"""CPU oracle for the exact-cosine top-k path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``multimodal-image-similarity-search_b200``) never imports it and has no CPU fallback.

PARITY STATUS
-------------
* The arithmetic of the reference's ``collection.query`` lives in the un-vendored,
  un-pinned third-party package ``chromadb`` (``/root/reference/requirements.txt:10``,
  ``chromadb>=0.4.13``), which wraps hnswlib (``space="cosine"``: vectors are L2-normalised
  at insert, distance = 1 - <q^, x^>).  Neither package is installable here, the reference
  holds no tests / golden vectors for it, so **the cosine top-k itself is "parity unpinned"**:
  it is restated from the published definition and cross-checked against float64.  The DEFINITION
  (cosine distance, ascending) is pinned against two independent third-party implementations that
  are installed -- ``scipy.spatial.distance.cdist(.., "cosine")`` and
  ``sklearn.neighbors.NearestNeighbors(metric="cosine", algorithm="brute")`` -- through
  ``tests/golden/make_thirdparty_golden.py`` -> ``thirdparty_golden.npz``
  (``tests/test_oracle.py::test_cosine_definition_matches_scipy_and_sklearn``); chromadb's own
  outputs remain unpinned.
* The parts of the path that ARE reference source -- the multimodal blend
  (``backend/app/main.py:850-860``), the distance->similarity map (``main.py:782``,
  ``app.py:326``), the limit rule (``main.py:757``) and the post-filter predicate
  (``main.py:201-222``) -- are pinned: ``tests/golden/make_reference_golden.py`` imports the
  real reference module (third-party imports stubbed) and records its outputs in
  ``tests/golden/reference_golden.json``; ``tests/test_oracle.py`` checks this file against
  them bit-for-bit.

Conventions shared with the CUDA engine:
* scores are float32 ``<q^, x> * inv_norm(x)`` with ``q^ = q / ||q||`` and
  ``inv_norm(x) = 1 / (||x|| + 1e-30)`` computed in float32 from the *stored* (possibly
  bf16-rounded) row; a zero row or zero query therefore scores 0 (distance 1), never NaN
  (hnswlib's cosine normalisation uses the same 1e-30 guard).
* ranking is by (score descending, row index ascending); distance = 1 - score.
* ``corpus_dtype="bf16"`` rounds the corpus to bfloat16 (round-to-nearest-even) first and
  then computes in float32 -- the oracle consumes the SAME rounded inputs as the engine.
  ``round_queries=True`` additionally rounds the *normalised* queries to bf16 (what the
  tensor-core batched kernel does).
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

__all__ = [
    "bf16_round", "inv_norms", "normalize_rows", "blend", "resolve_limit",
    "similarity_from_distance", "cosine_scores", "cosine_topk", "post_filter",
    "filter_mask", "pack_mask_bits", "dedup_pairs", "merge_topk", "topk_matches",
]


# --------------------------------------------------------------------------------------
# element-wise helpers
# --------------------------------------------------------------------------------------
def bf16_round(x: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 (round-to-nearest-even) -> float32, pure numpy bit arithmetic.

    Same rounding as CUDA ``__float2bfloat16_rn`` / ``torch.Tensor.to(torch.bfloat16)``.
    NaN stays NaN (quiet bit forced)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) >> 16) << 16
    out = r.astype(np.uint32)
    nan = np.isnan(x)
    if nan.any():
        out = np.where(nan, (x.view(np.uint32) | 0x00400000) & 0xFFFF0000, out).astype(np.uint32)
    return out.view(np.float32).reshape(x.shape)


def inv_norms(x: np.ndarray) -> np.ndarray:
    """1 / (||row||_2 + 1e-30) in float32, one per row (sum of squares accumulated in f64
    then rounded once, i.e. the correctly-rounded value the engine's f32 tree sum approximates)."""
    x = np.asarray(x, dtype=np.float32)
    ss = np.einsum("ij,ij->i", x.astype(np.float64), x.astype(np.float64))
    return (1.0 / (np.sqrt(ss) + 1e-30)).astype(np.float32)


def normalize_rows(x: np.ndarray) -> np.ndarray:
    """x / ||x||_2 per row.  Follows ``backend/app/utils.py:77-78, 97-98``
    (``features / features.norm(dim=1, keepdim=True)``; no epsilon there -- the 1e-30 guard
    only changes the all-zero row, which the reference would turn into NaN)."""
    x = np.asarray(x, dtype=np.float32)
    if x.ndim == 1:
        return (x * inv_norms(x[None])[0]).astype(np.float32)
    return (x * inv_norms(x)[:, None]).astype(np.float32)


def blend(image_embedding: np.ndarray, text_embedding: np.ndarray, weight_image: float) -> np.ndarray:
    """Multimodal query blend, statement-for-statement what
    ``backend/app/main.py:850-860`` (``search_multimodal``) does in numpy float32:
    normalise both, ``w*i^ + (1-w)*t^``, normalise the sum.  (No zero-norm guard in the
    backend; the legacy ``app.py:417-419`` guards it.  We keep the backend's arithmetic.)"""
    image_embedding = np.asarray(image_embedding, dtype=np.float32)
    text_embedding = np.asarray(text_embedding, dtype=np.float32)
    image_embedding_norm = image_embedding / np.linalg.norm(image_embedding)
    text_embedding_norm = text_embedding / np.linalg.norm(text_embedding)
    combined = (weight_image * image_embedding_norm + (1 - weight_image) * text_embedding_norm)
    combined = combined / np.linalg.norm(combined)
    return combined


def resolve_limit(limit: int) -> int:
    """``backend/app/main.py:757``: the UI's "All" (limit <= 0) means n_results = 1000."""
    return 1000 if limit <= 0 else limit


def similarity_from_distance(distances: Iterable[float], legacy: bool = False) -> List[float]:
    """``backend/app/main.py:782`` (``1 - d/2``) or legacy ``app.py:326`` (``1 - d``)."""
    if legacy:
        return [1.0 - distance for distance in distances]
    return [1 - (distance / 2) for distance in distances]


# --------------------------------------------------------------------------------------
# cosine scores / top-k  (restates chromadb+hnswlib cosine space; parity unpinned, see header)
# --------------------------------------------------------------------------------------
def _prepare(Q, X, corpus_dtype: str, round_queries: bool):
    X = np.asarray(X, dtype=np.float32)
    Q = np.asarray(Q, dtype=np.float32)
    if Q.ndim == 1:
        Q = Q[None]
    if corpus_dtype == "bf16":
        X = bf16_round(X)
    elif corpus_dtype != "f32":
        raise ValueError(corpus_dtype)
    Qn = normalize_rows(Q)
    if round_queries:
        Qn = bf16_round(Qn)
    return Qn, X


def cosine_scores(Q, X, corpus_dtype: str = "f32", round_queries: bool = False,
                  accumulate: str = "f32") -> np.ndarray:
    """[B, N] float32 cosine similarities under the engine's conventions.

    ``accumulate="f64"`` gives the float64-accumulated value (rounded to f32 at the end),
    used to bound the f32 summation-order error in the oracle self-tests."""
    Qn, Xr = _prepare(Q, X, corpus_dtype, round_queries)
    inv = inv_norms(Xr)
    if accumulate == "f64":
        dots = Qn.astype(np.float64) @ Xr.astype(np.float64).T
        return (dots * inv.astype(np.float64)[None, :]).astype(np.float32)
    dots = Qn @ Xr.T
    return (dots * inv[None, :]).astype(np.float32)


def cosine_topk(Q, X, k: int, corpus_dtype: str = "f32", round_queries: bool = False,
                valid: np.ndarray | None = None, block: int = 262144
                ) -> Tuple[np.ndarray, np.ndarray]:
    """Exact top-k by cosine similarity: returns (scores [B,k'] f32, rows [B,k'] int64),
    ordered by (score desc, row asc), k' = min(k, #valid rows)  (Chroma clamps n_results to
    the collection size).  ``valid`` is an optional boolean row mask ("pre" filter mode).
    Corpus is processed in row blocks so 10M-row cases do not materialise [B,N]."""
    Qn, Xr = _prepare(Q, X, corpus_dtype, round_queries)
    B, N = Qn.shape[0], Xr.shape[0]
    n_valid = N if valid is None else int(np.count_nonzero(valid))
    kk = min(k, n_valid)
    best_s = np.full((B, 0), -np.inf, dtype=np.float32)
    best_r = np.zeros((B, 0), dtype=np.int64)
    for lo in range(0, N, block):
        hi = min(N, lo + block)
        xb = Xr[lo:hi]
        s = ((Qn @ xb.T) * inv_norms(xb)[None, :]).astype(np.float32)
        rows = np.arange(lo, hi, dtype=np.int64)
        if valid is not None:
            keep = valid[lo:hi]
            s, rows = s[:, keep], rows[keep]
        cs = np.concatenate([best_s, s], axis=1)
        cr = np.concatenate([best_r, np.broadcast_to(rows, (B, rows.shape[0]))], axis=1)
        # lexsort: last key is primary.  (-score asc, row asc)
        order = np.stack([np.lexsort((cr[b], -cs[b].astype(np.float64)))[:kk] for b in range(B)])
        best_s = np.take_along_axis(cs, order, axis=1)
        best_r = np.take_along_axis(cr, order, axis=1)
    return best_s, best_r


def merge_topk(cand_scores: np.ndarray, cand_rows: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """k-way merge of per-shard candidates: [G,B,k] (global rows, <0 = empty slot) -> [B,k'].
    This is the step after the NCCL all-gather."""
    G, B, kk = cand_scores.shape
    cs = np.transpose(cand_scores, (1, 0, 2)).reshape(B, G * kk)
    cr = np.transpose(cand_rows, (1, 0, 2)).reshape(B, G * kk)
    out_s, out_r = [], []
    for b in range(B):
        ok = cr[b] >= 0
        s, r = cs[b][ok], cr[b][ok]
        order = np.lexsort((r, -s.astype(np.float64)))[:k]
        out_s.append(s[order]); out_r.append(r[order])
    n = min(len(x) for x in out_s) if out_s else 0
    return (np.stack([x[:n] for x in out_s]).astype(np.float32),
            np.stack([x[:n] for x in out_r]).astype(np.int64))


# --------------------------------------------------------------------------------------
# filter pass (reference semantics) and the new CLIP-cosine sweep / dedup
# --------------------------------------------------------------------------------------
def post_filter(results: Sequence[Dict], filters: Sequence[str]) -> List[Dict]:
    """The filter-application pass of the three search routes,
    ``backend/app/main.py:201-222`` == ``:257-278`` == ``:319-340``: keep a ranked result iff
    every selected filter's stored answer, lower-cased and stripped, equals "yes"; invalid
    or missing ``filter_results_json`` means no answers, so the result is dropped."""
    if not filters:
        return list(results)
    kept = []
    for r in results:
        filter_results = {}
        if "filter_results_json" in r:
            try:
                filter_results = json.loads(r["filter_results_json"])
            except (json.JSONDecodeError, TypeError):
                filter_results = {}
        if all(filter_results.get(f, "").lower().strip() == "yes" for f in filters):
            kept.append(r)
    return kept


def filter_mask(F, X, tau: float, corpus_dtype: str = "bf16", round_queries: bool = True) -> np.ndarray:
    """[F, N] bool: cos(prompt_f, row_n) >= tau.  New capability (BASELINE config 4), no
    reference arithmetic; defined on the same rounded inputs as the tensor-core kernel."""
    return cosine_scores(F, X, corpus_dtype, round_queries) >= np.float32(tau)


def pack_mask_bits(mask: np.ndarray) -> np.ndarray:
    """[F, N] bool -> [F, ceil(N/32)] uint32, bit (n % 32) of word n // 32 (little-endian bits)."""
    F, N = mask.shape
    W = (N + 31) // 32
    padded = np.zeros((F, W * 32), dtype=bool)
    padded[:, :N] = mask
    bits = padded.reshape(F, W, 32).astype(np.uint32)
    return (bits << np.arange(32, dtype=np.uint32)[None, None, :]).sum(axis=2, dtype=np.uint64).astype(np.uint32)


def dedup_pairs(X, tau: float, corpus_dtype: str = "bf16", block: int = 4096
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """All pairs i<j with cos(x_i, x_j) >= tau -> (i, j, score), sorted by (i, j).
    New capability (BASELINE config 5).  Both sides use the stored (rounded) rows; like the
    tensor-core kernel the dot is taken on raw stored rows and scaled by both inv-norms."""
    X = np.asarray(X, dtype=np.float32)
    Xr = bf16_round(X) if corpus_dtype == "bf16" else X
    inv = inv_norms(Xr)
    N = Xr.shape[0]
    I, J, S = [], [], []
    for lo in range(0, N, block):
        hi = min(N, lo + block)
        s = (Xr[lo:hi] @ Xr[lo:].T).astype(np.float32)
        s = (s * inv[lo:hi, None]).astype(np.float32) * inv[None, lo:]
        ii, jj = np.nonzero(s >= np.float32(tau))
        gi, gj = ii + lo, jj + lo
        keep = gi < gj
        I.append(gi[keep]); J.append(gj[keep]); S.append(s[ii[keep], jj[keep]])
    I = np.concatenate(I) if I else np.zeros(0, np.int64)
    J = np.concatenate(J) if J else np.zeros(0, np.int64)
    S = np.concatenate(S) if S else np.zeros(0, np.float32)
    order = np.lexsort((J, I))
    return I[order].astype(np.int64), J[order].astype(np.int64), S[order].astype(np.float32)


# --------------------------------------------------------------------------------------
# tie-aware comparator used by every parity test
# --------------------------------------------------------------------------------------
def topk_matches(got_scores, got_rows, ref_scores_full: np.ndarray, k: int, tol: float) -> Tuple[bool, str]:
    """Check one query's result against the oracle's FULL score vector.

    Passes iff (1) every returned score is within ``tol`` of the oracle's score for that row,
    (2) returned scores are non-increasing, (3) rows are distinct, (4) the id set equals the
    oracle's top-k "modulo ties within tolerance": every returned row's oracle score is
    >= (oracle k-th score - tol), and every row whose oracle score is > (oracle k-th + tol)
    is present."""
    got_scores = np.asarray(got_scores, dtype=np.float32)
    got_rows = np.asarray(got_rows, dtype=np.int64)
    n = ref_scores_full.shape[0]
    kk = min(k, n)
    if got_rows.shape[0] != kk:
        return False, f"expected {kk} results, got {got_rows.shape[0]}"
    if len(set(got_rows.tolist())) != kk:
        return False, "duplicate rows in result"
    if (got_rows < 0).any() or (got_rows >= n).any():
        return False, "row out of range"
    ref_at = ref_scores_full[got_rows]
    err = np.abs(ref_at - got_scores).max() if kk else 0.0
    if err > tol:
        return False, f"score error {err:.3e} > tol {tol:.1e}"
    if kk and (np.diff(got_scores) > 0).any():
        return False, "scores not sorted descending"
    if kk:
        kth = np.partition(ref_scores_full, n - kk)[n - kk]
        if (ref_at < kth - tol).any():
            return False, "returned a row below the oracle's k-th score (beyond tolerance)"
        must = np.nonzero(ref_scores_full > kth + tol)[0]
        missing = np.setdiff1d(must, got_rows)
        if missing.size:
            return False, f"missing rows clearly inside the top-k: {missing[:5].tolist()}"
    return True, "ok"
