"""CPU: pin the oracle.  (1) against outputs of the REAL reference code recorded in
tests/golden/reference_golden.json (blend, score map, limit rule, filter pass, duplicate check);
(2) against float64 and torch for the parts whose arithmetic lives in the un-installable chromadb
("parity unpinned" -- see oracle/cosine_oracle.py header); (3) committed known-answer vectors."""
import json
import os

import numpy as np
import pytest

from oracle import cosine_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_golden.json")) as f:
    REF = json.load(f)
GOLD = np.load(os.path.join(HERE, "golden", "oracle_golden.npz"))


def _f32(bits):
    return np.asarray(bits, dtype=np.uint32).view(np.float32)


# ---- (1) reference-pinned -----------------------------------------------------------------
@pytest.mark.parametrize("case", REF["blend"], ids=lambda c: f"w={c['weight_image']}")
def test_blend_is_bit_identical_to_reference(case):
    got = O.blend(_f32(case["image_bits"]), _f32(case["text_bits"]), case["weight_image"])
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got.view(np.uint32), np.asarray(case["sent_bits"], dtype=np.uint32))
    assert case["sent_is_f32_exact"] and case["include"] == ["metadatas", "distances"]


@pytest.mark.parametrize("case", REF["search_similar"], ids=lambda c: f"limit={c['limit']}")
def test_limit_rule_and_score_map_match_reference(case):
    assert O.resolve_limit(case["limit"]) == case["n_results"]
    sims = O.similarity_from_distance(case["distances"])
    assert [r["similarity_score"] for r in case["results"]] == sims
    assert case["query_kwargs"] == ["include", "n_results", "query_embeddings"]


@pytest.mark.parametrize("case", REF["filter_pass"], ids=lambda c: "+".join(c["filters"]) or "none")
def test_post_filter_matches_reference(case):
    ranked = [dict(m, similarity_score=1 - d / 2) for m, d in zip(case["input_metadatas"], case["distances"])]
    kept = O.post_filter(ranked, case["filters"])
    assert [r["id"] for r in kept] == case["kept_ids"]


def test_reference_duplicate_check_is_an_id_lookup():
    (c,) = REF["duplicate_check"]
    assert c["is_new"] is False and c["calls"] == ["get"]
    assert c["get_kwargs"] == {"ids": [c["existing_id"]], "include": ["metadatas"]}


# ---- (2) independent cross-checks ------------------------------------------------------------
def test_bf16_round_matches_torch():
    import torch
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(100000).astype(np.float32) * 10 ** rng.uniform(-20, 20, 100000).astype(np.float32),
                        np.array([0.0, -0.0, np.inf, -np.inf, 1.0, 1.00390625, 1.01171875, 3.3895314e38], np.float32)])
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(O.bf16_round(x).view(np.uint32), want.view(np.uint32))
    assert np.isnan(O.bf16_round(np.array([np.nan], np.float32)))[0]


@pytest.mark.parametrize("cd,rq", [("f32", False), ("bf16", False), ("bf16", True)])
def test_scores_match_float64(cd, rq):
    rng = np.random.default_rng(1)
    X = rng.standard_normal((2000, 512)).astype(np.float32)
    Q = rng.standard_normal((3, 512)).astype(np.float32)
    a = O.cosine_scores(Q, X, cd, rq)
    b = O.cosine_scores(Q, X, cd, rq, accumulate="f64")
    assert np.abs(a - b).max() < 1e-6
    # and against the textbook definition on the same (rounded) inputs
    Xr = O.bf16_round(X) if cd == "bf16" else X
    Qn = Q.astype(np.float64) / np.linalg.norm(Q.astype(np.float64), axis=1, keepdims=True)
    if rq:
        Qn = O.bf16_round(O.normalize_rows(Q)).astype(np.float64)
    ref = (Qn @ Xr.astype(np.float64).T) / np.linalg.norm(Xr.astype(np.float64), axis=1)[None]
    assert np.abs(a - ref).max() < 1e-6


def test_topk_blocked_equals_full_sort_and_ties_by_row():
    rng = np.random.default_rng(2)
    X = rng.standard_normal((5000, 64)).astype(np.float32)
    X[100] = X[7]
    X[4000] = X[7]
    Q = np.stack([X[7], rng.standard_normal(64).astype(np.float32)])
    s, r = O.cosine_topk(Q, X, 20, block=777)
    full = O.cosine_scores(Q, X)
    for b in range(2):
        # BLAS blocks differently for different row counts: scores may differ in the last ulp
        ok, why = O.topk_matches(s[b], r[b], full[b], 20, 1e-6)
        assert ok, why
    assert sorted(r[0][:3].tolist()) == [7, 100, 4000]
    Xs = np.tile(X[7], (50, 1))                      # bit-identical rows -> bit-identical scores
    _, rs = O.cosine_topk(Q[:1], Xs, 20, block=16)
    assert rs[0].tolist() == list(range(20))
    s2, r2 = O.cosine_topk(Q, X[:5], 20)
    assert s2.shape == (2, 5)                      # clamped to the collection size


def test_merge_topk_equals_global_topk():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((999, 32)).astype(np.float32)
    Q = rng.standard_normal((4, 32)).astype(np.float32)
    G, k = 4, 10
    per = (999 + G - 1) // G
    cs = np.full((G, 4, k), -np.inf, np.float32)
    cr = np.full((G, 4, k), -1, np.int64)
    for g in range(G):
        s, r = O.cosine_topk(Q, X[g * per:(g + 1) * per], k)
        cs[g, :, :s.shape[1]], cr[g, :, :r.shape[1]] = s, r + g * per
    ms, mr = O.merge_topk(cs, cr, k)
    full = O.cosine_scores(Q, X)
    for b in range(4):      # shard-sized and full-sized BLAS calls may differ in the last ulp
        ok, why = O.topk_matches(ms[b], mr[b], full[b], k, 1e-6)
        assert ok, why
        # and the merge itself is exact on the candidates it was given
        flat_s, flat_r = cs[:, b].ravel(), cr[:, b].ravel()
        order = np.lexsort((flat_r, -flat_s.astype(np.float64)))[:k]
        np.testing.assert_array_equal(mr[b], flat_r[order])


def test_comparator_accepts_ties_rejects_errors():
    full = np.array([0.9, 0.5, 0.5000001, 0.1, -0.3], np.float32)
    ok, _ = O.topk_matches([0.9, 0.5], [0, 1], full, 2, 1e-5)
    assert ok
    ok, _ = O.topk_matches([0.9, 0.5000001], [0, 2], full, 2, 1e-5)
    assert ok                                        # tie within tolerance: either id acceptable
    assert not O.topk_matches([0.9, 0.1], [0, 3], full, 2, 1e-5)[0]
    assert not O.topk_matches([0.9, 0.6], [0, 1], full, 2, 1e-5)[0]
    assert not O.topk_matches([0.5, 0.9], [1, 0], full, 2, 1e-5)[0]
    assert not O.topk_matches([0.9, 0.9], [0, 0], full, 2, 1e-5)[0]


def test_pack_mask_bits_layout():
    m = np.zeros((2, 70), bool)
    m[0, [0, 31, 32, 69]] = True
    m[1, 33] = True
    p = O.pack_mask_bits(m)
    assert p.shape == (2, 3)
    assert p[0].tolist() == [1 | (1 << 31), 1, 1 << 5] and p[1].tolist() == [0, 2, 0]


def test_dedup_pairs_bruteforce():
    rng = np.random.default_rng(4)
    X = rng.standard_normal((200, 48)).astype(np.float32)
    X[150:160] = X[0:10] + 0.02 * rng.standard_normal((10, 48)).astype(np.float32)
    i, j, s = O.dedup_pairs(X, 0.95, block=64)
    S = O.cosine_scores(X, X, "bf16", False)          # queries unrounded-normalised: close enough to list pairs
    want = {(a, b) for a in range(200) for b in range(a + 1, 200) if S[a, b] >= 0.951}
    got = set(zip(i.tolist(), j.tolist()))
    assert want <= got and all(b - a == 150 for a, b in got)


# ---- (3) committed known answers ------------------------------------------------------------------
@pytest.mark.parametrize("tag,cd,rq", [("f32", "f32", False), ("bf16", "bf16", False), ("bf16q", "bf16", True)])
def test_oracle_reproduces_committed_golden(tag, cd, rq):
    s, r = O.cosine_topk(GOLD["Q"], GOLD["X"], int(GOLD["k"]), corpus_dtype=cd, round_queries=rq)
    np.testing.assert_allclose(s, GOLD[f"scores_{tag}"], atol=1e-6, rtol=0)
    full = O.cosine_scores(GOLD["Q"], GOLD["X"], cd, rq)
    for b in range(s.shape[0]):                       # the committed rows are a valid top-k too
        ok, why = O.topk_matches(GOLD[f"scores_{tag}"][b], GOLD[f"rows_{tag}"][b], full[b], int(GOLD["k"]), 2e-6)
        assert ok, why
    assert sorted(r[1][:3].tolist()) == [3, 17, 400]  # the planted ties (cos = 1)


# ---- (4) third-party pin of the cosine-space definition (utils.py:129) ---------------------------------
def test_cosine_definition_matches_scipy_and_sklearn():
    """scipy.spatial.distance.cdist(..., "cosine") and sklearn NearestNeighbors(metric="cosine", brute) on the golden
    corpus (fixtures: tests/golden/make_thirdparty_golden.py).  chromadb itself stays unpinned (not installable);
    this pins the DEFINITION it implements against two independent implementations."""
    tp = np.load(os.path.join(HERE, "golden", "thirdparty_golden.npz"))
    X, Q, k = GOLD["X"], GOLD["Q"], int(GOLD["k"])
    keep = tp["keep"]                                                     # all rows but the zero row
    scores = O.cosine_scores(Q, X, corpus_dtype="f32")                    # [B, N] float32 cosine similarities
    np.testing.assert_allclose(1.0 - scores[:, keep].astype(np.float64), tp["cdist_cosine"], atol=2e-6)
    s, r = O.cosine_topk(Q, X, k, corpus_dtype="f32")
    zero_row = int(np.setdiff1d(np.arange(X.shape[0]), keep)[0])
    dist_of = np.full((Q.shape[0], X.shape[0]), np.inf)
    dist_of[:, keep] = tp["cdist_cosine"]
    for b in range(Q.shape[0]):
        assert zero_row not in r[b]                                       # a zero row scores 0: never in these top-10s
        np.testing.assert_allclose(1.0 - s[b].astype(np.float64), tp["sk_dist"][b], atol=2e-6)
        # same neighbours: identical id sets, or ids whose scipy distances agree within rounding (exact ties)
        want = tp["sk_rows"][b]
        if set(r[b].tolist()) != set(want.tolist()):
            np.testing.assert_allclose(np.sort(dist_of[b][r[b]]), np.sort(dist_of[b][want]), atol=2e-6)
    # query 1 is parallel to rows 3, 17 (an exact duplicate) and 400 (2.5 x row 3): three neighbours at distance 0;
    # sklearn orders such ties arbitrarily, the oracle (and the engine) by ascending row
    top3 = r[1][:3].tolist()                    # (row 400's f32 score may differ from the duplicates' by one ulp)
    assert set(top3) == {3, 17, 400} and top3.index(3) < top3.index(17) and np.abs(tp["sk_dist"][1][:3]).max() < 1e-9
    assert set(tp["sk_rows"][1][:3].tolist()) == {3, 17, 400}
