"""-m gpu: parity of the tcgen05 paths (K2 batched top-k, K3 filter sweep, K4 dedup) against the
CPU oracle fed the SAME rounded inputs (bf16 corpus, queries normalised then rounded to bf16).
Tolerance: 2e-3 on scores (BASELINE.json, bf16), id sets identical modulo ties within tolerance;
threshold outputs may differ from the oracle only where the oracle's score is within 1e-4 of tau."""
import os

import numpy as np
import pytest

from oracle import cosine_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz")


def _check_topk(s, r, Q, X, k, valid=None):
    full = O.cosine_scores(Q, X, corpus_dtype="bf16", round_queries=True)
    if valid is not None:
        full = np.where(valid[None, :], full, -np.inf).astype(np.float32)
    for b in range(full.shape[0]):
        kk = min(k, int(np.isfinite(full[b]).sum()))
        got_r, got_s = r[b][r[b] >= 0], s[b][r[b] >= 0]
        assert len(got_r) == kk, (b, len(got_r), kk)
        ok, why = O.topk_matches(got_s, got_r, full[b], kk, 2e-3)
        assert ok, f"query {b}: {why}"
        # far tighter than 2e-3 in practice: fp32 accumulation; the residual ~1e-5 comes from query
        # components whose f32 normalised value sits on a bf16 rounding boundary (the device's and
        # numpy's 1/||q|| differ in the last ulp, flipping that component by one bf16 ulp)
        assert np.abs(full[b][got_r] - got_s).max() < 2e-4


@pytest.mark.parametrize("n,d,B,k", [(1000, 512, 16, 10), (70000, 512, 130, 10), (5000, 256, 300, 32),
                                     (3000, 768, 20, 5), (4096, 64, 128, 10), (2500, 200, 17, 10),
                                     (130, 512, 1024, 10),
                                     # dim > 512: the A block is streamed with every stage instead of resident
                                     (40000, 768, 300, 10), (6000, 1024, 130, 10), (3000, 1536, 40, 32), (2000, 576, 16, 10)])
def test_tensor_topk_matches_oracle(gpu, n, d, B, k):
    rng = np.random.default_rng(n + d + B)
    X = rng.standard_normal((n, d)).astype(np.float32) * rng.uniform(0.2, 3.0, (n, 1)).astype(np.float32)
    Q = rng.standard_normal((B, d)).astype(np.float32) * 2.0
    ix = gpu.DeviceIndex(d, "bf16")
    ix.add(X)
    s, r = ix.query(Q, k, mode="tensor")
    assert ix.last_query_path == "tensor"
    _check_topk(s, r, Q, X, k)
    ix.close()


def test_tensor_and_scan_paths_agree(gpu):
    """Same index, both kernels: identical id sets (modulo near-ties) and scores within the bf16
    query-rounding error."""
    rng = np.random.default_rng(11)
    X = rng.standard_normal((50000, 512)).astype(np.float32)
    Q = rng.standard_normal((64, 512)).astype(np.float32)
    ix = gpu.DeviceIndex(512, "bf16")
    ix.add(X)
    st, rt = ix.query(Q, 10, mode="tensor")
    ss, rs = ix.query(Q, 10, mode="scan")
    assert np.abs(st - ss).max() < 2e-3
    overlap = np.mean([len(set(rt[b]) & set(rs[b])) / 10 for b in range(64)])
    assert overlap > 0.9
    s_auto, r_auto = ix.query(Q, 10)                    # auto: B >= 16 on bf16 -> tensor
    assert ix.last_query_path == "tensor"
    np.testing.assert_array_equal(r_auto, rt)
    ix.close()


def test_tensor_golden_and_ties(gpu):
    g = np.load(GOLD)
    X, Q, k = g["X"], g["Q"], int(g["k"])
    Qb = np.tile(Q, (4, 1))                              # 20 queries
    ix = gpu.DeviceIndex(X.shape[1], "bf16")
    ix.add(X)
    s, r = ix.query(Qb, k, mode="tensor")
    np.testing.assert_allclose(s[:5], g["scores_bf16q"], atol=2e-4, rtol=0)
    _check_topk(s, r, Qb, X, k)
    assert set(r[1][:3].tolist()) == {3, 17, 400}
    assert r[1][:2].tolist() == [3, 17]                  # bit-identical rows: ranked by row index
    ix.close()


def test_tensor_pre_filter_bits(gpu):
    rng = np.random.default_rng(12)
    n, d = 4000, 128
    X = rng.standard_normal((n, d)).astype(np.float32)
    ix = gpu.DeviceIndex(d, "bf16")
    ix.add(X)
    has = rng.random(n) < 0.25
    for row in np.nonzero(has)[0]:
        ix.set_filter_bits(int(row), [5])
    Q = rng.standard_normal((40, d)).astype(np.float32)
    s, r = ix.query(Q, 10, require_bits=[5], mode="tensor")
    _check_topk(s, r, Q, X, 10, valid=has)
    ix.close()


@pytest.mark.parametrize("n,d,F,tau", [(300, 96, 6, 0.1), (10000, 512, 256, 0.08), (1100, 512, 3, 0.0),
                                       (5000, 768, 130, 0.05), (3000, 1024, 200, 0.04)])
def test_filter_sweep_matches_oracle(gpu, n, d, F, tau):
    if (n, d, F) == (300, 96, 6):
        g = np.load(GOLD)
        X, P = g["Xd"], g["F"]
    else:
        rng = np.random.default_rng(n + F)
        X = rng.standard_normal((n, d)).astype(np.float32)
        P = rng.standard_normal((F, d)).astype(np.float32)
    ix = gpu.DeviceIndex(d, "bf16")
    ix.add(X)
    bits = ix.filter_sweep(P, tau)
    assert bits.shape == (F, ix.filter_words())
    got = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :n].astype(bool)
    assert not np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, n:].any()   # tail bits clear
    scores = O.cosine_scores(P, X, "bf16", True)
    want = scores >= np.float32(tau)
    bad = got != want
    assert np.all(np.abs(scores[bad] - tau) < 1e-4), f"{bad.sum()} mismatches away from the threshold"
    assert 0.001 < want.mean() < 0.9
    ix.close()


@pytest.mark.parametrize("n,d", [(300, 96), (20000, 768), (9000, 512), (5000, 1024),
                                 (40000, 512)])      # resident-A pair mode, several work items per cluster
def test_dedup_matches_oracle(gpu, n, d):
    if (n, d) == (300, 96):
        g = np.load(GOLD)
        X, tau = g["Xd"], float(g["dedup_tau"])
    else:
        rng = np.random.default_rng(n)
        X = rng.standard_normal((n, d)).astype(np.float32)
        src = rng.choice(n // 2, size=200, replace=False)
        dst = n // 2 + rng.choice(n // 2, size=200, replace=False)
        X[dst] = X[src] + (0.1 / np.sqrt(d)) * rng.standard_normal((200, d)).astype(np.float32)   # cos ~ 0.995
        tau = 0.95
    ix = gpu.DeviceIndex(d, "bf16")
    ix.add(X)
    i, j, s = ix.dedup(tau)
    wi, wj, ws = O.dedup_pairs(X, tau)
    got = {(a, b): c for a, b, c in zip(i.tolist(), j.tolist(), s.tolist())}
    want = {(a, b): c for a, b, c in zip(wi.tolist(), wj.tolist(), ws.tolist())}
    for key in set(got) ^ set(want):
        sc = got.get(key, want.get(key))
        assert abs(sc - tau) < 1e-4, f"pair {key} score {sc} differs from the oracle away from tau"
    for key in set(got) & set(want):
        assert abs(got[key] - want[key]) < 2e-5
    assert len(want) >= 10
    # row-range form (how ranks split the triangle): union over ranges == whole
    parts = [ix.dedup(tau, row_lo=lo, row_hi=min(n, lo + 3000)) for lo in range(0, n, 3000)]
    assert sum(len(p[0]) for p in parts) == len(i)
    ix.close()


@pytest.mark.parametrize("n,d,B,k", [(30000, 512, 40, 50), (30000, 512, 100, 100), (9000, 768, 20, 128), (20000, 256, 17, 33),
                                     (90, 512, 16, 64),          # fewer rows than k: the later rounds run out of rows
                                     (50000, 512, 1100, 10),     # 9 query blocks: two A groups of 8 share each corpus slice
                                     (20000, 512, 2100, 10)])
def test_tensor_topk_rounds_and_a_groups(gpu, n, d, B, k):
    """32 < k <= 128 on the tensor path = ceil(k/32) exact rounds (round r only admits rows ranking after the last
    entry of round r-1), incl. duplicate rows whose tie is cut by the row number across a round boundary; B above
    one cluster's 1024 queries = several A groups per corpus slice in ONE launch."""
    rng = np.random.default_rng(n + d + B + k)
    X = rng.standard_normal((n, d)).astype(np.float32) * rng.uniform(0.2, 3.0, (n, 1)).astype(np.float32)
    Q = rng.standard_normal((B, d)).astype(np.float32)
    if n > 1000:
        X[5000:5040] = Q[0] * 0.7                                 # 40 exact duplicates straddle the 32-entry round boundary
        X[7000:7100:2] = Q[1] * 1.3
    ix = gpu.DeviceIndex(d, "bf16")
    ix.add(X)
    s, r = ix.query(Q, k, mode="tensor")
    assert ix.last_query_path == "tensor"
    _check_topk(s, r, Q, X, k)
    if n > 1000:
        assert r[0][:min(40, k)].tolist() == list(range(5000, 5000 + min(40, k)))      # ties in ascending row order, across rounds
        assert (np.diff(s, axis=1) <= 0).all()
    s2, r2 = ix.query(Q[:16], k, mode="auto")                      # auto: B >= 16 and k <= 128 -> tensor path
    assert ix.last_query_path == "tensor"
    np.testing.assert_array_equal(r2, r[:16])
    ix.close()


def test_tensor_rounds_with_pre_filter(gpu):
    rng = np.random.default_rng(4)
    n, d, B, k = 20000, 512, 24, 70
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((B, d)).astype(np.float32)
    ix = gpu.DeviceIndex(d, "bf16")
    ix.add(X)
    keep = rng.random(n) < 0.3
    ix.set_filter_bits_range(0, [[3] if kp else [] for kp in keep])
    s, r = ix.query(Q, k, require_bits=[3], mode="tensor")
    _check_topk(s, r, Q, X, k, valid=keep)
    ix.close()
