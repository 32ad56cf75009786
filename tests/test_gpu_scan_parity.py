"""-m gpu: parity of the CUDA scan path (K1/K5/K6 + select) against the CPU oracle, through the
C ABI (ctypes).  Tolerances are BASELINE.json's: scores within 1e-5 (fp32 storage) / 2e-3 (bf16
storage), identical top-k id sets modulo ties within tolerance."""
import os

import numpy as np
import pytest

from oracle import cosine_oracle as O

pytestmark = pytest.mark.gpu
TOL = {"f32": 1e-5, "bf16": 2e-3}
GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz")


def _check(s, r, Q, X, k, dtype, valid=None, tol=None):
    full = O.cosine_scores(Q, X, corpus_dtype=dtype)
    if valid is not None:
        full = np.where(valid[None, :], full, -np.inf).astype(np.float32)
    for b in range(full.shape[0]):
        kk = min(k, int(np.isfinite(full[b]).sum()) if valid is not None else X.shape[0])
        got_r, got_s = r[b][r[b] >= 0], s[b][r[b] >= 0]
        assert len(got_r) == kk, (b, len(got_r), kk)
        assert (r[b][kk:] == -1).all() and np.isneginf(s[b][kk:]).all()
        ok, why = O.topk_matches(got_s, got_r, full[b], kk, tol or TOL[dtype])
        assert ok, f"query {b}: {why}"


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("n,d,k", [(6, 768, 5), (1000, 512, 10), (4097, 512, 25), (20000, 768, 10),
                                   (3000, 64, 50), (2500, 1024, 100), (777, 40, 10), (5000, 2048, 10)])
def test_scan_topk_matches_oracle(gpu, dtype, n, d, k):
    if dtype == "f32" and d * 4 > 4096:
        pytest.skip("row pitch > 4096 B not supported by the scan kernel")
    rng = np.random.default_rng(n * 31 + d)
    X = rng.standard_normal((n, d)).astype(np.float32) * rng.uniform(0.1, 4.0, (n, 1)).astype(np.float32)
    Q = rng.standard_normal((4, d)).astype(np.float32) * 3.0
    ix = gpu.DeviceIndex(d, dtype)
    assert ix.add(X) == 0 and len(ix) == n
    s, r = ix.query(Q, k, mode="scan")
    assert ix.last_query_path == "scan"
    _check(s, r, Q, X, k, dtype)
    ix.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_large_corpus_and_large_k(gpu, dtype):
    """k = 1000 is the UI's "All" (backend/app/main.py:757): radix-select path."""
    rng = np.random.default_rng(7)
    n, d = 300_000, 512
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((2, d)).astype(np.float32)
    ix = gpu.DeviceIndex(d, dtype)
    for lo in range(0, n, 70_000):          # several adds: exercises slab growth
        ix.add(X[lo:lo + 70_000])
    for k in (10, 128, 129, 1000):
        s, r = ix.query(Q, k, mode="scan")
        _check(s, r, Q, X, k, dtype)
    ix.close()


def test_golden_vectors(gpu):
    g = np.load(GOLD)
    X, Q, k = g["X"], g["Q"], int(g["k"])
    for dtype in ("f32", "bf16"):
        ix = gpu.DeviceIndex(X.shape[1], dtype)
        ix.add(X)
        s, r = ix.query(Q, k, mode="scan")
        np.testing.assert_allclose(s, g[f"scores_{dtype}"], atol=TOL[dtype], rtol=0)
        _check(s, r, Q, X, k, dtype)
        # query 1 is parallel to rows 3, 17 (exact duplicates) and 400 (scaled copy): all cos = 1
        assert set(r[1][:3].tolist()) == {3, 17, 400}
        ix.close()


def test_exact_ties_break_by_row(gpu):
    """Identical rows -> identical scores bit-for-bit -> ranking must be by ascending row."""
    rng = np.random.default_rng(3)
    d = 512
    base = rng.standard_normal(d).astype(np.float32)
    X = np.tile(base, (5000, 1))
    ix = gpu.DeviceIndex(d, "f32")
    ix.add(X)
    s, r = ix.query(base[None], 40, mode="scan")
    assert r[0].tolist() == list(range(40))
    assert np.all(s[0] == s[0][0])
    s, r = ix.query(base[None], 200, mode="scan")       # select path
    assert r[0].tolist() == list(range(200))
    ix.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("n", [40, 700, 50_000])
def test_ties_on_the_query_tail(gpu, dtype, n):
    """k <= 32 takes the cursor selection over packed keys (per-CTA lists, the last CTA's merge of every CTA list):
    blocks of identical rows give equal scores in different warps, CTAs and lists -- the order must still be
    (score desc, row asc), for one CTA (n = 40), a few (700: some warps of the last CTA get no list) and all 296."""
    rng = np.random.default_rng(11)
    d = 256
    a, b, c = (rng.standard_normal(d).astype(np.float32) for _ in range(3))
    q = a + 0.2 * b                                   # cos(q, a) ~ 0.98 > cos(q, a + b) ~ 0.83 > cos(q, b) ~ 0.2 > cos(q, c) ~ 0
    X = np.tile(c, (n, 1))
    idx_a = np.arange(5, n, 97)[:6]                   # copies of a, spread over the tiles
    idx_ab = np.arange(11, n, 89)[:9]
    idx_b = np.arange(2, n, 83)[:30]
    X[idx_b] = b
    X[idx_ab] = a + b
    X[idx_a] = a
    grp = np.full(n, 3)
    grp[idx_b], grp[idx_ab], grp[idx_a] = 2, 1, 0
    want = np.lexsort((np.arange(n), grp)).tolist()                # (score desc, row asc)
    ix = gpu.DeviceIndex(d, dtype)
    ix.add(X)
    for k in (1, 7, 10, 32):
        s, r = ix.query(q[None], k, mode="scan")
        kk = min(k, n)
        assert r[0][:kk].tolist() == want[:kk], (dtype, n, k, r[0].tolist(), want[:kk])
        assert np.all(np.diff(s[0][:kk]) <= 0)
    # a filter that leaves fewer rows than k: the tail must come back empty, the survivors in row order
    keep = sorted(np.nonzero(grp == 2)[0][:4].tolist())
    for row in keep:
        ix.set_filter_bits(row, [3])
    s, r = ix.query(q[None], 10, require_bits=[3], mode="scan")
    assert r[0][:len(keep)].tolist() == keep and (r[0][len(keep):] == -1).all() and np.isneginf(s[0][len(keep):]).all()
    ix.close()


def test_zero_rows_and_zero_query(gpu):
    rng = np.random.default_rng(4)
    X = rng.standard_normal((100, 128)).astype(np.float32)
    X[10] = 0
    ix = gpu.DeviceIndex(128, "f32")
    ix.add(X)
    s, r = ix.query(np.zeros((1, 128), np.float32), 5, mode="scan")
    assert r[0].tolist() == [0, 1, 2, 3, 4] and np.all(s[0] == 0)      # all scores 0 -> row order
    s, r = ix.query(X[20][None], 100, mode="scan")
    assert r[0][0] == 20 and abs(s[0][0] - 1) < 1e-6
    assert s[0][r[0].tolist().index(10)] == 0
    ix.close()


def test_k_larger_than_count_and_empty(gpu):
    ix = gpu.DeviceIndex(32, "bf16")
    q = np.ones((2, 32), np.float32)
    s, r = ix.query(q, 7)
    assert (r == -1).all() and np.isneginf(s).all()
    X = np.random.default_rng(0).standard_normal((3, 32)).astype(np.float32)
    ix.add(X)
    s, r = ix.query(q, 7, mode="scan")
    assert (r[:, 3:] == -1).all() and sorted(r[0][:3].tolist()) == [0, 1, 2]
    ix.close()


def test_remove_moves_last_row(gpu):
    rng = np.random.default_rng(5)
    n, d = 2000, 256
    X = rng.standard_normal((n, d)).astype(np.float32)
    ix = gpu.DeviceIndex(d, "f32")
    ix.add(X)
    assert ix.remove(n - 1) == -1
    assert ix.remove(5) == n - 2
    Xh = X[:n - 1].copy()
    Xh[5] = Xh[n - 2]
    Xh = Xh[:n - 2]
    np.testing.assert_array_equal(ix.get_rows(0, len(ix)), Xh)
    Q = rng.standard_normal((3, d)).astype(np.float32)
    s, r = ix.query(Q, 10, mode="scan")
    _check(s, r, Q, Xh, 10, "f32")
    ix.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("k", [10, 300])
def test_pre_filter_bits(gpu, dtype, k):
    """'pre' mode of the filter pass: only rows whose bits contain the required bits compete."""
    rng = np.random.default_rng(6)
    n, d = 3000, 128
    X = rng.standard_normal((n, d)).astype(np.float32)
    ix = gpu.DeviceIndex(d, dtype)
    ix.add(X)
    has0 = rng.random(n) < 0.3
    has70 = rng.random(n) < 0.5
    for row in range(n):
        bits = ([0] if has0[row] else []) + ([70] if has70[row] else [])
        if bits:
            ix.set_filter_bits(row, bits)
    assert ix.get_filter_bits(int(np.nonzero(has0 & has70)[0][0])) == [0, 70]
    Q = rng.standard_normal((2, d)).astype(np.float32)
    s, r = ix.query(Q, k, require_bits=[0, 70], mode="scan")
    _check(s, r, Q, X, k, dtype, valid=has0 & has70)
    s, r = ix.query(Q, k, require_bits=[200], mode="scan")     # nobody has bit 200
    assert (r == -1).all()
    ix.close()


def test_bf16_storage_is_round_to_nearest_even(gpu):
    rng = np.random.default_rng(8)
    X = rng.standard_normal((64, 72)).astype(np.float32) * 100
    ix = gpu.DeviceIndex(72, "bf16")
    ix.add(X)
    np.testing.assert_array_equal(ix.get_rows(0, 64), O.bf16_round(X))
    ix.close()


def test_merge_kernel_matches_oracle(gpu):
    import torch
    rng = np.random.default_rng(9)
    for G, B, k in ((2, 3, 10), (8, 5, 10), (8, 2, 1000), (4, 1, 128)):
        cs = rng.standard_normal((G, B, k)).astype(np.float32)
        cs = -np.sort(-cs, axis=2)
        cr = np.stack([rng.choice(10**6, size=(B, k), replace=False) + g * 10**6 for g in range(G)]).astype(np.int64)
        cr[G - 1, :, k - 2:] = -1                     # empty slots in the last shard
        cs[0, 0, 1] = cs[1, 0, 0]                     # a cross-shard exact tie
        es, er = O.merge_topk(cs, cr, k)
        s, r = gpu.DeviceIndex(8, "f32").merge_dev(torch.from_numpy(cs).cuda(), torch.from_numpy(cr).cuda())
        torch.cuda.synchronize()
        np.testing.assert_array_equal(r.cpu().numpy()[:, :er.shape[1]], er)
        np.testing.assert_array_equal(s.cpu().numpy()[:, :es.shape[1]], es)


def test_blend_kernel_and_multimodal_query(gpu):
    import torch
    g = np.load(GOLD)
    img, txt, w = g["blend_img"], g["blend_txt"], g["blend_w"]
    X = g["X"]
    ix = gpu.DeviceIndex(X.shape[1], "f32")
    ix.add(X)
    out = ix.blend_dev(torch.from_numpy(img).cuda(), torch.from_numpy(txt).cuda(), torch.from_numpy(w).cuda())
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), g["blend_out"], atol=2e-7, rtol=0)
    s, r = ix.query_multimodal(img, txt, w, 10, mode="scan")
    _check(s, r, g["blend_out"], X, 10, "f32")
    ix.close()


def test_device_query_api_and_launch_counter(gpu):
    import torch
    rng = np.random.default_rng(10)
    X = rng.standard_normal((10000, 512)).astype(np.float32)
    Q = rng.standard_normal((3, 512)).astype(np.float32)
    ix = gpu.DeviceIndex(512, "bf16", row_base=1_000_000)
    ix.add(torch.from_numpy(X).cuda())
    before = gpu.launch_count()
    s, r = ix.query_dev(torch.from_numpy(Q).cuda(), 10, mode="scan")
    torch.cuda.synchronize()
    assert gpu.launch_count() == before + 1               # ONE fused kernel for the whole batch
    _check(s.cpu().numpy(), r.cpu().numpy() - 1_000_000, Q, X, 10, "bf16")
    ix.close()


@pytest.mark.parametrize("k", [10, 200, 1000])
def test_two_shards_merged_equal_one_index(gpu, k):
    """The all-gather arm of the sharded query (k > 128 always takes it): per-shard top-k with shard
    row bases (contiguous shards) or the ShardedIndex's row map (striped shards) merged by K5 == one index."""
    import torch
    rng = np.random.default_rng(k)
    n, d, B = 30_001, 512, 6
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((B, d)).astype(np.float32)
    whole = gpu.DeviceIndex(d, "bf16")
    whole.add(X)
    s0, r0 = whole.query(Q, k, mode="scan")
    qd = torch.from_numpy(Q).cuda()
    for striped in (False, True):
        cs, cr, shards = [], [], []
        for g in range(2):
            # contiguous shards (row_base) or row-striped shards (ShardedIndex: reported row = g + local * 2)
            ix = gpu.DeviceIndex(d, "bf16", row_base=g, row_stride=2) if striped else \
                gpu.DeviceIndex(d, "bf16", row_base=g * 15_001)
            ix.add(X[g::2] if striped else X[g * 15_001:(g + 1) * 15_001])
            s, r = ix.query_dev(qd, k, mode="scan")
            cs.append(s); cr.append(r); shards.append(ix)
        s, r = shards[0].merge_dev(torch.stack(cs), torch.stack(cr))
        s, r = s.cpu().numpy(), r.cpu().numpy()
        assert np.array_equal(r, r0) and np.array_equal(s, s0)
        for ix in shards:
            ix.close()
    whole.close()
