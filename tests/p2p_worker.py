"""Worker of tests/test_gpu_exchange.py::test_two_gpu_processes_ipc_exchange (one process per GPU,
launched by torch.distributed.run): fused peer exchange == NCCL all-gather + merge == oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmiss_b200 as M  # noqa: E402
from oracle import cosine_oracle as O  # noqa: E402


def ix_err(sh):
    return sh.local.exchange_error()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(0)
    n, d = 200_000, 512
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[n - 1] = X[7]
    Q = np.concatenate([rng.standard_normal((95, d)).astype(np.float32), X[7:8]])
    lo, hi = M.shard_bounds(n, world, rank)
    for dtype, tol in (("bf16", 2e-3), ("f32", 1e-5)):
        ix = M.DeviceIndex(d, dtype, device=local, row_base=lo)
        ix.add(X[lo:hi])
        p2p = M.ShardedSearcher.for_index(ix, mode="scan", exchange="p2p", b_max=128, k_max=32)
        nccl = M.ShardedSearcher.for_index(ix, mode="scan", exchange="nccl")
        assert p2p.exchange == "p2p" and nccl.exchange == "nccl"
        qd = torch.from_numpy(Q).to(dev)
        full = O.cosine_scores(Q, X, corpus_dtype=dtype)
        for k in (10, 32):
            for B in (1, 7, 64, 96):
                for rep in range(3):
                    s, r = p2p.search(qd[:B], k)
                    s2, r2 = nccl.search(qd[:B], k)
                    torch.cuda.synchronize()
                    assert ix.exchange_error() == 0
                    assert torch.equal(r, r2) and torch.equal(s, s2), (dtype, k, B, rep)
                sn, rn = s.cpu().numpy(), r.cpu().numpy()
                for b in range(B):
                    ok, why = O.topk_matches(sn[b], rn[b], full[b], k, tol)
                    assert ok, why
        # single-query launches back to back, no host synchronisation in between
        outs = [p2p.search(qd[b:b + 1], 10) for b in range(32)]
        torch.cuda.synchronize()
        for b, (s, r) in enumerate(outs):
            ok, why = O.topk_matches(s[0].cpu().numpy(), r[0].cpu().numpy(), full[b], 10, tol)
            assert ok, why
        # host-buffer flavour (one C call per request on every rank)
        hs, hr = ix.query_sharded(Q[:5], 10, mode="scan")
        s2, r2 = nccl.search(qd[:5], 10)
        assert np.array_equal(hr, r2.cpu().numpy()) and np.array_equal(hs, s2.cpu().numpy())
        # throughput mode: 32 pushes, one collect
        s, r = p2p.peer_exchange.search_stream([qd[b:b + 1] for b in range(32)], 10, mode="scan")
        s2, r2 = nccl.search(qd[:32], 10)
        torch.cuda.synchronize()
        assert ix.exchange_error() == 0 and torch.equal(r, r2) and torch.equal(s, s2)
        if dtype == "bf16":
            pt = M.ShardedSearcher(p2p.local_topk, p2p.merge, None, p2p.peer_exchange, "tensor")
            s, r = pt.search(qd, 10)
            torch.cuda.synchronize()
            fullr = O.cosine_scores(Q, X, corpus_dtype="bf16", round_queries=True)
            for b in range(Q.shape[0]):
                ok, why = O.topk_matches(s[b].cpu().numpy(), r[b].cpu().numpy(), fullr[b], 10, tol)
                assert ok, why
        if dtype == "bf16":
            # all-pairs pass: replicate the shards, split the triangle, gather the pairs
            Xd = X[:20_000].copy()
            Xd[15_000:15_100] = Xd[100:200] + (0.1 / np.sqrt(d)) * rng.standard_normal((100, d)).astype(np.float32)
            dlo, dhi = M.shard_bounds(len(Xd), world, rank)
            sh = M.DeviceIndex(d, "bf16", device=local, row_base=dlo)
            sh.add(Xd[dlo:dhi])
            full_ix = M.replicate_index(sh, len(Xd))
            assert np.array_equal(full_ix.get_rows(0, 64), O.bf16_round(Xd[:64]))
            pi, pj, ps = M.find_duplicates_sharded(lambda lo_, hi_: full_ix.dedup(0.95, lo_, hi_), len(Xd))
            wi, wj, ws = O.dedup_pairs(Xd, 0.95)
            assert set(zip(pi.tolist(), pj.tolist())) == set(zip(wi.tolist(), wj.tolist())) and len(wi) >= 100
            full_ix.close()
            sh.close()
        got = [torch.zeros_like(r) for _ in range(world)]
        dist.all_gather(got, r)
        assert all(torch.equal(got[0], g) for g in got)
        dist.barrier()
        ix.close()
    # ---- the reference's single server process on rank 0, shards served by the other ranks ----------
    sh = M.ShardedIndex(d, "bf16", device=local, exchange="p2p", mode="scan", b_max=64, k_max=128)
    if rank != 0:
        sh.serve()
    else:
        n2 = 30_001
        X = X[:n2].copy()
        X[n2 - 1] = X[7]                               # exact tie across shards: ordered by global row, like one index
        Q = np.concatenate([Q[:5], X[7:8]])
        ids = [f"img_{i:05x}" for i in range(n2)]
        metas = [{"filename": f"{i}.jpg", "filter_results_json": '{"is it red?": "%s"}' % ("yes" if i % 4 == 0 else "no")}
                 for i in range(n2)]
        many = M.Collection("many", {"hnsw:space": "cosine"}, index=sh)
        one = M.Collection("one", {"hnsw:space": "cosine"}, dtype="bf16", device=local)
        for col in (one, many):
            col.add(ids=ids[:20_000], embeddings=X[:20_000], metadatas=metas[:20_000])
            col.add(ids=ids[20_000:], embeddings=X[20_000:n2], metadatas=metas[20_000:])
        assert many.count() == n2 and len(sh) == n2

        def same(**kw):
            a = one.query(query_embeddings=Q[:6].tolist(), **kw)
            b = many.query(query_embeddings=Q[:6].tolist(), **kw)
            assert a["ids"] == b["ids"], kw
            for da, db in zip(a["distances"], b["distances"]):
                np.testing.assert_allclose(da, db, atol=1e-6)
        same(n_results=10, include=["metadatas", "distances"])
        same(n_results=10, include=["distances"], where_filters=["is it red?"], filter_mode="pre")
        same(n_results=1000, include=["distances"])                          # k > k_max: all-gather + merge path
        qd6 = torch.from_numpy(Q[:6]).to(dev)                                # device-resident queries: broadcast over NCCL as they are
        assert one.query(query_embeddings=qd6, n_results=10, include=["distances"])["ids"] == \
            many.query(query_embeddings=qd6, n_results=10, include=["distances"])["ids"]
        for col in (one, many):
            col.delete(ids=[ids[3], ids[n2 - 1], ids[12_345]])
            col.update(ids=[ids[4]], metadatas=[{"filter_results_json": '{"is it red?": "no"}'}])
        same(n_results=10, include=["metadatas", "distances"])
        same(n_results=10, include=["distances"], where_filters=["is it red?"], filter_mode="pre")
        # configs 4 / 5 and the multimodal blend on the sharded collection (every rank's device does its part)
        P = np.random.default_rng(77).standard_normal((3, d)).astype(np.float32)
        np.testing.assert_array_equal(many.filter_sweep(P, 0.08), one.filter_sweep(P, 0.08))
        for col in (one, many):
            assert col.apply_filter_sweep("like P2", P[2], 0.08) == int(one.filter_sweep(P[2:3], 0.08).sum())
        same(n_results=10, include=["metadatas", "distances"], where_filters=["like P2"], filter_mode="pre")
        a = one.query_multimodal(Q[:3], Q[3:6], [0.2, 0.5, 0.8], n_results=10, include=["distances"])
        b = many.query_multimodal(Q[:3], Q[3:6], [0.2, 0.5, 0.8], n_results=10, include=["distances"])
        assert a["ids"] == b["ids"]
        for col in (one, many):                                              # plant near-duplicates of rows 10 and 11
            col.add(ids=["dup_a", "dup_b"], embeddings=X[10:12] + 0.002 * P[:2])
        assert set((a_, b_) for a_, b_, _ in many.find_duplicates(0.95)) == \
            set((a_, b_) for a_, b_, _ in one.find_duplicates(0.95)) == {(ids[10], "dup_a"), (ids[11], "dup_b")}
        assert ix_err(sh) == 0
        many.close()
        one.close()
    dist.barrier()
    if rank == 0:
        print("p2p worker ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
