"""-m gpu: the fused peer exchange (vs_query_topk_sharded_dev / vs_exchange_merge_dev) against the
CPU oracle.

* one GPU: G shards of one corpus live on cuda:0 as G handles of THIS process, their exchange
  buffers wired together by pointer; every "rank" issues its kernels on its own stream and the
  kernels of the G ranks meet on the device exactly as they would across NVLink (same protocol,
  same flags, same bounded spin).
* two GPUs (skipped on a 1-GPU box): two processes under torch.distributed.run, buffers mapped with
  CUDA IPC, fused exchange vs the NCCL all-gather path vs the oracle.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import cosine_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = {"f32": 1e-5, "bf16": 2e-3}


def _shards(gpu, X, G, dtype, b_max, k_max, uneven=False):
    import torch
    n = X.shape[0]
    if uneven:                                   # ragged shards, the last one empty
        cuts = [0] + [min(n, (n * (g + 1)) // (G - 1) + (13 * g) % 7) for g in range(G - 2)] + [n, n]
    else:
        cuts = [gpu.shard_bounds(n, G, g)[0] for g in range(G)] + [n]
    idx = []
    for g in range(G):
        ix = gpu.DeviceIndex(X.shape[1], dtype, device=0, row_base=cuts[g])
        if cuts[g + 1] > cuts[g]:
            ix.add(X[cuts[g]:cuts[g + 1]])
        ix.exchange_create(G, g, b_max, k_max)
        idx.append(ix)
    ptrs = [ix.exchange_local_ptr() for ix in idx]
    for ix in idx:
        ix.exchange_attach(peer_ptrs=ptrs)
    streams = [torch.cuda.Stream() for _ in range(G)]
    return idx, streams


def _run_all(idx, streams, q, k, mode):
    import torch
    outs = []
    for ix, st in zip(idx, streams):             # asynchronous launches: the ranks meet on the device
        with torch.cuda.stream(st):
            outs.append(ix.query_sharded_dev(q, k, mode=mode))
    torch.cuda.synchronize()
    for ix in idx:
        assert ix.exchange_error() == 0
    return outs


def _check_all(outs, Q, X, k, dtype, round_queries=False):
    full = O.cosine_scores(Q, X, corpus_dtype=dtype, round_queries=round_queries) if round_queries else \
        O.cosine_scores(Q, X, corpus_dtype=dtype)
    s0, r0 = outs[0][0].cpu().numpy(), outs[0][1].cpu().numpy()
    kk = min(k, X.shape[0])
    for b in range(Q.shape[0]):
        ok, why = O.topk_matches(s0[b][:kk], r0[b][:kk], full[b], kk, TOL[dtype])
        assert ok, f"query {b}: {why}"
        assert (r0[b][kk:] == -1).all()
    for s, r in outs[1:]:                        # every rank ends with the same answer, bit for bit
        assert np.array_equal(r.cpu().numpy(), r0) and np.array_equal(s.cpu().numpy(), s0)


@pytest.mark.parametrize("G", [2, 4, 8])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_fused_scan_exchange_one_gpu(gpu, G, dtype):
    import torch
    rng = np.random.default_rng(G)
    n, d = 30_000, 512
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[n - 1] = X[5]                              # exact tie across the first and the last shard
    Q = np.concatenate([rng.standard_normal((5, d)).astype(np.float32), X[5:6]])
    idx, streams = _shards(gpu, X, G, dtype, b_max=256, k_max=128)
    qd = torch.from_numpy(Q).cuda()
    for k in (10, 1, 32, 100):
        for rep in range(3):                     # consecutive epochs reuse both buffer halves
            outs = _run_all(idx, streams, qd, k, "scan")
            _check_all(outs, Q, X, k, dtype)
        assert outs[0][1][5][:2].tolist() == [5, n - 1][:k]
    # single-query launches back to back (the bench's pattern: PDL + one exchange per launch)
    for rep in range(4):
        for b in range(Q.shape[0]):
            outs = _run_all(idx, streams, qd[b:b + 1], 10, "scan")
            _check_all(outs, Q[b:b + 1], X, 10, dtype)
    for ix in idx:
        ix.close()


def test_exchange_kernel_batches_and_tensor_path_one_gpu(gpu):
    """B > one scan launch (64) and the tcgen05 path go through the stand-alone exchange kernel."""
    import torch
    rng = np.random.default_rng(11)
    n, d, G = 40_000, 512, 4
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((300, d)).astype(np.float32)
    idx, streams = _shards(gpu, X, G, "bf16", b_max=128, k_max=32)     # 300 queries -> 3 chunks of <=128 slots
    qd = torch.from_numpy(Q).cuda()
    outs = _run_all(idx, streams, qd, 10, "scan")
    _check_all(outs, Q, X, 10, "bf16")
    outs = _run_all(idx, streams, qd, 10, "tensor")
    _check_all(outs, Q, X, 10, "bf16", round_queries=True)
    outs = _run_all(idx, streams, qd[:40], 32, "auto")
    _check_all(outs, Q[:40], X, 32, "bf16", round_queries=True)
    for ix in idx:
        ix.close()


def test_deferred_push_collect_one_gpu(gpu):
    """Throughput mode: every query kernel pushes into its own slots, ONE collect kernel per batch."""
    import torch
    rng = np.random.default_rng(21)
    n, d, G = 50_000, 512, 4
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((40, d)).astype(np.float32)
    idx, streams = _shards(gpu, X, G, "bf16", b_max=64, k_max=32)
    qd = torch.from_numpy(Q).cuda()
    for rep in range(3):
        outs = []
        for ix, st in zip(idx, streams):
            with torch.cuda.stream(st):
                ix.exchange_begin()
                for b in range(24):                               # 24 single-query scans (fused push)
                    ix.query_push_dev(qd[b:b + 1], 10, b, mode="scan")
                ix.query_push_dev(qd[24:40], 10, 24, mode="tensor")   # 16 more through K2 + push kernel
                outs.append(ix.exchange_collect_dev(40, 10))
        torch.cuda.synchronize()
        assert all(ix.exchange_error() == 0 for ix in idx)
        _check_all([(s[:24], r[:24]) for s, r in outs], Q[:24], X, 10, "bf16")
        _check_all([(s[24:], r[24:]) for s, r in outs], Q[24:], X, 10, "bf16", round_queries=True)
    with pytest.raises(gpu.VecSearchError):
        idx[0].query_push_dev(qd[:8], 10, 60)                     # slots [60,68) exceed B_max=64
    for ix in idx:
        ix.close()


def test_ragged_and_empty_shards_one_gpu(gpu):
    import torch
    rng = np.random.default_rng(3)
    n, d, G = 1000, 64, 4
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((3, d)).astype(np.float32)
    idx, streams = _shards(gpu, X, G, "f32", b_max=64, k_max=32, uneven=True)
    assert len(idx[-1]) == 0 and sum(len(ix) for ix in idx) == n
    qd = torch.from_numpy(Q).cuda()
    for k in (10, 32):
        outs = _run_all(idx, streams, qd, k, "scan")
        _check_all(outs, Q, X, k, "f32")
    for ix in idx:
        ix.close()


def test_exchange_argument_errors(gpu):
    import torch
    ix = gpu.DeviceIndex(64, "f32")
    ix.add(np.ones((4, 64), np.float32))
    q = torch.ones((1, 64), device="cuda")
    with pytest.raises(gpu.VecSearchError):
        ix.query_sharded_dev(q, 2)                       # exchange not created
    with pytest.raises(gpu.VecSearchError):
        ix.exchange_create(9, 0)                         # > 8 ranks
    ix.exchange_create(2, 0, 16, 8)
    with pytest.raises(gpu.VecSearchError):
        ix.query_sharded_dev(q, 2)                       # not attached
    with pytest.raises(gpu.VecSearchError):
        ix.exchange_attach()                             # no handle, no pointer for peer 1
    ix.close()


def test_two_gpu_processes_ipc_exchange(gpu):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "p2p_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "p2p worker ok" in out.stdout
