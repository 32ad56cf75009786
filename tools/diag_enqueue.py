"""Does anything a process did EARLIER make the single-process group's launches slower?  One process: measure the host-side
enqueue time of a group request, run an operation, measure again.  (bench.py's full default run showed 7.6 us where a fresh
process shows 4.2 us.)

    python tools/diag_enqueue.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmiss_b200 as M  # noqa: E402

dev = torch.device("cuda", 0)
DIM = 512


def fill(ix, n, dim=DIM, seed=1):
    g = torch.Generator(device=dev)
    for c0 in range(0, n, 1 << 19):
        m = min(1 << 19, n - c0)
        g.manual_seed(seed + c0)
        ix.add(torch.nn.functional.normalize(torch.randn((m, dim), generator=g, device=dev), dim=1))
    torch.cuda.synchronize()


gx = M.GroupIndex(DIM, "bf16", devices=[0], capacity=1_250_000, b_max=1024, k_max=32)
fill(gx.shards[0], 1_250_000)
qs = np.random.default_rng(3).standard_normal((256, DIM)).astype(np.float32)


def measure(tag, n=1500):
    for i in range(32):
        gx.query(qs[i:i + 1], 10, mode="scan")
    tl = np.empty((n, 4)); lat = np.empty(n)
    for i in range(n):
        t0 = time.perf_counter()
        gx.query(qs[i % 256:i % 256 + 1], 10, mode="scan")
        lat[i] = time.perf_counter() - t0
        tl[i] = gx.last_timing_us()
    m = np.median(tl, axis=0)
    print(json.dumps({"after": tag, "enqueued_us": round(float(m[1]), 2), "flag_seen_us": round(float(m[2]), 1),
                      "p50_us": round(float(np.percentile(lat, 50) * 1e6), 1),
                      "python_us": round(float(np.median(lat) * 1e6 - m[3]), 2)}), flush=True)


measure("fresh")
evs = [torch.cuda.Event(enable_timing=True) for _ in range(4000)]
for e in evs:
    e.record()
torch.cuda.synchronize()
measure("4000 recorded timing events alive")
del evs
measure("events freed")
big = M.DeviceIndex(DIM, "bf16", device=0, capacity=10_000_000)
fill(big, 10_000_000, seed=77)
measure("second index, 10M rows, open")
q1 = torch.randn((1, DIM), device=dev)
s1 = torch.empty((1, 10), device=dev); r1 = torch.empty((1, 10), dtype=torch.int64, device=dev)
for _ in range(300):
    big.query_dev(q1, 10, out_scores=s1, out_rows=r1, mode="scan", pipelined=True)
torch.cuda.synchronize()
measure("300 pipelined scans on it")
qb = torch.randn((1024, DIM), device=dev)
bs = torch.empty((1024, 10), device=dev); br = torch.empty((1024, 10), dtype=torch.int64, device=dev)
for _ in range(20):
    big.query_dev(qb, 10, out_scores=bs, out_rows=br, mode="tensor")
torch.cuda.synchronize()
measure("K2 batches (B=1024) on it")
pr = torch.randn((256, DIM), device=dev)
fb = torch.zeros((256, big.filter_words()), dtype=torch.int32, device=dev)
for _ in range(20):
    big.filter_sweep_dev(pr, 0.103, out_bits=fb)
torch.cuda.synchronize()
measure("K3 filter sweeps on it")
hq = np.random.default_rng(5).standard_normal((32, DIM)).astype(np.float32)
for _ in range(5):
    big.query(hq, 10, mode="scan")
measure("host-buffer queries (vs_query_topk_host) on it")
fx = M.DeviceIndex(DIM, "f32", device=0, capacity=1_000_000)
fill(fx, 1_000_000, seed=5)
for _ in range(100):
    fx.query_dev(q1, 10, out_scores=s1, out_rows=r1, mode="scan", pipelined=True)
torch.cuda.synchronize()
fx.close()
measure("f32 index created, scanned, closed")
dx = M.DeviceIndex(768, "bf16", device=0, capacity=400_000)
fill(dx, 400_000, dim=768, seed=9)
cap = 1 << 17
oi = torch.empty(cap, dtype=torch.int64, device=dev); oj = torch.empty(cap, dtype=torch.int64, device=dev)
osc = torch.empty(cap, dtype=torch.float32, device=dev); cnt = torch.zeros(2, dtype=torch.int64, device=dev)
dx.dedup_dev(0.95, 0, 400_000, oi, oj, osc, cnt)
torch.cuda.synchronize()
dx.close()
measure("K4 dedup index created, run, closed")
big.close()
measure("10M index closed")
a = np.random.default_rng(1).standard_normal((2000, 2000)).astype(np.float32)
for _ in range(5):
    a @ a
measure("numpy matmuls (BLAS thread pool started)")
gx.close()
