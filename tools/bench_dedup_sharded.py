"""BASELINE config 5: all-pairs duplicate detection (cos >= 0.95) over N x 768 bf16 rows, the triangle
split over the ranks of one box (one process per GPU, launched with torch.distributed.run).

    python -m torch.distributed.run --nproc-per-node G tools/bench_dedup_sharded.py [--rows 2000000] [--planted 20000]

Every rank holds the full corpus (replicated at load: identical seeded generation), runs K4 on its
triangle slice (triangle_bounds: equal work, not equal rows), device-timed; rank 0 prints one JSON line:
useful flops = D*N*(N-1), time = max over ranks, and the planted pairs must all be found.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmiss_b200 as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--planted", type=int, default=20_000)
    ap.add_argument("--tau", type=float, default=0.95)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, d, P = a.rows, a.dim, a.planted
    ix = M.DeviceIndex(d, "bf16", device=local, capacity=n)
    gen = torch.Generator(device=dev)
    chunk = 1 << 18
    first = None
    for c0 in range(0, n, chunk):                                   # identical on every rank
        m = min(chunk, n - c0)
        gen.manual_seed(4242 + c0)
        x = torch.nn.functional.normalize(torch.randn((m, d), generator=gen, device=dev), dim=1)
        if c0 == 0:
            first = x[:P].clone()
        if c0 + m >= n and P > 0:                                   # last chunk: noisy copies of the first P rows
            p = min(P, m)
            x[m - p:] = torch.nn.functional.normalize(
                first[:p] + 0.1 / d ** 0.5 * torch.randn((p, d), generator=gen, device=dev), dim=1)
        ix.add(x)
    torch.cuda.synchronize()
    lo, hi = M.triangle_bounds(n, world, rank)
    cap = max(1 << 20, 4 * P)
    oi = torch.empty(cap, dtype=torch.int64, device=dev)
    oj = torch.empty(cap, dtype=torch.int64, device=dev)
    os_ = torch.empty(cap, dtype=torch.float32, device=dev)
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    ix.dedup_dev(a.tau, lo, min(hi, lo + 4096), oi, oj, os_, cnt)    # warm-up on a thin slice
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ix.dedup_dev(a.tau, lo, hi, oi, oj, os_, cnt)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), float(cnt[0].item())], dtype=torch.float64, device=dev)
    tmax, tsum = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peaks = json.load(open(pk)) if os.path.exists(pk) else {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
        ms = float(tmax[0].item())
        tf = float(d) * n * (n - 1) / (ms / 1e3) / 1e12
        print(json.dumps({"kernel": "dedup_allpairs_sharded", "rows": n, "dim": d, "n_gpus": world, "ms": ms,
                          "useful_tflops_total": tf, "useful_tflops_per_gpu": tf / world,
                          "frac_burst_per_gpu": tf / world / peaks["bf16_tflops"],
                          "frac_sustained_per_gpu": tf / world / peaks["bf16_tflops_sustained"],
                          "pairs_found": int(tsum[1].item()), "planted": P,
                          "rows_of_rank0": [lo, hi]}), flush=True)
        assert int(tsum[1].item()) >= P, "planted duplicates missing"
    ix.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
