"""Device-timed micro-benchmark of the scan kernel (K1) on ONE GPU for a list of shard sizes.

    [VS_SCAN_SMEM_KB=110] [VS_SCAN_CTAS_PER_SM=2] [VS_SCAN_PDL=0] python tools/bench_scan.py \
        [--rows 10000000,5000000,2500000,1250000] [--dtype bf16] [--queries 32] [--iters 5]

Prints one JSON line per shard size: ms per scan averaged over back-to-back single-query launches
(the bench.py pattern), GB/s against MEASURED_PEAKS.json, and the latency of ONE isolated query.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmiss_b200 as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="10000000,5000000,2500000,1250000")
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--queries", type=int, default=32)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    sizes = [int(x) for x in a.rows.split(",")]
    ix = M.DeviceIndex(a.dim, a.dtype, device=0, capacity=max(sizes))
    gen = torch.Generator(device=dev)
    env = {k: v for k, v in os.environ.items() if k.startswith("VS_SCAN")}
    q = torch.nn.functional.normalize(torch.randn((a.queries, a.dim), device=dev), dim=1)
    s = torch.empty((a.queries, a.k), dtype=torch.float32, device=dev)
    r = torch.empty((a.queries, a.k), dtype=torch.int64, device=dev)
    for n in sorted(sizes):                                     # grow the same index
        while len(ix) < n:
            m = min(1 << 19, n - len(ix))
            gen.manual_seed(1234 + len(ix))
            ix.add(torch.nn.functional.normalize(torch.randn((m, a.dim), generator=gen, device=dev), dim=1))
        torch.cuda.synchronize()

        def burst():
            for i in range(a.queries):
                ix.query_dev(q[i:i + 1], a.k, out_scores=s[i:i + 1], out_rows=r[i:i + 1], mode="scan", pipelined=True)

        for _ in range(3):
            burst()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            burst()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (a.iters * a.queries)
        lat = []
        for i in range(8):                                       # isolated launches
            e0.record()
            ix.query_dev(q[i:i + 1], a.k, out_scores=s[i:i + 1], out_rows=r[i:i + 1], mode="scan")
            e1.record()
            torch.cuda.synchronize()
            lat.append(e0.elapsed_time(e1))
        bytes_ = n * (a.dim * (2 if a.dtype == "bf16" else 4) + 4)
        print(json.dumps({"rows": n, "dim": a.dim, "dtype": a.dtype, "env": env, "ms_per_scan": round(ms, 5),
                          "gbs": round(bytes_ / ms / 1e6, 1), "frac": round(bytes_ / ms / 1e6 / peak, 4),
                          "isolated_ms": round(sorted(lat)[len(lat) // 2], 5),
                          "isolated_frac": round(bytes_ / sorted(lat)[len(lat) // 2] / 1e6 / peak, 4)}), flush=True)
    ix.close()


if __name__ == "__main__":
    main()
