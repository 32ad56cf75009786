"""One request at a time through the single-process group (vs_group_query_host): latency percentiles and the host-side
timeline of a request, for a given shard size per GPU.

    python tools/bench_group.py [--rows-per-gpu 1250000] [--devices 0,1,...] [--queries 500]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmiss_b200 as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=1_250_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--devices", default="")
    ap.add_argument("--queries", type=int, default=500)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--b-max", type=int, default=64)
    a = ap.parse_args()
    devs = [int(x) for x in a.devices.split(",")] if a.devices else list(range(torch.cuda.device_count()))
    G = len(devs)
    gx = M.GroupIndex(a.dim, a.dtype, devices=devs, capacity=a.rows_per_gpu * G, b_max=a.b_max, k_max=32)
    for s, sh in enumerate(gx.shards):
        dv = torch.device("cuda", sh.device)
        g = torch.Generator(device=dv)
        with torch.cuda.device(dv):
            for c0 in range(0, a.rows_per_gpu, 1 << 19):
                n = min(1 << 19, a.rows_per_gpu - c0)
                g.manual_seed(17 + c0 + s * 7919)
                sh.add(torch.nn.functional.normalize(torch.randn((n, a.dim), generator=g, device=dv), dim=1))
            torch.cuda.synchronize()
    rng = np.random.default_rng(3)
    qs = rng.standard_normal((256, a.dim)).astype(np.float32)
    for i in range(32):
        gx.query(qs[i:i + 1], a.k, mode="scan")
    lat = np.empty(a.queries)
    tl = np.empty((a.queries, 4))
    w0 = time.perf_counter()
    for i in range(a.queries):
        t0 = time.perf_counter()
        gx.query(qs[i % 256:i % 256 + 1], a.k, mode="scan")
        lat[i] = time.perf_counter() - t0
        tl[i] = gx.last_timing_us()
    wall = time.perf_counter() - w0
    m = np.median(tl, axis=0)
    print(json.dumps({"devices": devs, "rows_per_gpu": a.rows_per_gpu, "dim": a.dim, "dtype": a.dtype, "b_max": a.b_max, "qps": a.queries / wall,
                      "latency_us": {"p50": float(np.percentile(lat, 50) * 1e6), "p99": float(np.percentile(lat, 99) * 1e6),
                                     "min": float(lat.min() * 1e6)},
                      "timeline_us_median": {"published": float(m[0]), "enqueued": float(m[1]), "flag_seen": float(m[2]),
                                             "returned": float(m[3])}}), flush=True)
    gx.close()


if __name__ == "__main__":
    main()
