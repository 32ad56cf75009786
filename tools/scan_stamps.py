"""Where does the time of ONE isolated request go?  Needs a tuning build with -DVS_SCAN_STAMPS:

    VS_BUILD_TUNING=1 VS_BUILD_DEFS=-DVS_SCAN_STAMPS python multimodal-image-similarity-search_b200/build.py
    VS_LIB_PATH=.../libvecsearch_b200_tuning.so python tools/scan_stamps.py [--rows 1250000] [--k 10]

The scan kernel marks %globaltimer at: first/last CTA entering, first tile landed (earliest/latest CTA), streaming done
(earliest/latest CTA), shard top-k selected (last CTA), result + flag written.  Combined with the host-side timeline of
vs_group_query_host (vs_group_last_timing) this splits a request into launch latency, ramp, streaming, merge tail, and the
way back over PCIe.  Medians over --queries requests on a one-GPU group.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmiss_b200 as M  # noqa: E402
from mmiss_b200 import _native  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_250_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--queries", type=int, default=300)
    a = ap.parse_args()
    lib = ctypes.CDLL(_native.LIB_PATH)
    fn = lib.vs_debug_scan_stamps
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_uint64)]
    dev = torch.device("cuda", 0)
    gx = M.GroupIndex(a.dim, "bf16", devices=[0], capacity=a.rows, b_max=64, k_max=32)
    g = torch.Generator(device=dev)
    for c0 in range(0, a.rows, 1 << 19):
        n = min(1 << 19, a.rows - c0)
        g.manual_seed(17 + c0)
        gx.shards[0].add(torch.nn.functional.normalize(torch.randn((n, a.dim), generator=g, device=dev), dim=1))
    torch.cuda.synchronize()
    torch.cuda.set_device(0)
    qs = np.random.default_rng(3).standard_normal((256, a.dim)).astype(np.float32)
    for i in range(32):
        gx.query(qs[i:i + 1], a.k, mode="scan")
    st = (ctypes.c_uint64 * 8)()
    rows = []
    for i in range(a.queries):
        assert fn(1, st) == 0
        gx.query(qs[i % 256:i % 256 + 1], a.k, mode="scan")
        tl = gx.last_timing_us()
        assert fn(0, st) == 0
        s = [int(x) for x in st]
        d = lambda x, y: (s[x] - s[y]) / 1e3   # noqa: E731
        rows.append([tl[1], tl[2] - tl[1], d(7, 0), d(1, 0), d(2, 0), d(3, 0), d(4, 0), d(5, 0), d(6, 5), d(7, 6)])
    m = np.median(np.array(rows), axis=0)
    names = ["host_enqueue_us", "host_enqueued_to_flag_seen_us", "device_first_cta_to_flag_written_us", "last_cta_entered_us",
             "first_tile_landed_earliest_us", "first_tile_landed_latest_us", "streaming_done_earliest_us",
             "streaming_done_latest_us", "merge_tail_us(last streaming done -> shard top-k)", "emit_us(result+fence+flag)"]
    out = {n: round(float(v), 2) for n, v in zip(names, m)}
    out["launch_latency_plus_pcie_return_us"] = round(float(m[1] - m[2]), 2)
    out["rows"], out["k"] = a.rows, a.k
    print(json.dumps(out), flush=True)
    gx.close()


if __name__ == "__main__":
    main()
