"""Device-timed benchmarks of the tcgen05 paths (BASELINE configs 3, 4, 5) on ONE GPU.

    python tools/bench_tensor.py [--rows 1250000] [--batch 1024] [--filters 256] [--dedup-rows 200000]

Prints one JSON line per kernel: time, TFLOP/s and the fraction of the measured bf16 peaks
(MEASURED_PEAKS.json: burst for a kernel timed alone).  Inputs are generated on the device and are
far larger than L2; timing is CUDA events on the launch stream after 3 warm-ups.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmiss_b200 as M  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def timed(fn, iters, min_seconds=0.3):
    """ms per call: >= `iters` calls and >= `min_seconds` of device time, after a warm-up of the same length (a few
    milliseconds are not enough for the clocks / the power governor to settle)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    n = max(iters, int(min_seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)) + 1) if iters > 1 else 1
    for _ in range(n if iters > 1 else 0):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def build_index(rows, dim, dev, planted=0):
    ix = M.DeviceIndex(dim, "bf16", device=dev.index or 0, capacity=rows)
    gen = torch.Generator(device=dev)
    chunk = min(1 << 18, max(1, rows // 2)) if planted else 1 << 18
    first = None
    for c0 in range(0, rows, chunk):
        n = min(chunk, rows - c0)
        gen.manual_seed(77 + c0)
        x = torch.nn.functional.normalize(torch.randn((n, dim), generator=gen, device=dev), dim=1)
        if planted and c0 == 0:
            first = x[:planted].clone()
        elif planted and first is not None and c0 + n >= rows:      # last chunk: noisy copies (cos ~ 0.995)
            m = min(first.shape[0], n)
            x[:m] = torch.nn.functional.normalize(
                first[:m] + 0.1 / dim ** 0.5 * torch.randn((m, dim), generator=gen, device=dev), dim=1)
        ix.add(x)
    torch.cuda.synchronize()
    return ix


def cublas_now(seconds=1.5):
    """torch.matmul bf16 8192^3 back to back on THIS box right now (the way MEASURED_PEAKS.json's sustained figure was
    taken): boxes differ by ~10 % under the power cap, so fractions against this number compare across runs."""
    a = torch.randn((8192, 8192), device="cuda", dtype=torch.bfloat16)
    b = torch.randn((8192, 8192), device="cuda", dtype=torch.bfloat16)
    for _ in range(5):
        a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    e0.record()
    import time
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            a @ b
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return 2.0 * 8192 ** 3 * n / (e0.elapsed_time(e1) / 1e3) / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_250_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--filters", type=int, default=256)
    ap.add_argument("--dedup-rows", type=int, default=200_000)
    ap.add_argument("--dedup-dim", type=int, default=768)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--skip", default="")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    pk, kind = peaks()
    cub = cublas_now()
    knobs = {k_: v for k_, v in os.environ.items() if k_.startswith("VS_TC_") or k_ == "VS_LIB_PATH"}

    def report(name, ms, flops, extra):
        tf = flops / (ms / 1e3) / 1e12
        print(json.dumps({"kernel": name, "tag": a.tag, "ms": ms, "tflops": tf, "frac_burst": tf / pk["bf16_tflops"],
                          "frac_sustained": tf / pk["bf16_tflops_sustained"], "peak_kind": kind,
                          "cublas_bf16_now_tflops": cub, "frac_of_cublas_now": tf / cub, "knobs": knobs, **extra}), flush=True)

    if "topk" not in a.skip or "filter" not in a.skip:
        ix = build_index(a.rows, a.dim, dev)
    if "topk" not in a.skip:
        g = torch.Generator(device=dev).manual_seed(5)
        img = torch.randn((a.batch, a.dim), generator=g, device=dev)
        txt = torch.randn((a.batch, a.dim), generator=g, device=dev)
        w = torch.rand((a.batch,), generator=g, device=dev, dtype=torch.float64)
        q = torch.empty((a.batch, a.dim), device=dev)
        out_s = torch.empty((a.batch, a.k), device=dev)
        out_r = torch.empty((a.batch, a.k), dtype=torch.int64, device=dev)

        def f():
            ix.blend_dev(img, txt, w, out=q)                       # multimodal blend (main.py:850-860)
            ix.query_dev(q, a.k, out_scores=out_s, out_rows=out_r, mode="tensor")
        ms = timed(f, a.iters)
        report("multimodal_topk_tensor", ms, 2.0 * a.batch * a.rows * a.dim,
               {"rows": a.rows, "dim": a.dim, "batch": a.batch, "k": a.k, "qps": a.batch / (ms / 1e3),
                "corpus_gb_per_s": a.rows * a.dim * 2 / (ms / 1e3) / 1e9})
    if "filter" not in a.skip:
        g = torch.Generator(device=dev).manual_seed(6)
        prompts = torch.randn((a.filters, a.dim), generator=g, device=dev)
        bits = torch.zeros((a.filters, ix.filter_words()), dtype=torch.int32, device=dev)
        ms = timed(lambda: ix.filter_sweep_dev(prompts, 0.103, out_bits=bits), a.iters)
        nbytes = a.rows * a.dim * 2 + a.filters * a.rows / 8
        report("filter_sweep", ms, 2.0 * a.filters * a.rows * a.dim,
               {"rows": a.rows, "dim": a.dim, "filters": a.filters, "gb_per_s": nbytes / (ms / 1e3) / 1e9,
                "frac_hbm": nbytes / (ms / 1e3) / 1e9 / pk["hbm_gbs"],
                "pass_rate": float((bits != 0).float().mean())})
    if "topk" not in a.skip or "filter" not in a.skip:
        ix.close()
    if "dedup" not in a.skip:
        n = a.dedup_rows
        ix = build_index(n, a.dedup_dim, dev, planted=2000)
        cap = 1 << 20
        oi = torch.empty(cap, dtype=torch.int64, device=dev)
        oj = torch.empty(cap, dtype=torch.int64, device=dev)
        os_ = torch.empty(cap, dtype=torch.float32, device=dev)
        cnt = torch.zeros(2, dtype=torch.int64, device=dev)
        ms = timed(lambda: ix.dedup_dev(0.95, 0, n, oi, oj, os_, cnt), max(1, a.iters // 2))
        report("dedup_allpairs", ms, float(a.dedup_dim) * n * (n - 1),
               {"rows": n, "dim": a.dedup_dim, "pairs_found": int(cnt[0].item()), "flops_counted": "useful triangle"})
        ix.close()


if __name__ == "__main__":
    main()
