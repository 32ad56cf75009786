#!/usr/bin/env python
"""bench.py -- headline benchmark: exact top-10 cosine search over 10M x 512 bf16 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA, C ABI)
    python bench.py --impl reference [...]                         # reference arm (CPU exact path)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # N ranks, one GPU each

A *step* is one batch of Q independent single queries over the whole corpus: every query is its
own fused scan kernel launch that streams the rank's shard from HBM once (K1).  With N > 1 the LAST
CTA of that same kernel stores the shard's k candidates into every peer's exchange buffer over
NVLink (P2P stores + flags) and merges the N lists, so every rank ends with the global top-k and
the query is still ONE kernel, no NCCL call (--exchange nccl: all-gather + merge kernel K5).  The
corpus (10M rows total, row-sharded ceil(N/G) per rank => "strong" scaling) is resident in HBM
before the timed region; it is >> L2 (126 MB), so no flush is needed between iterations.

`value`   = queries/s over all ranks, device-timed (CUDA events, max over ranks).
`e2e`     = same metric through the host-buffer API: per step ONE H2D of the step's queries from pinned
            memory, the Q single-query scans (+ exchange when N>1), D2H of the [Q, k] result, sync;
            `e2e.per_query_sync` = the same with copy + sync per query (one request at a time).
`roofline`= the scan kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
`cpu_baseline` = the oracle port of the reference's exact path (numpy matmul + top-k) on this
            box's host cores, on a bounded row sample, scaled to the full corpus.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOTAL_ROWS = 10_000_000
DIM = 512
TOPK = 10
METRIC = "exact top-10 cosine search QPS, 10Mx512 bf16 corpus, single-query scans"
HBM_FALLBACK_GBS = 6650.0


def ncu_traffic_bytes(rows_local: int, dim: int, dtype: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per scan launch from the committed `ncu --set full`
    capture (profiles/r01c_ncu_scan_topk.txt, first launch; taken on 10M x 512 bf16); None for any other shape."""
    if (rows_local, dim, dtype) != (10_000_000, 512, "bf16"):
        return None
    try:
        total = 0.0
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        seen = set()
        for ln in open(os.path.join(ROOT, "profiles", "r01c_ncu_scan_topk.txt")):
            key = ln.split("=")[0].strip()
            if key in ("dram__bytes_read.sum", "dram__bytes_write.sum") and key not in seen:
                seen.add(key)
                val, unit = ln.split("=")[1].split()
                total += float(val) * scale[unit]
        return total or None
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": HBM_FALLBACK_GBS, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's exact path (numpy `X @ q` + argpartition),
# restated in oracle/cosine_oracle.py; timed on a bounded sample.
# ------------------------------------------------------------------------------------------------
def cpu_reference_qps(total_rows: int, dim: int, k: int, budget_s: float, sample_rows: int = 1_000_000,
                      min_queries: int = 3):
    from oracle import cosine_oracle as O
    # all the host threads BLAS can use, also under torchrun (which exports OMP_NUM_THREADS=1)
    ncores = int(os.cpu_count() or 1)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        _limit = threadpool_limits(limits=ncores)
        threads = max([d.get("num_threads", 1) for d in threadpool_info()] or [1])
    except Exception:                                 # noqa: BLE001
        import torch
        _limit, threads = None, torch.get_num_threads()
    rng = np.random.default_rng(0)
    sample_rows = min(sample_rows, total_rows)
    X = O.bf16_round(O.normalize_rows(rng.standard_normal((sample_rows, dim), dtype=np.float32)))
    inv = O.inv_norms(X)
    Q = O.normalize_rows(rng.standard_normal((64, dim), dtype=np.float32))

    def one(q):
        s = (X @ q) * inv                      # the scan
        idx = np.argpartition(-s, k)[:k]       # top-k
        return idx[np.argsort(-s[idx], kind="stable")]

    one(Q[0])                                  # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        one(Q[n % 64])
        n += 1
        dt = time.perf_counter() - t0
        if (dt >= budget_s and n >= min_queries) or n >= 4096:
            break
    qps_sample = n / dt
    if _limit is not None:
        _limit.restore_original_limits()
    return {
        "value": qps_sample * sample_rows / total_rows,
        "unit": "queries/s",
        "cores": int(os.cpu_count() or 1),
        "threads": int(threads),
        "kind": "port",
        "sample": f"{n} single queries over a {sample_rows}x{dim} row sample (fp32 view of the bf16 corpus, numpy "
                  f"matmul+argpartition, {dt:.1f}s), QPS scaled by {sample_rows}/{total_rows} rows",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps_total = max(1, args.steps)
    t0 = time.perf_counter()
    base = cpu_reference_qps(args.rows, args.dim, args.k, budget_s=min(60.0, 4.0 * steps_total))
    wall = time.perf_counter() - t0
    # the index the reference really queries is chromadb's HNSW (approximate): its recall against the
    # exact result and its CPU query rate, from the hnswlib restatement in oracle/ on a bounded sample
    hnsw = None
    if args.hnsw_rows > 0:
        try:
            from oracle import hnsw_oracle
            hnsw = [hnsw_oracle.measure(args.hnsw_rows, args.dim, 200, args.k, clustered=True),
                    hnsw_oracle.measure(args.hnsw_rows, args.dim, 200, args.k, clustered=False)]
        except Exception as e:                        # noqa: BLE001 -- gcc missing etc.: report, do not fail the arm
            hnsw = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / base["value"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall, "hnsw_recall_vs_exact": hnsw,
        "note": "chromadb/hnswlib are not installable offline; this is the reference's exact CPU path "
                "(numpy matmul + top-k) as restated in oracle/, all host threads BLAS uses",
    }
    print(json.dumps(line))


def workload_config(args, G):
    return {"workload": f"{args.rows}x{args.dim} {args.dtype} unit-norm corpus, exact top-{args.k} cosine, "
                        f"{args.queries} single-query scans per step",
            "rows_total": args.rows, "rows_per_gpu": (args.rows + G - 1) // G, "dim": args.dim, "k": args.k,
            "queries_per_step": args.queries, "parallelism": f"row-shard x{G}",
            "l2": "inputs larger than L2 (no flush needed)"}


# ------------------------------------------------------------------------------------------------
# clocks sampling (recipe: /opt/skills/guides/B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region: an NVML polling thread
    (5 ms period; `nvidia-smi -lms` needs ~100 ms to start, too slow for short multi-GPU runs)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        import threading
        self.samples, self.err, self._stop = [], None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception as e:                        # noqa: BLE001
            self.err = f"nvml unavailable: {type(e).__name__}"
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                     int(self.reasons_fn(self.h))))
            except Exception as e:                    # noqa: BLE001
                self.err = f"nvml sample failed: {type(e).__name__}"
                return
            self._stop.wait(0.005)

    def stop(self):
        if self.err and not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err]}
        self._stop.set()
        self.t.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        mask = 0
        for _, _, m in self.samples:
            mask |= m
        return {"sm_mhz": float(np.median([c for c, _, _ in self.samples])), "sm_max_mhz": self.max_mhz,
                "power_w_max": float(max(p for _, p, _ in self.samples)), "samples": len(self.samples),
                "reasons": sorted(name for bit, name in self.REASONS.items() if mask & bit)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import mmiss_b200 as M

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    G = world
    lo, hi = M.shard_bounds(args.rows, G, rank)
    n_local = hi - lo

    # ---- build the shard (synthetic unit-norm rows, generated on the device in chunks) --------
    ix = M.DeviceIndex(args.dim, args.dtype, device=local_rank, capacity=n_local, row_base=lo)
    chunk = 1 << 19
    gen = torch.Generator(device=dev)
    t_build = time.perf_counter()
    for c0 in range(lo, hi, chunk):
        n = min(chunk, hi - c0)
        gen.manual_seed(1234 + c0)               # chunk content depends only on its global offset
        x = torch.randn((n, args.dim), generator=gen, device=dev, dtype=torch.float32)
        x = torch.nn.functional.normalize(x, dim=1)
        ix.add(x)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    assert len(ix) == n_local
    del x

    Q = args.queries
    gq = torch.Generator(device="cpu").manual_seed(99)
    q_host = torch.nn.functional.normalize(torch.randn((Q, args.dim), generator=gq), dim=1).pin_memory()
    q_dev = q_host.to(dev)
    k = args.k
    cand_s = torch.empty((Q, k), dtype=torch.float32, device=dev)
    cand_r = torch.empty((Q, k), dtype=torch.int64, device=dev)
    searcher, exchange = None, "none"
    if G > 1:
        exchange = args.exchange
        if exchange == "p2p":
            try:
                searcher = M.ShardedSearcher.for_index(ix, mode="scan", exchange="p2p",
                                                       b_max=max(Q, args.batch, 1), k_max=max(32, args.k))
            except Exception as e:                                   # e.g. no peer access between the GPUs
                exchange = f"nccl (p2p unavailable: {type(e).__name__}: {e})"[:200]
            flag = torch.tensor([1 if searcher is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)              # all ranks use the same exchange
            if int(flag.item()) == 0:
                if searcher is not None:
                    exchange = "nccl (p2p unavailable on a peer)"
                searcher = None
        if searcher is None:
            searcher = M.ShardedSearcher.for_index(ix, mode="scan", exchange="nccl")
    p2p = searcher is not None and searcher.exchange == "p2p"
    gath_s = torch.empty((G, Q, k), dtype=torch.float32, device=dev) if G > 1 else None
    gath_r = torch.empty((G, Q, k), dtype=torch.int64, device=dev) if G > 1 else None
    out_s = torch.empty((Q, k), dtype=torch.float32, device=dev)
    out_r = torch.empty((Q, k), dtype=torch.int64, device=dev)
    scan_ev = []

    def step(timed: bool):
        if timed:
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
        if p2p:        # every scan kernel pushes its top-k to all peers over NVLink; one collect kernel per step
            ix.exchange_begin()
            for i in range(Q):
                ix.query_push_dev(q_dev[i:i + 1], k, i, mode="scan")
        else:
            for i in range(Q):                   # Q independent single-query scans (one kernel each)
                ix.query_dev(q_dev[i:i + 1], k, out_scores=cand_s[i:i + 1], out_rows=cand_r[i:i + 1], mode="scan")
        if timed:
            e1.record()
            scan_ev.append((e0, e1))
        if p2p:
            return ix.exchange_collect_dev(Q, k, out_scores=out_s, out_rows=out_r)
        if G > 1:
            dist.all_gather_into_tensor(gath_s, cand_s)
            dist.all_gather_into_tensor(gath_r, cand_r)
            ix.merge_dev(gath_s, gath_r, out_scores=out_s, out_rows=out_r)
            return out_s, out_r
        return cand_s, cand_r

    def barrier():
        if G > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        res_s, res_r = step(False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = M.launch_count()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        res_s, res_r = step(True)
    t1.record()
    barrier()
    elapsed_ms = t0.elapsed_time(t1)
    launches = M.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    scan_ms = sum(a.elapsed_time(b) for a, b in scan_ev) / (len(scan_ev) * Q)

    # ---- correctness spot-check of the timed result (first 2 queries vs torch fp32 on this shard
    #      is not possible globally without the full corpus; check local candidates instead) ----
    rows_bytes = n_local * (args.dim * (2 if args.dtype == "bf16" else 4) + 4)

    # ---- e2e: host buffers through the public API, copies inside the timed region --------------
    # (a) per step (the contract's definition): the step's Q queries go host->device in ONE copy from
    #     pinned memory, are scored by Q independent single-query scans (one launch, grid.y = Q: every
    #     query still streams the whole shard by itself), and the [Q, k] result is read back.
    # (b) per query (a search service answering one request at a time): copy, scan, read back, sync.
    e2e_steps = max(1, min(args.steps, 5))
    h_out_s = torch.empty((Q, k), dtype=torch.float32).pin_memory()
    h_out_r = torch.empty((Q, k), dtype=torch.int64).pin_memory()
    q_np = q_host.numpy()

    def e2e_step():
        if G == 1:
            return ix.query(q_np, k, mode="scan")                 # vs_query_topk_host: H2D + Q scans + D2H + sync
        qd = q_host.to(dev, non_blocking=True)                    # H2D from pinned memory
        s, r = searcher.search(qd, k)                             # Q scans + exchange (+ merge)
        h_out_s.copy_(s, non_blocking=True); h_out_r.copy_(r, non_blocking=True)
        torch.cuda.synchronize()
        return h_out_s, h_out_r

    def e2e_query_step():
        for i in range(Q):
            if G == 1:
                ix.query(q_np[i:i + 1], k, mode="scan")
            elif p2p:
                ix.query_sharded(q_np[i:i + 1], k, mode="scan")   # vs_query_topk_sharded_host: H2D + kernel/exchange + D2H + sync
            else:
                qd = q_host[i:i + 1].to(dev, non_blocking=True)
                s, r = searcher.search(qd, k)
                h_out_s[i:i + 1].copy_(s, non_blocking=True); h_out_r[i:i + 1].copy_(r, non_blocking=True)
                torch.cuda.synchronize()

    def wall(fn, reps):
        fn()
        barrier()
        w0 = time.perf_counter()
        for _ in range(reps):
            fn()
        barrier()
        return time.perf_counter() - w0

    e2e_res = e2e_step()
    # the timed device result and the host-API result of the same queries must agree
    assert np.array_equal(np.asarray(e2e_res[1]), res_r.cpu().numpy()), "e2e result differs from the device-timed result"
    e2e_s = wall(e2e_step, e2e_steps)
    e2e_q_s = wall(e2e_query_step, e2e_steps)

    # ---- extra (not the headline): BASELINE config 3, B blended text+image queries on tcgen05 --------
    batched_ms = 0.0
    if args.dtype == "bf16" and args.batch > 0:
        B = args.batch
        gb = torch.Generator(device=dev).manual_seed(4242)          # same queries on every rank
        img = torch.randn((B, args.dim), generator=gb, device=dev)
        txt = torch.randn((B, args.dim), generator=gb, device=dev)
        wts = torch.rand((B,), generator=gb, device=dev, dtype=torch.float64)
        qb = torch.empty((B, args.dim), device=dev)
        bs = torch.empty((B, k), dtype=torch.float32, device=dev)
        br = torch.empty((B, k), dtype=torch.int64, device=dev)
        bgs = torch.empty((G, B, k), dtype=torch.float32, device=dev) if G > 1 else None
        bgr = torch.empty((G, B, k), dtype=torch.int64, device=dev) if G > 1 else None
        bos = torch.empty((B, k), dtype=torch.float32, device=dev)
        bor = torch.empty((B, k), dtype=torch.int64, device=dev)

        def bstep():
            ix.blend_dev(img, txt, wts, out=qb)                      # multimodal blend (main.py:850-860)
            if p2p:
                ix.query_sharded_dev(qb, k, out_scores=bos, out_rows=bor, mode="tensor")   # K2 + exchange kernel
                return
            ix.query_dev(qb, k, out_scores=bs, out_rows=br, mode="tensor")
            if G > 1:
                dist.all_gather_into_tensor(bgs.view(-1, k), bs)
                dist.all_gather_into_tensor(bgr.view(-1, k), br)
                ix.merge_dev(bgs, bgr, out_scores=bos, out_rows=bor)

        for _ in range(3):
            bstep()
        barrier()
        b0 = torch.cuda.Event(enable_timing=True); b1 = torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(5):
            bstep()
        b1.record()
        barrier()
        batched_ms = b0.elapsed_time(b1) / 5

    # ---- extras (not the headline): BASELINE configs 4 and 5 on this rank's shard -------------------------
    filter_ms, dedup_ms, dedup_rows, dedup_dim = 0.0, 0.0, 0, 768
    if args.dtype == "bf16" and args.filters > 0:
        gf = torch.Generator(device=dev).manual_seed(4343)
        prompts = torch.randn((args.filters, args.dim), generator=gf, device=dev)
        fbits = torch.zeros((args.filters, ix.filter_words()), dtype=torch.int32, device=dev)
        for _ in range(3):
            ix.filter_sweep_dev(prompts, 0.103, out_bits=fbits)           # ~1 % of random 512-d rows pass
        barrier()
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(5):
            ix.filter_sweep_dev(prompts, 0.103, out_bits=fbits)
        f1.record()
        barrier()
        filter_ms = f0.elapsed_time(f1) / 5
        del fbits
    if args.dtype == "bf16" and args.dedup_rows > 0 and G == 1:
        dedup_rows = args.dedup_rows
        dx = M.DeviceIndex(dedup_dim, "bf16", device=local_rank, capacity=dedup_rows)
        for c0 in range(0, dedup_rows, chunk):
            n = min(chunk, dedup_rows - c0)
            gen.manual_seed(777 + c0)
            dx.add(torch.nn.functional.normalize(torch.randn((n, dedup_dim), generator=gen, device=dev), dim=1))
        cap = 1 << 16
        oi = torch.empty(cap, dtype=torch.int64, device=dev); oj = torch.empty(cap, dtype=torch.int64, device=dev)
        osc = torch.empty(cap, dtype=torch.float32, device=dev); cnt = torch.zeros(2, dtype=torch.int64, device=dev)
        dx.dedup_dev(0.95, 0, dedup_rows, oi, oj, osc, cnt)
        torch.cuda.synchronize()
        d0 = torch.cuda.Event(enable_timing=True); d1 = torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(2):
            dx.dedup_dev(0.95, 0, dedup_rows, oi, oj, osc, cnt)
        d1.record()
        torch.cuda.synchronize()
        dedup_ms = d0.elapsed_time(d1) / 2
        dx.close()

    # ---- reduce over ranks (max time) -----------------------------------------------------------------
    xerr = ix.exchange_error() if p2p else 0
    if xerr:
        raise SystemExit(f"bench.py: peer exchange timed out on rank {rank} (results invalid)")
    tvals = torch.tensor([elapsed_ms, scan_ms, e2e_s, batched_ms, e2e_q_s, filter_ms], dtype=torch.float64, device=dev)
    if G > 1:
        dist.all_reduce(tvals, op=dist.ReduceOp.MAX)
    elapsed_ms, scan_ms, e2e_s, batched_ms, e2e_q_s, filter_ms = tvals.tolist()

    if rank == 0:
        peaks, peaks_kind = measured_peaks()
        qps = Q * args.steps / (elapsed_ms / 1e3)
        achieved = rows_bytes / (scan_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": G, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, G),
            "scanned_gb_per_s": qps * args.rows * args.dim * (2 if args.dtype == "bf16" else 4) / 1e9,
            "roofline": {"bound": "hbm", "kernel": "scan_topk_kernel", "achieved": achieved,
                         "peak": peaks["hbm_gbs"], "peak_kind": peaks_kind, "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"],
                         "traffic": ncu_traffic_bytes(n_local, args.dim, args.dtype),
                         "bytes_per_launch": rows_bytes, "avg_launch_ms": scan_ms,
                         "note": "peak = MEASURED_PEAKS.json hbm_gbs, a read+write COPY bandwidth; this kernel only "
                                 "reads, so frac can exceed 1 (ncu: gpu__dram_throughput 88.8 % of the DRAM peak, "
                                 "profiles/r01c_ncu_scan_topk.txt); avg_launch_ms is per scan inside a burst of "
                                 "back-to-back launches (consecutive scans overlap under PDL)"},
            "e2e": {"value": Q * e2e_steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": Q * args.dim * 4,
                    "d2h_bytes_per_step": Q * k * 12, "steps": e2e_steps,
                    "path": ("vs_query_topk_host(B=%d, scan) via ctypes" % Q) if G == 1 else
                    f"ShardedSearcher.search ({'scan + fused p2p exchange' if p2p else 'scan + nccl all-gather + merge'}) "
                    "+ pinned H2D/D2H",
                    "per_query_sync": {"value": Q * e2e_steps / e2e_q_s, "unit": "queries/s",
                                       "note": "one request at a time: H2D, scan, D2H, sync per query"}},
            "gpu_launches": int(launches), "clocks": clocks, "build_s": t_build, "exchange": exchange,
        }
        if batched_ms > 0:
            tf = 2.0 * args.batch * args.rows * args.dim / (batched_ms / 1e3) / 1e12
            line["batched"] = {"workload": f"{args.batch} blended text+image queries, top-{k}, tcgen05 path "
                                           f"(blend kernel + K2{(' + p2p exchange kernel' if p2p else ' + all-gather + merge') if G > 1 else ''})",
                               "ms_per_batch": batched_ms, "qps": args.batch / (batched_ms / 1e3), "tflops": tf,
                               "frac_of_bf16_burst": tf / G / peaks["bf16_tflops"],
                               "frac_of_bf16_sustained": tf / G / peaks["bf16_tflops_sustained"]}
        if filter_ms > 0:
            tf = 2.0 * args.filters * args.rows * args.dim / (filter_ms / 1e3) / 1e12
            fbytes = args.rows * args.dim * 2 + args.filters * args.rows / 8
            line["filter_sweep"] = {"workload": f"{args.filters} filter prompts x {args.rows} rows, cos >= 0.103 -> bit mask "
                                                "(tcgen05, no exchange: mask rows are shard-local)",
                                    "ms": filter_ms, "tflops": tf, "frac_of_bf16_burst": tf / G / peaks["bf16_tflops"],
                                    "frac_of_bf16_sustained": tf / G / peaks["bf16_tflops_sustained"],
                                    "gb_per_s": fbytes / (filter_ms / 1e3) / 1e9,
                                    "frac_of_hbm": fbytes / (filter_ms / 1e3) / 1e9 / G / peaks["hbm_gbs"]}
        if dedup_ms > 0:
            tf = float(dedup_dim) * dedup_rows * (dedup_rows - 1) / (dedup_ms / 1e3) / 1e12
            line["dedup"] = {"workload": f"all pairs cos >= 0.95 over {dedup_rows} x {dedup_dim} bf16 rows (tcgen05; "
                                         "useful-triangle flops; the 8-GPU 2M-row run is tools/bench_dedup_sharded.py)",
                             "ms": dedup_ms, "tflops": tf, "frac_of_bf16_burst": tf / peaks["bf16_tflops"],
                             "frac_of_bf16_sustained": tf / peaks["bf16_tflops_sustained"]}
        if G == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_qps(args.rows, args.dim, k, budget_s=args.cpu_budget)
        print(json.dumps(line))
    ix.close()
    if G > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=TOTAL_ROWS)
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--k", type=int, default=TOPK)
    ap.add_argument("--queries", type=int, default=32, help="single-query scans per step")
    ap.add_argument("--batch", type=int, default=1024, help="batched tensor-path extra (0 = skip)")
    ap.add_argument("--filters", type=int, default=256, help="filter-sweep extra: number of prompts (0 = skip)")
    ap.add_argument("--dedup-rows", type=int, default=200_000, help="dedup extra (N=1 only): rows x 768 (0 = skip)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: candidate exchange fused into the query kernel over NVLink peer memory, or NCCL all-gather")
    ap.add_argument("--hnsw-rows", type=int, default=20_000,
                    help="reference arm: rows of the bounded HNSW recall sample (0 = skip)")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
