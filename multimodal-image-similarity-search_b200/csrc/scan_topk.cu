// scan_topk.cu -- K1: single-query exact cosine top-k as ONE fused kernel (sm_100a).
//
// Replaces the arithmetic of chromadb Collection.query for one vector
// (reference call site backend/app/main.py:761-765).  The corpus shard streams from HBM
// exactly once per query:
//   * a producer thread per CTA moves tiles of R whole rows (R*pitch contiguous bytes) plus
//     their R inverse norms into a ring of shared-memory stages with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP), L2 evict_first -- except a fixed ~64 MB slice of the
//     shard that is loaded evict_last and so stays L2-resident from one query to the next (kL2KeepBytes);
//   * 8 consumer warps read the staged rows with conflict-free 128-bit LDS, FMA against the
//     L2-normalised query held in registers (query normalisation is fused: every warp derives
//     1/||q|| itself), butterfly-reduce, scale by the row's inverse norm;
//   * each warp keeps a register-resident sorted top-k list distributed over its lanes
//     (WarpTopK<M>, k <= 32*M) updated with warp shuffles; nearly every row fails the
//     `score > k-th` test so the list code is off the hot path;
//   * filter bits ("pre" mode of the filter pass, backend/app/main.py:215) are only looked
//     up for rows that would enter the list -- zero extra traffic;
//   * the per-warp lists are merged per CTA, written to a small global buffer, and the last
//     CTA to finish (ticket) merges all CTA lists and writes the final [k] result: no second
//     kernel, no score materialisation.  For k <= 32 the lists on this tail are sorted runs of packed
//     64-bit keys in shared memory, merged by a register-held bitonic merge tree (8/16 lists, k <= 16:
//     WarpTopK::merge_sorted_bitonic) or by cursor selection (WarpTopK::select_sorted_smem): the tail is
//     pure latency for an isolated request (tools/scan_stamps.py, profiles/r02_group_latency.md).
// Algorithmic HBM bytes per query = n_rows * (pitch + 4).
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "exchange.cuh"
#include "kernels.h"

namespace vs {

constexpr int kMaxStages = 8;
constexpr int kMaxScanDevices = 64;
// Shared memory per CTA.  Deliberately about a third of the SM: two CTAs fit on one SM, so with
// programmatic dependent launch the NEXT query's CTAs are already resident and streaming while this
// query's CTAs run their merge/exchange tails -- the HBM pipe never drains between queries
// (profiles/r01_scan_ab.md: 6.4 -> 7.3 TB/s on back-to-back 10M x 512 bf16 scans).
constexpr int kSmemBudget = 76 * 1024;
constexpr int kCtasPerSm = 2;
constexpr int kSmemMax = 220 * 1024;
// Bytes of each shard kept L2-resident across queries (evict_last tiles, see the producer loop).  0 disables.
constexpr int64_t kL2KeepBytes = 64ll << 20;

// -DVS_SCAN_STAMPS (tuning builds only, tools/scan_stamps.py): %globaltimer marks of one query's phases
#ifdef VS_SCAN_STAMPS
__device__ unsigned long long g_scan_stamps[8];
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define VS_STAMP_MIN(i) atomicMin(&g_scan_stamps[i], gtime_ns())
#define VS_STAMP_MAX(i) atomicMax(&g_scan_stamps[i], gtime_ns())
#else
#define VS_STAMP_MIN(i) ((void)0)
#define VS_STAMP_MAX(i) ((void)0)
#endif

// rows per consumer warp per tile, chosen so a stage (W warps x U rows x pitch) is <= 32 KB
#ifndef VS_SCAN_TILE_SHIFT
#define VS_SCAN_TILE_SHIFT 0   // experiments: 1 halves the tile (rows per warp), see profiles/r01_scan_ab.md
#endif
__host__ __device__ constexpr int rows_per_tile_target(int cpl) {
  return (cpl <= 1 ? 64 : cpl <= 2 ? 32 : cpl <= 4 ? 16 : 8) >> VS_SCAN_TILE_SHIFT;
}
__host__ __device__ constexpr int rows_per_warp(int cpl, int w) {
  return rows_per_tile_target(cpl) / w > 0 ? rows_per_tile_target(cpl) / w : 1;
}

struct ScanKernelParams {
  const uint8_t* rows;
  const float* inv_norm;
  const uint64_t* mask;
  uint64_t req[kMaskWords];
  const float* q;
  float* part_s;
  uint32_t* part_r;
  unsigned int* tickets;
  float* out_s;
  int64_t* out_r;
  float* scores_full;
  int64_t row_base;
  int64_t row_stride;
  uint32_t n_rows;
  uint32_t n_tiles;
  int dim;
  int ld_bytes;
  int k;
  int stages;
  int stage_stride;   // bytes between row stages
  int use_mask;
  int early_wait;     // 1: griddepcontrol.wait before the first read of q / the corpus (see launch_one)
  unsigned int* done_flag;   // host-mapped [B] or nullptr: done_flag[qi] = done_seq once query qi's result is written
  unsigned int done_seq;
  XchgParams xg;      // xg.G > 0: exchange the shard's result with the peers before writing it
  uint32_t keep_mod;  // > 0: every keep_mod-th tile of each CTA is loaded L2 evict_last (the shard's L2-resident slice)
  int q_inline;       // 1: the (single) query travels IN the launch packet (qv) instead of through device memory
  float qv[kMaxInlineQ];
};

// Final step of a query, executed by ONE warp holding the shard's top-k (local rows): either write
// it out, or (row-sharded collection) push it to every peer over NVLink, wait for the peers'
// lists and write the merged GLOBAL top-k.  Exactly one warp per query reaches this point, and it
// never waits before its own push, so ranks cannot deadlock on each other.
template <int ML>
__device__ __forceinline__ void emit_result(const ScanKernelParams& p, WarpTopK<ML>& top, int qi, int k, int lane,
                                            uint64_t* sm_keys, int sm_entries) {
  if (p.xg.G > 0) {
#pragma unroll
    for (int m = 0; m < ML; ++m)
      if (top.r[m] != kEmptyRow) top.r[m] = top.r[m] * (uint32_t)p.row_stride + (uint32_t)p.row_base;   // global rows < 2^32 (checked on the host)
    xchg_push(p.xg, top, p.xg.slot0 + qi, k, lane);
    if (p.xg.push_only) return;   // vs_exchange_collect_dev merges the whole epoch later
    xchg_wait_merge(p.xg, top, p.xg.slot0 + qi, k, lane, sm_keys, sm_entries);
  }
  const int64_t add = p.xg.G > 0 ? 0 : p.row_base;
  const int64_t mul = p.xg.G > 0 ? 1 : p.row_stride;
  for (int e = lane; e < k; e += 32) {
#pragma unroll
    for (int m = 0; m < ML; ++m)
      if ((e >> 5) == m) {
        p.out_s[(size_t)qi * k + e] = top.s[m];
        p.out_r[(size_t)qi * k + e] = top.r[m] == kEmptyRow ? -1 : (int64_t)top.r[m] * mul + add;
      }
  }
  if (p.done_flag) {
    // request/response without a stream synchronise: out_s/out_r (and the flag) live in host-mapped
    // pinned memory; the host thread polls the flag (vs_group_query_host)
    __threadfence_system();
    __syncwarp();
    if (lane == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.done_flag + qi), "r"(p.done_seq) : "memory");
  }
}

template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int kPerChunk = 4;
  static __device__ __forceinline__ float dot(const uint4& v, const float* q) {
    return fmaf(__uint_as_float(v.w), q[3],
                fmaf(__uint_as_float(v.z), q[2], fmaf(__uint_as_float(v.y), q[1], __uint_as_float(v.x) * q[0])));
  }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kPerChunk = 8;
  // bf16 -> f32 is a 16-bit shift: low half `x << 16`, high half `x & 0xffff0000`
  static __device__ __forceinline__ float dot(const uint4& v, const float* q) {
    float a = __uint_as_float(v.x << 16) * q[0];
    a = fmaf(__uint_as_float(v.x & 0xffff0000u), q[1], a);
    a = fmaf(__uint_as_float(v.y << 16), q[2], a);
    a = fmaf(__uint_as_float(v.y & 0xffff0000u), q[3], a);
    a = fmaf(__uint_as_float(v.z << 16), q[4], a);
    a = fmaf(__uint_as_float(v.z & 0xffff0000u), q[5], a);
    a = fmaf(__uint_as_float(v.w << 16), q[6], a);
    a = fmaf(__uint_as_float(v.w & 0xffff0000u), q[7], a);
    return a;
  }
};

__device__ __forceinline__ bool mask_ok(const uint64_t* mask, uint32_t row, const uint64_t* req) {
  const uint64_t* m = mask + (size_t)row * kMaskWords;
  bool ok = true;
#pragma unroll
  for (int w = 0; w < kMaskWords; ++w) ok = ok && ((__ldg(m + w) & req[w]) == req[w]);
  return ok;
}

// M == 0: materialise scores (large-k path) instead of keeping lists.
template <typename T, int CPL, int M, int W, bool FULL>
__global__ void __launch_bounds__((W + 1) * 32, (M == 4 && W == 16) ? 1 : 2) scan_topk_kernel(const ScanKernelParams p) {
  constexpr int kConsumerWarps = W;
  constexpr int EPC = Elem<T>::kPerChunk;
  constexpr int U = rows_per_warp(CPL, W);
  constexpr int R = kConsumerWarps * U;
  constexpr int ML = M > 0 ? M : 1;

  extern __shared__ __align__(128) uint8_t smem[];
  // layout: [stages][stage_stride] rows | [kMaxStages][R] inv norms | barriers | candidate lists
  uint8_t* s_rows = smem;
  float* s_inv = reinterpret_cast<float*>(smem + (size_t)p.stages * p.stage_stride);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_inv + kMaxStages * R);
  uint64_t* empty = full + kMaxStages;
  float* cand_s = reinterpret_cast<float*>(empty + kMaxStages);
  uint32_t* cand_r = reinterpret_cast<uint32_t*>(cand_s + kConsumerWarps * 32 * ML);
  __shared__ int s_is_last;
  // a query that rides in the launch packet is staged through shared memory once per CTA: reading the constant bank
  // with a lane-varying index costs one pass per distinct address, and every consumer warp would pay it again
  __shared__ float s_q[kMaxInlineQ];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.y;
  const int k = p.k;
  const int chunks = p.ld_bytes >> 4;

  if (threadIdx.x == 0) {
    VS_STAMP_MIN(0);   // first / last CTA entering
    VS_STAMP_MAX(1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();
  // PDL: the next query's kernel may start filling SMs as this grid's CTAs retire.  This kernel
  // only READS the corpus/query until the partial lists are written, so -- when the host knows that
  // nothing this kernel reads was produced by the grid right before it in the stream (early_wait = 0:
  // the previous launch was a scan of this library and the caller vouches for q, VS_Q_PIPELINED) --
  // the wait on the previous grid (which may still be merging into the same workspace) is deferred
  // to that point.  Otherwise the wait comes first (producer: before its first load; consumers: before
  // they read q): the PDL contract makes the primary's writes visible only after griddepcontrol.wait.
  pdl_launch_dependents();
  const bool wait_first = (M == 0) || p.early_wait;   // (the materialising variant writes shared scratch while scanning)

  WarpTopK<ML> top;
  top.init();

  if (warp == kConsumerWarps) {
    // ===================== producer: one thread drives the TMA bulk pipeline ==============
    if (lane == 0) {
      if (wait_first) pdl_wait();
      // Everything streams evict_first EXCEPT a fixed slice of the shard (tile `it` of CTA `x` with (it + x) % keep_mod
      // == 0, about kL2KeepBytes in total) that is loaded evict_last: a shard is many times the 126 MB L2, so without
      // a hint nothing survives from one query to the next; with it the same ~5 % of the tiles hit in L2 on EVERY
      // query and HBM only has to deliver the rest.  The slice is staggered over CTAs and iterations so that the hits
      // are spread over the whole scan (L2 serves them while HBM streams the other tiles) instead of front-loaded.
      const uint64_t pol_stream = policy_evict_first();
      const uint64_t pol_keep = policy_evict_last();
      uint32_t it = 0;
      uint32_t phase = p.keep_mod ? blockIdx.x % p.keep_mod : 1u;   // == (it + blockIdx.x) % keep_mod, kept incrementally
      for (uint32_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
        const uint64_t pol = (p.keep_mod && phase == 0) ? pol_keep : pol_stream;
        if (p.keep_mod && ++phase == p.keep_mod) phase = 0;
        const int st = it % p.stages;
        const uint32_t use = it / p.stages;
        if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
        const uint32_t row0 = t * R;
        const uint32_t nr = min((uint32_t)R, p.n_rows - row0);
        const uint32_t bytes_rows = nr * (uint32_t)p.ld_bytes;
        const uint32_t bytes_inv = ((nr + 3u) & ~3u) * 4u;
        mbar_expect_tx(&full[st], bytes_rows + bytes_inv);
        bulk_g2s(s_rows + (size_t)st * p.stage_stride, p.rows + (size_t)row0 * p.ld_bytes, bytes_rows, &full[st],
                 pol);
        bulk_g2s(s_inv + st * R, p.inv_norm + row0, bytes_inv, &full[st], pol);
      }
    }
    __syncwarp();
  } else {
    // ===================== consumers =======================================================
    // a query that rides in the launch packet is staged through shared memory by the consumer warps while the
    // producer's first tiles are already in flight (named barrier: the producer warp does not take part)
    if (p.q_inline) {
      for (int e = threadIdx.x; e < p.dim; e += kConsumerWarps * 32) s_q[e] = p.qv[e];
      asm volatile("bar.sync 1, %0;" ::"r"(kConsumerWarps * 32) : "memory");
    }
    if (wait_first) pdl_wait();
    // normalised query in registers: chunk c*32+lane of the row <-> q elements [ch*EPC, +EPC)
    float q[CPL][EPC];
    {
      const float* qp = p.q + (size_t)qi * p.dim;
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int e0 = (c * 32 + lane) * EPC;
#pragma unroll
        for (int e = 0; e < EPC; ++e) {
          const float v = (e0 + e < p.dim) ? (p.q_inline ? s_q[e0 + e] : __ldg(qp + e0 + e)) : 0.f;
          q[c][e] = v;
          ss = fmaf(v, v, ss);
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
      const float qinv = 1.0f / (sqrtf(ss) + 1e-30f);
#pragma unroll
      for (int c = 0; c < CPL; ++c)
#pragma unroll
        for (int e = 0; e < EPC; ++e) q[c][e] *= qinv;
    }

    uint32_t it = 0;
    for (uint32_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
      const int st = it % p.stages;
      mbar_wait(&full[st], (it / p.stages) & 1);
#ifdef VS_SCAN_STAMPS
      if (it == 0 && threadIdx.x == 0) {
        VS_STAMP_MIN(2);   // first tile landed: earliest / latest CTA
        VS_STAMP_MAX(3);
      }
#endif
      const uint8_t* tile = s_rows + (size_t)st * p.stage_stride + (size_t)(warp * U) * p.ld_bytes;
      // all LDS.128 of the warp's U rows are issued before any FMA (U*CPL loads in flight);
      // FULL = every lane owns a valid chunk in every pass, so the loads are unpredicated
      uint4 v[U][CPL];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint4* rp = reinterpret_cast<const uint4*>(tile + (size_t)u * p.ld_bytes);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const int ch = c * 32 + lane;
          if (FULL || ch < chunks) v[u][c] = rp[ch];
          else v[u][c] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      float acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float part[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) part[c] = Elem<T>::dot(v[u][c], q[c]);   // U*CPL independent chains
        float a = part[0];
#pragma unroll
        for (int c = 1; c < CPL; ++c) a += part[c];
        acc[u] = a;
      }
      float inv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) inv[u] = s_inv[st * R + warp * U + u];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], off);
      // all lanes have consumed the stage (the shuffles above are warp-convergent)
      if (lane == 0) mbar_arrive(&empty[st]);

      const uint32_t row0 = t * R + warp * U;
      if constexpr (M == 0) {
        float mine = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (lane == u) mine = acc[u] * inv[u];
        const uint32_t row = row0 + lane;
        if (lane < U && row < p.n_rows) {
          if (p.use_mask && !mask_ok(p.mask, row, p.req)) mine = VS_NEG_INF;
          p.scores_full[(size_t)qi * p.n_rows + row] = mine;
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float s = acc[u] * inv[u];
          const uint32_t row = row0 + u;
          // rows arrive in increasing order within a warp, so on equal score the earlier row
          // stays: a strict compare implements the (score desc, row asc) order here.
          if (s > top.thr_s && row < p.n_rows) {
            if (!p.use_mask || mask_ok(p.mask, row, p.req)) top.insert(s, row, k - 1, lane);
          }
        }
      }
    }
  }

  if constexpr (M == 0) return;
  if (threadIdx.x == 0) {
    VS_STAMP_MIN(4);   // streaming done: earliest / latest CTA
    VS_STAMP_MAX(5);
  }
  if (!p.early_wait) pdl_wait();   // previous grid fully done: its partial lists / tickets / result rows are no longer in use

  // ---- per-CTA merge of the 8 warp lists ---------------------------------------------------
  // k <= 32 (ML == 1): lists travel through shared memory as packed 64-bit keys; the W warp lists are merged by a
  // register-held bitonic merge tree (k <= 16, WarpTopK::merge_sorted_bitonic) or by cursor selection
  // (WarpTopK::select_sorted_smem); the candidate area [W][32] x (f32 + u32) is exactly [W][32] keys.
  uint64_t* cand_k = reinterpret_cast<uint64_t*>(cand_s);
  if (warp < kConsumerWarps) {
    if constexpr (ML == 1) top.store_keys(cand_k + warp * 32, k, lane);
    else top.store(cand_s + warp * 32 * ML, cand_r + warp * 32 * ML, k, lane);
  }
  __syncthreads();
  const size_t pbase = ((size_t)qi * gridDim.x + blockIdx.x) * k;
  if (warp == 0) {
    // all warps' lists (this warp's own is list 0)
    if constexpr (ML == 1) {
      if (k <= 16) top.template merge_sorted_bitonic<kConsumerWarps>(cand_k, kConsumerWarps, 32, k, lane);
      else top.select_sorted_smem(cand_k, kConsumerWarps, 32, k, lane);
    } else {
      top.select_from(cand_s, cand_r, kConsumerWarps, 32 * ML, k, lane);
    }
    if (gridDim.x == 1) {
      emit_result<ML>(p, top, qi, k, lane, cand_k, kConsumerWarps * 32 * ML);   // single CTA: this is already the shard's answer
    } else {
      top.store(p.part_s + pbase, p.part_r + pbase, k, lane);
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        const unsigned int ticket = atomicAdd(&p.tickets[qi], 1u);
        s_is_last = (ticket == gridDim.x - 1);
      }
    }
  }
  if (gridDim.x == 1) return;
  __syncthreads();
  if (!s_is_last) return;

  // ---- last CTA: merge every CTA's list (8 warps in parallel, then warp 0) -----------------
  __threadfence();
  if (warp < kConsumerWarps) {
    top.init();
    const volatile float* gs = p.part_s + (size_t)qi * gridDim.x * k;
    const volatile uint32_t* gr = p.part_r + (size_t)qi * gridDim.x * k;
    // warp w takes CTA lists w, w+8, ...
    const int nl = ((int)gridDim.x - warp + kConsumerWarps - 1) / kConsumerWarps;
    if constexpr (ML == 1) {
      // The lists are copied (independent, coalesced loads; packed on the way) into this warp's slice of the row ring --
      // every tile has been consumed by now, the ring is free -- and selected there.  A slice holds `cap` lists; more
      // than that (large k) go in batches, with the running result carried along as list 0 of the next batch.
      const int slice_keys = (int)(((size_t)p.stages * p.stage_stride / kConsumerWarps) >> 3);
      uint64_t* lk = reinterpret_cast<uint64_t*>(s_rows) + (size_t)warp * slice_keys;
      int cap = slice_keys / k;
      if (cap > WarpTopK<ML>::kMaxSortedLists) cap = WarpTopK<ML>::kMaxSortedLists;
      int done = 0;
      if (cap < 2) {   // (cannot happen with the shipped tile shapes: a slice is >= 64 keys) -- register-scan form
        if (nl > 0) top.select_from(gs + (size_t)warp * k, gr + (size_t)warp * k, nl, kConsumerWarps * k, k, lane);
        done = nl;
      }
      while (done < nl) {
        const int carry = done > 0 ? 1 : 0;
        const int nb = nl - done < cap - carry ? nl - done : cap - carry;
        if (carry) top.store_keys(lk, k, lane);
        stage_lists_as_keys(gs + ((size_t)warp + (size_t)done * kConsumerWarps) * k, gr + ((size_t)warp + (size_t)done * kConsumerWarps) * k,
                            nb, (size_t)kConsumerWarps * k, k, lk + carry * k, lane);
        __syncwarp();
        top.select_sorted_smem(lk, nb + carry, k, k, lane);
        __syncwarp();
        done += nb;
      }
      top.store_keys(cand_k + warp * 32, k, lane);
    } else {
      if (nl > 0) top.select_from(gs + (size_t)warp * k, gr + (size_t)warp * k, nl, kConsumerWarps * k, k, lane);
      top.store(cand_s + warp * 32 * ML, cand_r + warp * 32 * ML, k, lane);
    }
  }
  __syncthreads();
  if (warp == 0) {
    if constexpr (ML == 1) {
      if (k <= 16) top.template merge_sorted_bitonic<kConsumerWarps>(cand_k, kConsumerWarps, 32, k, lane);
      else top.select_sorted_smem(cand_k, kConsumerWarps, 32, k, lane);
    } else {
      top.select_from(cand_s, cand_r, kConsumerWarps, 32 * ML, k, lane);
    }
    if (lane == 0) p.tickets[qi] = 0;  // ready for the next launch
    if (lane == 0) VS_STAMP_MAX(6);    // shard's top-k selected
    emit_result<ML>(p, top, qi, k, lane, cand_k, kConsumerWarps * 32 * ML);
    if (lane == 0) VS_STAMP_MAX(7);    // result (and flag) written
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int cpl_for(int64_t ld_bytes) {
  const int64_t chunks = ld_bytes / 16;
  const int need = (int)((chunks + 31) / 32);
  const int avail[] = {1, 2, 3, 4, 6, 8};
  for (int a : avail)
    if (need <= a) return a;
  return -1;
}

// consumer warps per CTA, from the B200 A/B run (profiles/r01_scan_ab.md): bf16 rows carry twice the
// FMA work per byte and run best with 8 warps x 4 rows in flight; f32 prefers 16 warps.
static int scan_warps(int dtype) { return dtype == 0 ? 16 : 8; }

int scan_rows_per_tile(int dtype, int64_t ld_bytes) {
  const int cpl = cpl_for(ld_bytes);
  return cpl < 0 ? -1 : scan_warps(dtype) * rows_per_warp(cpl, scan_warps(dtype));
}

template <typename T, int CPL, int M, int W, bool FULL>
static cudaError_t launch_one(const ScanArgs& a, int sm_count, cudaStream_t st) {
  constexpr int kConsumerWarps = W;
  constexpr int kScanThreads = (W + 1) * 32;
  constexpr int U = rows_per_warp(CPL, W);
  constexpr int R = kConsumerWarps * U;
  constexpr int ML = M > 0 ? M : 1;
  ScanKernelParams p;
  p.rows = static_cast<const uint8_t*>(a.rows);
  p.inv_norm = a.inv_norm;
  p.mask = a.mask;
  p.use_mask = 0;
  for (int w = 0; w < kMaskWords; ++w) {
    p.req[w] = a.req[w];
    if (a.req[w]) p.use_mask = 1;
  }
  if (!a.mask) p.use_mask = 0;
  p.q = a.q;
  p.q_inline = 0;
  if (a.q_host != nullptr) {
    if (a.B != 1 || a.dim > kMaxInlineQ) return cudaErrorInvalidValue;
    p.q_inline = 1;
    memcpy(p.qv, a.q_host, (size_t)a.dim * sizeof(float));
  }
  p.part_s = a.part_s;
  p.part_r = a.part_r;
  p.tickets = a.tickets;
  p.out_s = a.out_s;
  p.out_r = a.out_r;
  p.scores_full = a.scores_full;
  p.row_base = a.row_base;
  p.row_stride = a.row_stride;
  p.n_rows = (uint32_t)a.n_rows;
  p.n_tiles = (uint32_t)((a.n_rows + R - 1) / R);
  p.dim = a.dim;
  p.ld_bytes = (int)a.ld_bytes;
  p.k = a.k;
  p.early_wait = a.early_wait;
  p.done_flag = a.done_flag;
  p.done_seq = a.done_seq;
  p.xg = a.xg;
  p.stage_stride = (int)((R * a.ld_bytes + 127) & ~127LL);
  {
#ifdef VS_TUNING
    static const int64_t keep_bytes = [] {
      const char* v = getenv("VS_SCAN_KEEP_MB");
      return v ? (int64_t)atoi(v) << 20 : kL2KeepBytes;
    }();
#else
    constexpr int64_t keep_bytes = kL2KeepBytes;
#endif
    const int64_t shard_bytes = a.n_rows * a.ld_bytes;
    p.keep_mod = keep_bytes > 0 ? (uint32_t)((shard_bytes + keep_bytes - 1) / keep_bytes) : 0u;
    if (keep_bytes > 0 && p.keep_mod < 1) p.keep_mod = 1;
  }
  const int fixed = kMaxStages * R * 4 + 2 * kMaxStages * 8 + kConsumerWarps * 32 * ML * 8 + 256;
  // the product library has no knobs; a tuning build (-DVS_TUNING, build.py) reads them for A/B runs (tools/bench_scan.py)
#ifdef VS_TUNING
  static const int smem_budget = [] {
    const char* v = getenv("VS_SCAN_SMEM_KB");
    const int kb = v ? atoi(v) : 0;
    return kb >= 16 && kb <= 220 ? kb * 1024 : kSmemBudget;
  }();
  static const int ctas_per_sm = [] {
    const char* v = getenv("VS_SCAN_CTAS_PER_SM");
    const int c = v ? atoi(v) : 0;
    return c >= 1 && c <= 4 ? c : kCtasPerSm;
  }();
  static const bool pdl = [] {
    const char* v = getenv("VS_SCAN_PDL");
    return !(v && v[0] == '0');
  }();
#else
  constexpr int smem_budget = kSmemBudget;
  constexpr int ctas_per_sm = kCtasPerSm;
  constexpr bool pdl = true;
#endif
  int stages = (smem_budget - fixed) / p.stage_stride;
  if (stages < 2) stages = (kSmemMax - fixed) / p.stage_stride < 3 ? (kSmemMax - fixed) / p.stage_stride : 3;  // wide rows
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return cudaErrorInvalidValue;
  p.stages = stages;
  const size_t smem = (size_t)stages * p.stage_stride + fixed;
  auto kern = scan_topk_kernel<T, CPL, M, W, FULL>;
  cudaError_t e = cudaSuccess;
  {
    // the dynamic-smem attribute is per (kernel, device): set it once, not on every query (one driver call less on the
    // request path, where the worker threads of a group launch on 8 devices at the same moment)
    static std::atomic<size_t> smem_set[kMaxScanDevices] = {};
    int dev = 0;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxScanDevices || smem_set[dev].load(std::memory_order_acquire) < smem) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      if (dev >= 0 && dev < kMaxScanDevices) smem_set[dev].store(smem, std::memory_order_release);
    }
  }
  int gx = (a.grid_x > 0 ? a.grid_x : sm_count) * ctas_per_sm;
  if ((uint32_t)gx > p.n_tiles) gx = (int)p.n_tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, a.B, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kScanThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  e = cudaLaunchKernelEx(&cfg, kern, p);
  count_launch();
  return e;
}

template <typename T, int CPL, int W, bool FULL>
static cudaError_t launch_mw(const ScanArgs& a, int sm_count, cudaStream_t st) {
  if (a.scores_full) return launch_one<T, CPL, 0, W, FULL>(a, sm_count, st);
  if (a.k <= 32) return launch_one<T, CPL, 1, W, FULL>(a, sm_count, st);
  if (a.k <= 128) return launch_one<T, CPL, 4, W, FULL>(a, sm_count, st);
  return cudaErrorInvalidValue;
}
template <typename T, int CPL>
static cudaError_t launch_m(const ScanArgs& a, int sm_count, cudaStream_t st) {
  const bool full = (a.ld_bytes / 16) == CPL * 32;
  constexpr int W = sizeof(T) == 4 ? 16 : 8;
  return full ? launch_mw<T, CPL, W, true>(a, sm_count, st) : launch_mw<T, CPL, W, false>(a, sm_count, st);
}

template <typename T>
static cudaError_t launch_t(const ScanArgs& a, int sm_count, cudaStream_t st) {
  switch (cpl_for(a.ld_bytes)) {
    case 1: return launch_m<T, 1>(a, sm_count, st);
    case 2: return launch_m<T, 2>(a, sm_count, st);
    case 3: return launch_m<T, 3>(a, sm_count, st);
    case 4: return launch_m<T, 4>(a, sm_count, st);
    case 6: return launch_m<T, 6>(a, sm_count, st);
    case 8: return launch_m<T, 8>(a, sm_count, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_scan(const ScanArgs& a, int sm_count, cudaStream_t st) {
  if (a.n_rows <= 0 || a.B <= 0) return cudaErrorInvalidValue;
  if (a.ld_bytes % 16 != 0) return cudaErrorInvalidValue;
  if (a.dtype == 0) return launch_t<float>(a, sm_count, st);
  return launch_t<__nv_bfloat16>(a, sm_count, st);
}

}  // namespace vs

#ifdef VS_SCAN_STAMPS
// tuning builds only (not part of the ABI): reset != 0 re-arms the marks on the current device, else copies them out
extern "C" int vs_debug_scan_stamps(int reset, unsigned long long* out) {
  unsigned long long v[8];
  if (reset) {
    for (int i = 0; i < 8; ++i) v[i] = (i == 0 || i == 2 || i == 4) ? ~0ull : 0ull;
    return (int)cudaMemcpyToSymbol(vs::g_scan_stamps, v, sizeof(v));
  }
  const cudaError_t e = cudaMemcpyFromSymbol(v, vs::g_scan_stamps, sizeof(v));
  if (e == cudaSuccess && out) memcpy(out, v, sizeof(v));
  return (int)e;
}
#endif
