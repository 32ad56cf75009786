// common.cuh -- shared device helpers (sm_100a): mbarrier / bulk-copy (TMA) PTX wrappers,
// orderable score keys, warp-distributed top-k lists.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace vs {

constexpr int kWarp = 32;
constexpr uint32_t kEmptyRow = 0xFFFFFFFFu;
#define VS_NEG_INF (__int_as_float(0xff800000))

// ----------------------------------------------------------------------------------------
// mbarrier + bulk async copy (the non-tensor TMA path: SASS UBLKCP + SYNCS)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Programmatic dependent launch: let the next kernel in the stream start its CTAs as SMs free up
// (launch_dependents) and block until the previous grid has completed and flushed (wait).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0,
// both addresses 16-B aligned).  L2 policy: evict_first -- corpus rows are streamed once.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ----------------------------------------------------------------------------------------
// ranking order: (score desc, row asc).  better(a,b) <=> a ranks strictly before b.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ bool better(float sa, uint32_t ra, float sb, uint32_t rb) {
  return sa > sb || (sa == sb && ra < rb);
}
// monotone float -> uint32 map (larger float -> larger key); -0.0 and +0.0 map to distinct
// keys, so canonicalise zero first where keys are compared for equality.
__device__ __forceinline__ uint32_t score_key(float s) {
  uint32_t u = __float_as_uint(s + 0.0f);  // -0.0 + 0.0 = +0.0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_score(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(u);
}

// ----------------------------------------------------------------------------------------
// Warp-distributed sorted list of 32*M (score,row) entries: entry at rank p lives in slot
// p/32 of lane p%32.  All operations are warp-collective and branch warp-uniformly.
// ----------------------------------------------------------------------------------------
template <int M>
struct WarpTopK {
  float s[M];
  uint32_t r[M];
  float thr_s;      // entry k-1 (the current k-th best), warp-uniform
  uint32_t thr_r;

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int m = 0; m < M; ++m) {
      s[m] = VS_NEG_INF;
      r[m] = kEmptyRow;
    }
    thr_s = VS_NEG_INF;
    thr_r = kEmptyRow;
  }
  // would (ns,nr) enter the top-k?  (warp-uniform inputs)
  __device__ __forceinline__ bool accepts(float ns, uint32_t nr) const { return better(ns, nr, thr_s, thr_r); }

  // insert a warp-uniform candidate that passed accepts(); k-1 = index of the threshold entry.
  __device__ __forceinline__ void insert(float ns, uint32_t nr, int km1, int lane) {
    int pos = 0;
#pragma unroll
    for (int m = 0; m < M; ++m) pos += __popc(__ballot_sync(0xffffffffu, better(s[m], r[m], ns, nr)));
    float cs = ns;
    uint32_t cr = nr;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const int lo = m * 32;
      if (pos < lo + 32) {
        const int lp = pos > lo ? pos - lo : 0;
        const float last_s = __shfl_sync(0xffffffffu, s[m], 31);
        const uint32_t last_r = __shfl_sync(0xffffffffu, r[m], 31);
        const float up_s = __shfl_up_sync(0xffffffffu, s[m], 1);
        const uint32_t up_r = __shfl_up_sync(0xffffffffu, r[m], 1);
        if (lane > lp) {
          s[m] = up_s;
          r[m] = up_r;
        } else if (lane == lp) {
          s[m] = cs;
          r[m] = cr;
        }
        cs = last_s;
        cr = last_r;
      }
    }
    refresh_threshold(km1);
  }
  __device__ __forceinline__ void refresh_threshold(int km1) {
    float ts = VS_NEG_INF;
    uint32_t tr = kEmptyRow;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const float a = __shfl_sync(0xffffffffu, s[m], km1 & 31);
      const uint32_t b = __shfl_sync(0xffffffffu, r[m], km1 & 31);
      if ((km1 >> 5) == m) {
        ts = a;
        tr = b;
      }
    }
    thr_s = ts;
    thr_r = tr;
  }
  // write the first k entries to dst_s/dst_r (smem or global), coalesced.
  __device__ __forceinline__ void store(float* dst_s, uint32_t* dst_r, int k, int lane) const {
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const int p = m * 32 + lane;
      if (p < k) {
        dst_s[p] = s[m];
        dst_r[p] = r[m];
      }
    }
  }
  static constexpr int kPrefetch = 12;   // candidates per lane held in registers by select_from
  // REPLACE this list by the top-k of `nlists` lists of k entries each (list i at src + i*stride).
  // Selection instead of insertion: all candidates (<= 32*kPrefetch) are pulled into registers with independent loads,
  // then k rounds of "every lane offers its best remaining candidate, two warp reductions (redux.sync max on the
  // orderable score key, min on the row among the ties) pick the winner".  A round is ~100 instructions with a short
  // dependency chain; the insertion form costs ~8 dependent shuffles per accepted candidate and accepts ~k*ln(n/k) of
  // them: on the tail of a single query (per-CTA merge, last-CTA merge of 296 lists, peer exchange) that was ~16 us of
  // a 216 us request (tools/bench_group.py, k = 10 vs k = 1).  Falls back to init + merge_from for larger inputs.
  __device__ __forceinline__ void select_from(const volatile float* src_s, const volatile uint32_t* src_r,
                                              int nlists, int stride, int k, int lane) {
    const int total = nlists * k;
    if (M != 1 || total > 32 * kPrefetch) {
      init();
      merge_from(src_s, src_r, nlists, stride, k, lane);
      return;
    }
    const int rounds = (total + 31) >> 5;
    float ps[kPrefetch];
    uint32_t pr[kPrefetch];
#pragma unroll
    for (int i = 0; i < kPrefetch; ++i) {
      const int c = i * 32 + lane;
      ps[i] = VS_NEG_INF;
      pr[i] = kEmptyRow;
      if (i < rounds && c < total) {
        const int li = c / k, e = c - li * k;
        ps[i] = src_s[li * stride + e];
        pr[i] = src_r[li * stride + e];
      }
    }
    float out_s = VS_NEG_INF;
    uint32_t out_r = kEmptyRow;
    for (int j = 0; j < k; ++j) {
      // this lane's best remaining candidate
      float bs = ps[0];
      uint32_t br = pr[0];
#pragma unroll
      for (int i = 1; i < kPrefetch; ++i)
        if (i < rounds && better(ps[i], pr[i], bs, br)) {
          bs = ps[i];
          br = pr[i];
        }
      const uint32_t key = br == kEmptyRow ? 0u : score_key(bs);   // empty slots lose against everything
      const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
      const uint32_t rmin = __reduce_min_sync(0xffffffffu, key == kmax ? br : kEmptyRow);
      const unsigned who = __ballot_sync(0xffffffffu, key == kmax && br == rmin);
      const int src_lane = __ffs(who) - 1;
      const float ws = __shfl_sync(0xffffffffu, bs, src_lane);
      if (lane == j) {
        out_s = kmax == 0u ? VS_NEG_INF : ws;
        out_r = kmax == 0u ? kEmptyRow : rmin;
      }
      const bool mine = lane == src_lane;   // retire the winner (rows are unique among the candidates)
#pragma unroll
      for (int i = 0; i < kPrefetch; ++i) {
        const bool hit = mine && pr[i] == br;
        ps[i] = hit ? VS_NEG_INF : ps[i];
        pr[i] = hit ? kEmptyRow : pr[i];
      }
    }
    s[0] = out_s;
    r[0] = out_r;
    refresh_threshold(k - 1);
  }

  // merge `nlists` sorted lists of k entries each (list i at src + i*stride) into this list.
  // Lane-parallel prefilter against the threshold, then serial insertion of the survivors.
  __device__ __forceinline__ void merge_from(const volatile float* src_s, const volatile uint32_t* src_r,
                                             int nlists, int stride, int k, int lane) {
    const int total = nlists * k;
    for (int base = 0; base < total; base += 32) {
      const int c = base + lane;
      float cs = VS_NEG_INF;
      uint32_t cr = kEmptyRow;
      if (c < total) {
        const int li = c / k, e = c - li * k;
        cs = src_s[li * stride + e];
        cr = src_r[li * stride + e];
      }
      unsigned live = __ballot_sync(0xffffffffu, cr != kEmptyRow && better(cs, cr, thr_s, thr_r));
      while (live) {
        const int src_lane = __ffs(live) - 1;
        live &= live - 1;
        const float ns = __shfl_sync(0xffffffffu, cs, src_lane);
        const uint32_t nr = __shfl_sync(0xffffffffu, cr, src_lane);
        if (accepts(ns, nr)) insert(ns, nr, k - 1, lane);
      }
    }
  }
};

}  // namespace vs
