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

  // ---- selection over SORTED lists held in shared memory as packed keys -------------------------------------------
  // key = score_key : ~row (64 bit, larger ranks first, 0 = empty slot).  Used on the tail of a single query, where one
  // warp works alone and the dependent-instruction chain IS the latency (tools/scan_stamps.py: the register-scan form,
  // select_from, cost 3.4 us per call at k = 10 -- three calls per query).
  static __device__ __forceinline__ uint64_t pack_key(float sc, uint32_t row) {
    return row == kEmptyRow ? 0ull : ((uint64_t)score_key(sc) << 32) | (uint32_t)~row;
  }
  // write the first k entries as packed keys
  __device__ __forceinline__ void store_keys(uint64_t* dst, int k, int lane) const {
    static_assert(M == 1, "packed-key lists are a k <= 32 form");
    if (lane < k) dst[lane] = pack_key(s[0], r[0]);
  }
  // REPLACE this list by the top-k of `nlists` <= 64 sorted key lists (list i at keys + i*stride, k entries each, empties
  // last).  Lane l owns lists l and l+32 and keeps only their current head and the entry after it in registers.  A
  // round: max of the lane's two heads, ONE redux.sync.max on the score word, a ballot to find the owner (a second redux
  // on the row word only when two lanes tie on the score), then the owner moves to its prefetched next entry and issues
  // the load of the one after.  The winner reaches output lane j by a shuffle that nothing waits for.
  static constexpr int kMaxSortedLists = 64;
  __device__ __forceinline__ void select_sorted_smem(const uint64_t* keys, int nlists, int stride, int k, int lane) {
    static_assert(M == 1, "packed-key lists are a k <= 32 form");
    const uint64_t* l0 = keys + (size_t)lane * stride;
    const uint64_t* l1 = keys + (size_t)(lane + 32) * stride;
    const bool has0 = lane < nlists, has1 = lane + 32 < nlists;
    int c0 = 0, c1 = 0;
    uint64_t h0 = has0 ? l0[0] : 0ull, h1 = has1 ? l1[0] : 0ull;
    uint64_t n0 = (has0 && k > 1) ? l0[1] : 0ull, n1 = (has1 && k > 1) ? l1[1] : 0ull;
    uint32_t out_hi = 0u, out_lo = 0u;
    for (int j = 0; j < k; ++j) {
      const uint64_t m = h0 > h1 ? h0 : h1;
      const uint32_t mhi = (uint32_t)(m >> 32), mlo = (uint32_t)m;
      const uint32_t hi = __reduce_max_sync(0xffffffffu, mhi);
      if (hi == 0u) break;   // every list exhausted (warp-uniform); the remaining output slots stay empty
      bool cand = mhi == hi;
      unsigned who = __ballot_sync(0xffffffffu, cand);
      if (who & (who - 1u)) {   // equal scores in different lanes: the lower row (larger ~row) goes first
        const uint32_t lo = __reduce_max_sync(0xffffffffu, cand ? mlo : 0u);
        cand = cand && mlo == lo;
        who = __ballot_sync(0xffffffffu, cand);   // rows are unique among the candidates: exactly one lane is left
      }
      const int src = __ffs(who) - 1;
      const uint32_t wlo = __shfl_sync(0xffffffffu, mlo, src);
      if (lane == j) {
        out_hi = hi;
        out_lo = wlo;
      }
      if (lane == src) {
        if (h0 == m) {
          h0 = n0;
          ++c0;
          n0 = c0 + 1 < k ? l0[c0 + 1] : 0ull;
        } else {
          h1 = n1;
          ++c1;
          n1 = c1 + 1 < k ? l1[c1 + 1] : 0ull;
        }
      }
    }
    s[0] = out_hi == 0u ? VS_NEG_INF : key_score(out_hi);
    r[0] = out_hi == 0u ? kEmptyRow : ~out_lo;
    refresh_threshold(k - 1);
  }

  // REPLACE this list by the top-k (k <= 16) of `nlists` <= NL sorted key lists in shared memory -- no serial rounds at
  // all: a bitonic merge TREE held in registers.  A register slot holds two lists: lanes 0-15 list 2s in order, lanes
  // 16-31 list 2s+1 reversed = one bitonic sequence; five butterfly compare-exchange stages (64-bit shuffles) sort it, its
  // top 16 sit in lanes 0-15; two sorted slots are paired the same way (one shuffle reverses the second into lanes 16-31)
  // until one is left.  NL = 8: 3 levels x 5 stages on 4 + 2 + 1 slots, ~370 instructions, ~0.4 us, against ~2.3 us for
  // the k = 10 rounds of select_sorted_smem (profiles/r02_group_latency.md).  Empty slots (key 0) sort last by themselves.
  template <int NL>
  __device__ __forceinline__ void merge_sorted_bitonic(const uint64_t* keys, int nlists, int stride, int k, int lane) {
    static_assert(M == 1 && (NL == 8 || NL == 16), "8 or 16 lists of at most 16 entries");
    constexpr int NS = NL / 2;
    constexpr int LEVELS = NL == 8 ? 3 : 4;
    uint64_t v[NS];
    const int half = lane >> 4;
    const int e = half ? 31 - lane : lane;   // entry of the list this lane holds (0..15)
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
      const int li = 2 * sl + half;
      v[sl] = (li < nlists && e < k) ? keys[(size_t)li * stride + e] : 0ull;
    }
#pragma unroll
    for (int lvl = 0; lvl < LEVELS; ++lvl) {
      const int width = NS >> lvl;           // live slots on this level (compile-time after unrolling)
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        const bool up = (lane & d) == 0;     // this lane keeps the larger key of the pair
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
          if (sl < width) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, v[sl], d);
            // keys are unique except for the empty key 0, where exchanging is a no-op: take the partner's key iff
            // (it is larger) == (this lane keeps the larger one)
            if ((other > v[sl]) == up) v[sl] = other;
          }
        }
      }
      if (width > 1) {
#pragma unroll
        for (int t = 0; t < NS / 2; ++t) {
          if (t < width / 2) {
            const uint64_t second = __shfl_sync(0xffffffffu, v[2 * t + 1], 31 - lane);   // its top 16, reversed, for lanes 16-31
            v[t] = half ? second : v[2 * t];
          }
        }
      }
    }
    const uint64_t w = lane < k ? v[0] : 0ull;
    const uint32_t whi = (uint32_t)(w >> 32);
    s[0] = whi == 0u ? VS_NEG_INF : key_score(whi);
    r[0] = whi == 0u ? kEmptyRow : ~(uint32_t)w;
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

// Copy `nl` sorted lists of k (score, row) entries from GLOBAL memory (list i at gs/gr + i*gstride) into shared memory as
// packed keys, dst[i*k + e].  The loads of a chunk (32 lanes x kStageUnroll entries) are all issued before the first
// store: a rolled loop would pay one L2 round trip per 32 entries (0.4 us each on the tail of a query).
constexpr int kStageUnroll = 8;
__device__ __forceinline__ void stage_lists_as_keys(const volatile float* gs, const volatile uint32_t* gr, int nl,
                                                    size_t gstride, int k, uint64_t* dst, int lane) {
  const int total = nl * k;
  const uint32_t inv_k = ((1u << 20) + (uint32_t)k - 1u) / (uint32_t)k;   // c / k == (c * inv_k) >> 20 for c < 32768, k <= 32
  for (int c0 = 0; c0 < total; c0 += 32 * kStageUnroll) {
    float vs[kStageUnroll];
    uint32_t vr[kStageUnroll];
#pragma unroll
    for (int u = 0; u < kStageUnroll; ++u) {
      const int c = c0 + u * 32 + lane;
      vs[u] = VS_NEG_INF;
      vr[u] = kEmptyRow;
      if (c < total) {
        const int li = (int)(((uint32_t)c * inv_k) >> 20), e = c - li * k;
        vs[u] = gs[(size_t)li * gstride + e];
        vr[u] = gr[(size_t)li * gstride + e];
      }
    }
#pragma unroll
    for (int u = 0; u < kStageUnroll; ++u) {
      const int c = c0 + u * 32 + lane;
      if (c < total) dst[c] = WarpTopK<1>::pack_key(vs[u], vr[u]);
    }
  }
}

}  // namespace vs
