// exchange.cuh -- candidate exchange between the row shards over NVLink peer memory
// ("each GPU does its local top-k ... k-way merge produces the global
// result"; here the all-gather is replaced by P2P stores issued from inside the query kernel).
//
// Every rank owns one exchange buffer that all G ranks have mapped.  For one query slot a warp
//   1. push   : stores its k (score, global row) candidates into list [rank] of the slot in
//               EVERY rank's buffer (plain st.global on peer-mapped addresses -> NVLink writes),
//   2. signal : __threadfence_system, then st.release.sys of `epoch` into flag [slot][rank] of
//               every rank,
//   3. wait   : ld.acquire.sys spin on the G LOCAL flags of the slot (bounded),
//   4. merge  : merges the G lists that now sit in LOCAL memory into its register top-k list.
// A wait that times out (a peer never arrived within ~3 s) NEVER merges: the half-buffer may still
// hold lists of exchange e-2, so the slot's result is written as empty (-inf, -1) and the sticky
// error word -- host-mapped pinned memory, so every host entry point can test it without a CUDA
// call -- is raised; the host APIs turn it into VS_ERR_EXCHANGE.
// Steps 3+4 may be deferred: a batch of independent queries pushes into distinct slots of ONE epoch
// (push_only) and a single collect kernel waits for + merges all of them, so the ranks meet once
// per batch instead of once per query (per-query rendezvous makes every query as slow as the
// momentarily slowest GPU).
// The buffer has two halves selected by epoch parity.  A rank can be at most one exchange ahead
// of any peer: completing exchange e needs every peer's flag of e, and a peer raises it only after
// its own exchange e-1 finished (stream order / griddepcontrol.wait), so when half (e & 1) is
// overwritten by exchange e+2 every rank has already merged exchange e out of it.
// Half layout: flags u32 [Bmax][8] | scores f32 [G][Bmax][kmax] | rows u32 [G][Bmax][kmax].
#pragma once
#include "common.cuh"

namespace vs {

constexpr int kMaxPeers = 8;

struct XchgParams {
  unsigned char* peers[kMaxPeers];   // peer-mapped base pointers; peers[rank] is the local buffer
  unsigned int* err;                 // sticky error word in host-mapped pinned memory: 1 = a wait timed out
  int G;                             // 0 = no exchange
  int rank;
  int Bmax, kmax;
  int slot0;                         // slot of query 0 of this launch
  uint32_t epoch;
  int push_only;                     // 1: steps 1+2 only; a later collect kernel does 3+4 for the whole epoch
};

struct XchgLayout {
  size_t half_bytes, scores_off, rows_off;
};
__host__ __device__ inline XchgLayout xchg_layout(int Bmax, int kmax, int G) {
  XchgLayout L;
  size_t off = (size_t)Bmax * kMaxPeers * sizeof(uint32_t);
  off = (off + 255) & ~(size_t)255;
  L.scores_off = off;
  off += (size_t)G * Bmax * kmax * sizeof(float);
  off = (off + 255) & ~(size_t)255;
  L.rows_off = off;
  off += (size_t)G * Bmax * kmax * sizeof(uint32_t);
  L.half_bytes = (off + 255) & ~(size_t)255;
  return L;
}

// steps 1 + 2 (warp-collective).  `top` holds GLOBAL rows.
template <int M>
__device__ __forceinline__ void xchg_push(const XchgParams& x, const WarpTopK<M>& top, int slot, int k, int lane) {
  const XchgLayout L = xchg_layout(x.Bmax, x.kmax, x.G);
  const size_t half = (size_t)(x.epoch & 1u) * L.half_bytes;
  const size_t list = ((size_t)x.rank * x.Bmax + slot) * x.kmax;
  for (int g = 0; g < x.G; ++g) {
    unsigned char* base = x.peers[g] + half;
    top.store(reinterpret_cast<float*>(base + L.scores_off) + list,
              reinterpret_cast<uint32_t*>(base + L.rows_off) + list, k, lane);
  }
  __threadfence_system();
  __syncwarp();
  if (lane < x.G) {
    uint32_t* flag = reinterpret_cast<uint32_t*>(x.peers[lane] + half) + (size_t)slot * kMaxPeers + x.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(x.epoch) : "memory");
  }
}

// steps 3 + 4 (warp-collective): `top` is re-initialised and receives the global top-k, or stays
// empty when a peer never arrived (see the header comment).
// sm_keys: optional shared-memory scratch of sm_entries 64-bit slots owned by the calling warp: the G lists are copied
// there (packed) and selected with per-lane cursors (WarpTopK::select_sorted_smem) instead of the register scan.
template <int M>
__device__ __forceinline__ void xchg_wait_merge(const XchgParams& x, WarpTopK<M>& top, int slot, int k, int lane,
                                                uint64_t* sm_keys = nullptr, int sm_entries = 0) {
  const XchgLayout L = xchg_layout(x.Bmax, x.kmax, x.G);
  const size_t half = (size_t)(x.epoch & 1u) * L.half_bytes;
  const unsigned char* mine = x.peers[x.rank] + half;
  bool timed_out = false;
  if (lane < x.G) {
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(mine) + (size_t)slot * kMaxPeers + lane;
    const long long t0 = clock64();
    uint32_t v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if (v == x.epoch) break;
      if (clock64() - t0 > 6000000000LL) {   // ~3 s: a peer never arrived
        timed_out = true;
        break;
      }
      __nanosleep(64);
    }
  }
  top.init();
  if (__any_sync(0xffffffffu, timed_out)) {
    if (lane == 0) {
      *reinterpret_cast<volatile unsigned int*>(x.err) = 1u;
      __threadfence_system();
    }
    return;
  }
  const volatile float* ls = reinterpret_cast<const volatile float*>(mine + L.scores_off) + (size_t)slot * x.kmax;
  const volatile uint32_t* lr = reinterpret_cast<const volatile uint32_t*>(mine + L.rows_off) + (size_t)slot * x.kmax;
  if constexpr (M == 1) {
    if (sm_keys != nullptr && x.G * k <= sm_entries) {
      __syncwarp();
      stage_lists_as_keys(ls, lr, x.G, (size_t)x.Bmax * x.kmax, k, sm_keys, lane);
      __syncwarp();
      static_assert(kMaxPeers <= 8, "the merge tree below takes at most 8 lists");
      if (k <= 16) top.template merge_sorted_bitonic<8>(sm_keys, x.G, k, k, lane);
      else top.select_sorted_smem(sm_keys, x.G, k, k, lane);
      return;
    }
  }
  top.select_from(ls, lr, x.G, x.Bmax * x.kmax, k, lane);
}

}  // namespace vs
