// gemm_tc.cu -- K2/K3/K4: batched cosine scoring on the 5th-gen tensor cores (sm_100a).
//
//   K2  batched top-k      : batched form of Collection.query (backend/app/main.py:761-765)
//   K3  filter sweep       : F prompt embeddings x N rows, thresholded to a bit mask
//   K4  all-pairs dedup    : rows x rows, pairs with cos >= tau
//
// One warp-specialised kernel, three epilogues.  D[m, n] = sum_k A[m, k] * B[n, k] with
//   A = a block of 128 "query-side" vectors (queries / prompts / a block of corpus rows), bf16,
//       128B swizzle; RESIDENT in shared memory for the whole work item when that leaves a deep
//       enough ring (dim <= 512 in pair mode), else its 16 KB k-block travels with every stage,
//   B = corpus rows, streamed from HBM in tiles of BN rows x 64 k-elements by TMA into a ring of
//       stages.  The C CTAs of a thread-block CLUSTER own C different A blocks and share ONE
//       corpus stream: each CTA fetches 1/C of every B tile and TMA-multicasts it to the CTAs that
//       need it, so a corpus byte crosses L2->SM once per cluster instead of once per A block,
//   D = fp32 accumulators in TMEM: lane = query, column = corpus row, 512/BN buffers so the MMA
//       of tile t+1 overlaps the epilogue of tile t.
// Issue mode (template CG): CG = 2 (default) pairs the CTAs (2p, 2p+1) of the cluster with
// tcgen05 cta_group::2 -- the even CTA issues one M = 256 MMA for both A blocks and each CTA holds
// only its half of the B tile; CG = 1 issues M = 128 MMAs per CTA.
// Roles: warp 0 = TMA producer (rows, A k-blocks, the tile's inverse norms and group bounds),
// warp 1 = tcgen05.mma issuer (one elected thread), warps 2..9 = epilogue (tcgen05.ld 32x32b: each
// thread owns ONE query and sees that query's scores against 32 consecutive rows per load -> a
// private register top-k list, no cross-thread traffic).
// Epilogue arithmetic: score = acc * inv_norm[row] (queries are L2-normalised then rounded to
// bf16 by prep_queries_kernel; the oracle does the same); a fast reject on the raw accumulators
// against a per-query bound shared by all lists (pool of k class maxima) keeps the exact path rare;
// filter bits are only fetched for rows that would enter a list, nothing but the final candidates
// is written.
#include <cuda.h>

#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace vs {

constexpr int kTcM = 128;          // A rows per block (UMMA M)
constexpr int kTcKB = 64;          // k elements per stage (= 128 B = one swizzle span)
constexpr int kTcThreads = 320;      // TMA warp + MMA warp + 8 epilogue warps
constexpr int kTcMaxStages = 8;
constexpr int kInvRing = 4;          // tiles of inverse norms staged ahead of the epilogue
constexpr int kTcBarrierBytes = 512; // mbarriers + TMEM address holder
constexpr int kTcSmemMax = 232448; // 227 KB
constexpr int kTcMaxBatch = 4096;  // queries / prompts per launch (the partial-list workspace is sized for it)
constexpr int64_t kTopkPairRowsBelow = 7000000;   // shards shorter than this run K2 as CTA pairs + A groups
constexpr int kFilterClusterCap = 8;
constexpr int kMaxDevices = 64;

enum { kModeTopK = 0, kModeFilter = 1, kModeDedup = 2 };

struct TcParams {
  const float* inv_norm;
  const uint64_t* mask;
  uint64_t req[kMaskWords];
  int use_mask;
  int64_t row_base;
  int64_t row_stride;
  uint32_t n_rows;
  uint32_t n_tiles;        // ceil(n_rows / BN)
  int kb_count;            // ceil(dim / 64)
  int stages;
  int n_slices;            // corpus slices (top-k / filter); one cluster per slice per pass
  int tiles_per_slice;
  int n_items;             // work items per cluster sequence
  int a_stream;            // 1: the A block is NOT resident; its k-block travels with every B stage (wide rows)
  int prefetch;            // B stages to prefetch into L2 ahead of the smem ring
  int n_agroups;           // A-block groups of C blocks each: work item w = (slice w / n_agroups, A group w % n_agroups)
  const float* gmin;       // [ceil(n_rows/32)] (1 - 2^-20) * min row norm of each 32-row group (+inf if empty)
  // top-k
  uint32_t* gbound;        // [Bp][pool stride] orderable keys: per query k class maxima (row % k) over everything ANY
                           // slice / epilogue half has inserted; their minimum is a lower bound on the k-th score
  int k_real;              // slots of the pool in use (= k of this round)
  // rounds for 32 < k <= 128 (see launch_tensor_topk): only rows ranking strictly AFTER (ub_s[q], ub_r[q]) compete
  const float* ub_s;       // [B] stride ub_stride, or nullptr (first round)
  const int64_t* ub_r;
  int ub_stride;
  float* part_s;           // [n_slices][2][Bp][KL]  (2 = the two epilogue halves)
  int64_t* part_r;
  int Bp;                  // C * 128
  // filter
  uint32_t* out_bits;
  int64_t words_per_filter;
  float tau;
  int F;                   // valid A rows in this launch (rows >= F are padding)
  // dedup
  int64_t a_row_lo;        // first corpus row of A block 0 (128-aligned)
  int64_t a_row_min;       // caller's row_lo: rows below it are not reported
  int64_t a_row_hi;
  int64_t cap;
  int64_t* out_i;
  int64_t* out_j;
  float* out_score;
  unsigned long long* out_count;
};

// ------------------------------------------------------------------------------------------
// PTX wrappers: clusters, TMA tensor loads, tcgen05 (alloc / mma / commit / ld / fences)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "l"(policy)
      : "memory");
}
// same box delivered to the same CTA-relative smem offset (and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                               uint16_t mask, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%4, %5}], [%2], %3, %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0),
      "r"(c1), "l"(policy)
      : "memory");
}
// pull a box into L2 only (no smem, no barrier): hides HBM latency behind the shallow smem ring
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* holder_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// ---- cta_group::2 (CTA pair = cluster ranks 2p, 2p+1; the even CTA is the leader and issues the MMAs) ----
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the pair bit of a shared::cluster address -> the leader's copy
__device__ __forceinline__ void tmem_alloc2(uint32_t* holder_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA loads issued by either CTA of a pair; the transaction bytes are counted on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                                uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
      "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc_cg2(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                                   uint16_t mask, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      ".L2::cache_hint [%0], [%1, {%4, %5}], [%2], %3, %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "h"(mask),
      "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows per CTA] * B[N/2 rows per CTA]^T: M = 256 across the pair
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc_cg2(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// arrive on the same barrier of another CTA of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  // default semantics (release at CTA scope), as for the local arrives: the only thing the leader must
  // observe is that this warp's tcgen05.ld's have retired, which tcgen05.wait::ld + the
  // before_thread_sync fence already order; a cluster-scope release costs a MEMBAR per tile and warp
#ifdef VS_REMOTE_ARRIVE_RELEASE   // A/B builds only
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
#else
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// Waits of the producer thread (stage release) and of the epilogue warps (accumulator ready) are LONG (a whole tile): a
// suspend-time hint lets the hardware park the warp until the phase completes instead of re-issuing try_wait tens of times
// per microsecond (ncu, round 1 code: 12.7 M producer retries + 7.3 M epilogue retries per 2.4 ms sweep), which costs issue
// slots and, under the power cap, clock.  VS_TC_WAIT_HINT_NS = 0 compiles the plain form.
#ifndef VS_TC_WAIT_HINT_NS
#define VS_TC_WAIT_HINT_NS 0
#endif
__device__ __forceinline__ void mbar_wait_long(uint64_t* bar, uint32_t parity) {
#if VS_TC_WAIT_HINT_NS > 0
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)VS_TC_WAIT_HINT_NS)
        : "memory");
  } while (!ok);
#else
  mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> f32, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all tcgen05 ops issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// ... and on the same barrier of every CTA of the cluster in `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// 32 lanes x 32 consecutive columns: thread i of the warp receives lane (base+i), columns c..c+31
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory matrix descriptor (sm_100 "version 1"):
//   start address >> 4 | LBO (unused for swizzled K-major) = 1 | SBO = 1024 B (8 rows x 128 B) |
//   version = 1 | layout = SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// per-thread sorted list (descending), KL entries in registers
// ------------------------------------------------------------------------------------------
template <int KL>
struct ThreadTopK {
  float s[KL];
  uint32_t r[KL];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < KL; ++i) {
      s[i] = VS_NEG_INF;
      r[i] = kEmptyRow;
    }
  }
  __device__ __forceinline__ float threshold() const { return s[KL - 1]; }
  // rows arrive in increasing order per thread, so a strict compare keeps (score desc, row asc)
  __device__ __forceinline__ void insert(float ns, uint32_t nr) {
    s[KL - 1] = ns;
    r[KL - 1] = nr;
#pragma unroll
    for (int j = KL - 1; j > 0; --j) {
      const bool sw = s[j] > s[j - 1];
      const float a = s[j], b = s[j - 1];
      const uint32_t ra = r[j], rb = r[j - 1];
      s[j - 1] = sw ? a : b;
      s[j] = sw ? b : a;
      r[j - 1] = sw ? ra : rb;
      r[j] = sw ? rb : ra;
    }
  }
};

// ---- per-query shared pool of k class maxima (orderable keys) -------------------------------------
// Every epilogue thread keeps its own list for (slice, half); alone, each list only knows the k-th
// best of ITS rows, so all ~30 lists of a query warm up separately and the exact slow path runs for
// a third of the 32-row groups.  The pool has k slots; slot c holds the best score inserted so far, by
// ANY list, among corpus rows with row % k == c.  The k slots are scores of k DISTINCT rows, so their
// minimum is a lower bound on the query's final k-th best score -- and it reflects every row processed
// so far by every slice, not one list's share.  Updating it is one fire-and-forget atomicMax (RED):
// no read-modify-write loop, no latency on the insert path; reading it is 3 vector loads per tile.
__host__ __device__ constexpr int pool_stride(int KL) { return KL <= 12 ? 12 : 32; }   // 16-byte aligned rows
template <int KLP>
__device__ __forceinline__ uint32_t pool_min(const uint32_t* pool, int k, int& idx) {
  uint32_t mn = 0xFFFFFFFFu;
  idx = 0;
#pragma unroll
  for (int j4 = 0; j4 < KLP / 4; ++j4) {
    uint32_t a, b, c, d;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(pool + 4 * j4));
    if (4 * j4 < k && a < mn) { mn = a; idx = 4 * j4; }
    if (4 * j4 + 1 < k && b < mn) { mn = b; idx = 4 * j4 + 1; }
    if (4 * j4 + 2 < k && c < mn) { mn = c; idx = 4 * j4 + 2; }
    if (4 * j4 + 3 < k && d < mn) { mn = d; idx = 4 * j4 + 3; }
  }
  return mn;
}
__device__ __forceinline__ bool tc_mask_ok(const uint64_t* mask, uint32_t row, const uint64_t* req) {
  const uint64_t* m = mask + (size_t)row * kMaskWords;
  bool ok = true;
#pragma unroll
  for (int w = 0; w < kMaskWords; ++w) ok = ok && ((__ldg(m + w) & req[w]) == req[w]);
  return ok;
}

// ---- filter-sweep epilogue helpers: exact `acc * inv_norm >= tau` for 32 columns in ~2.5 instructions per column ----
// (packed f32x2 multiply = the same round-to-nearest product as the scalar FMUL the oracle's arithmetic implies, one
//  FSETP and one predicated OR with an immediate bit; the straightforward C compiles to ~5.)
__device__ __forceinline__ void mul_f32x2(uint32_t a0, uint32_t a1, float b0, float b1, float& d0, float& d1) {
  asm("{\n\t.reg .b64 a, b, d;\n\t"
      "mov.b64 a, {%2, %3};\n\t"
      "mov.b64 b, {%4, %5};\n\t"
      "mul.rn.f32x2 d, a, b;\n\t"
      "mov.b64 {%0, %1}, d;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "r"(a0), "r"(a1), "f"(b0), "f"(b1));
}
template <int J>
__device__ __forceinline__ void or_bit_if_ge(uint32_t& bits, float s, float tau) {
  asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(bits) : "f"(s), "f"(tau), "n"(1u << J));
}
template <int J>
__device__ __forceinline__ void threshold_bits(uint32_t& bits, const uint32_t (&acc)[32], const float (&inv)[32], float tau) {
  float s0, s1;
  mul_f32x2(acc[J], acc[J + 1], inv[J], inv[J + 1], s0, s1);
  or_bit_if_ge<J>(bits, s0, tau);
  or_bit_if_ge<J + 1>(bits, s1, tau);
  if constexpr (J + 2 < 32) threshold_bits<J + 2>(bits, acc, inv, tau);
}

// ------------------------------------------------------------------------------------------
// the kernel.  C = cluster size (CTAs sharing one corpus stream, one A block each)
// ------------------------------------------------------------------------------------------
// CG = 2: the CTAs (2p, 2p+1) of the cluster form tcgen05 CTA PAIRS.  The even CTA issues ONE
// M = 256 MMA for both A blocks; each CTA keeps only HALF of every B tile (its BN/2 rows) in smem, so
// per MMA a CTA reads 4 KB of A + 4 KB of B from smem instead of 4 + 8, and receives 16 + 16 KB per
// stage instead of 16 + 32.  Accumulators stay per CTA (128 lanes x BN columns), so the epilogue is
// unchanged.  CG = 2 needs an even cluster size; A is streamed per stage or (when it fits) resident.
template <int MODE, int BN, int KL, int C, int CG = 1>
__global__ void __launch_bounds__(kTcThreads, 1)
    tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  static_assert(CG == 1 || (C % 2 == 0), "CTA pairs need an even cluster");
  constexpr int ACC = 512 / BN;              // accumulator buffers in TMEM
  constexpr uint32_t kStageBytes = BN * 128 / CG; // B rows held per CTA per stage x 64 bf16
  constexpr uint32_t kPartRows = BN / C;     // rows of each B tile this CTA fetches (and multicasts)
  constexpr uint32_t kABlockBytes = kTcM * 128;
  constexpr uint32_t kIdesc = make_idesc(kTcM * CG, BN);
  constexpr uint16_t kMask = (uint16_t)((1u << C) - 1u);
  constexpr uint16_t kEvenMask = (uint16_t)(0x5555u & kMask);   // pair leaders
  constexpr uint32_t kHalfC = C >= 2 ? C / 2 : 1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 128B swizzle: 1024-B aligned
  // resident A: [A: kb_count blocks][stages x B tile]; streamed A: [stages x (B tile | A k-block)]
  uint8_t* sA = smem;
  uint8_t* sB = smem + (p.a_stream ? 0 : (size_t)p.kb_count * kABlockBytes);
  const uint32_t stage_stride = kStageBytes + (p.a_stream ? kABlockBytes : 0u);
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + (size_t)p.stages * stage_stride);
  uint64_t* empty = full + kTcMaxStages;
  uint64_t* tfull = empty + kTcMaxStages;
  uint64_t* tempty = tfull + 4;
  uint64_t* a_full = tempty + 4;
  uint64_t* a_empty = a_full + 1;
  uint64_t* ifull = a_empty + 1;             // inverse norms of a tile staged in smem: ring of kInvRing tiles
  uint64_t* iempty = ifull + kInvRing;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(iempty + kInvRing);
  // [kInvRing][BN inverse norms | BN/32 group min-norm bounds]
  float* s_invt = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + kTcBarrierBytes);
  constexpr uint32_t kInvStride = BN + BN / 32;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = C > 1 ? cluster_ctarank() : 0u;
  const bool leader = CG == 1 || (rank & 1u) == 0u;
  const int cluster_id = blockIdx.x / C;
  const int n_clusters = gridDim.x / C;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], C / CG);   // one tcgen05.commit arrival from every MMA issuer of the cluster
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8 * CG);   // one arrival per epilogue warp (of both CTAs of a pair)
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < kInvRing; ++i) {
      mbar_init(&ifull[i], 1);
      mbar_init(&iempty[i], 8);   // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc2(tmem_holder, 512); else tmem_alloc(tmem_holder, 512);
  }
  tc_fence_before();
  if (C > 1) cluster_sync_all(); else __syncthreads();   // barriers of every CTA initialised before remote arrivals
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // work item (one per cluster) -> this CTA's A block + the cluster's tile range
  auto item_range = [&](int w, int& ablock, uint32_t& t0, uint32_t& t1) {
    if (MODE == kModeDedup) {
      ablock = w * C + (int)rank;
      const uint32_t a_row0 = (uint32_t)p.a_row_lo + (uint32_t)(w * C) * kTcM;   // cluster's first row
      t0 = a_row0 / BN;       // only columns j > i can pair with row i
      t1 = p.n_tiles;
    } else {
      // the clusters (slice, 0 .. n_agroups-1) stream the same tiles at the same time: the slice comes from HBM
      // once and from L2 for the other A groups
      ablock = (w % p.n_agroups) * C + (int)rank;
      t0 = (uint32_t)(w / p.n_agroups) * p.tiles_per_slice;
      t1 = min(t0 + (uint32_t)p.tiles_per_slice, p.n_tiles);
    }
  };

  if (warp == 0) {
    // ================= TMA producer =========================================================
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();
      const uint64_t pol_keep = policy_evict_last();
      uint32_t it = 0, n_item = 0, ptile = 0;
      for (int w = cluster_id; w < p.n_items; w += n_clusters, ++n_item) {
        int ablock;
        uint32_t t0, t1;
        item_range(w, ablock, t0, t1);
        const int a_row = MODE == kModeDedup ? (int)(p.a_row_lo + (int64_t)ablock * kTcM) : ablock * kTcM;
        if (!p.a_stream) {
          // A block: resident for the whole item
          if (n_item > 0) mbar_wait(a_empty, (n_item - 1) & 1);
          if (CG == 2) {
            // pair mode: both CTAs' A blocks are counted on the LEADER's barrier (its MMA reads both)
            if (leader) mbar_expect_tx(a_full, 2u * (uint32_t)p.kb_count * kABlockBytes);
            for (int kb = 0; kb < p.kb_count; ++kb)
              tma_load_2d_cg2(sA + (size_t)kb * kABlockBytes, &tmA, kb * kTcKB, a_row, a_full, pol_keep);
          } else {
            mbar_expect_tx(a_full, (uint32_t)p.kb_count * kABlockBytes);
            for (int kb = 0; kb < p.kb_count; ++kb)
              tma_load_2d(sA + (size_t)kb * kABlockBytes, &tmA, kb * kTcKB, a_row, a_full, pol_keep);
          }
        }
        // L2 prefetch runs `prefetch` stages ahead of the smem ring (this CTA's part of each box)
        const uint32_t n_st = (t1 - t0) * (uint32_t)p.kb_count;
        auto prefetch_stage = [&](uint32_t idx) {
          if (idx < n_st) {
            const uint32_t pt = t0 + idx / (uint32_t)p.kb_count;
            const int pkb = (int)(idx % (uint32_t)p.kb_count);
            tma_prefetch_2d(&tmB, pkb * kTcKB, (int)(pt * BN + rank * kPartRows));
          }
        };
        for (uint32_t i = 0; i < (uint32_t)p.prefetch; ++i) prefetch_stage(i);
        uint32_t idx = 0;
        for (uint32_t t = t0; t < t1; ++t, ++ptile) {
          {
            // this tile's BN inverse norms -> smem (the epilogue's exact path reads them with LDS instead
            // of stalling on L2); the array is padded past n_rows, so a whole tile is always in bounds
            const uint32_t ib = ptile % kInvRing;
            if (ptile >= (uint32_t)kInvRing) mbar_wait_long(&iempty[ib], (ptile / kInvRing - 1) & 1);
            mbar_expect_tx(&ifull[ib], BN * 4 + (p.gmin ? BN / 32 * 4 : 0));
            bulk_g2s(s_invt + ib * kInvStride, p.inv_norm + (size_t)t * BN, BN * 4, &ifull[ib], pol_keep);
            if (p.gmin) bulk_g2s(s_invt + ib * kInvStride + BN, p.gmin + (size_t)t * (BN / 32), BN / 32 * 4, &ifull[ib], pol_keep);
          }
          for (int kb = 0; kb < p.kb_count; ++kb, ++it, ++idx) {
            const int st = it % p.stages;
            const uint32_t use = it / p.stages;
            if (p.prefetch > 0) prefetch_stage(idx + (uint32_t)p.prefetch);
            if (use > 0) mbar_wait_long(&empty[st], (use - 1) & 1);   // all C CTAs have consumed this stage
            const int row = (int)(t * BN + rank * kPartRows);
            if (CG == 2) {
              // pair mode: the LEADER's barrier counts everything both CTAs of the pair receive for this
              // stage (2 A k-blocks + 2 half B tiles).  This CTA's part of the B tile belongs to half
              // h = rank / (C/2) of the tile and goes to every CTA of parity h (leaders hold half 0).
              if (leader) mbar_expect_tx(&full[st], 2u * (kStageBytes + (p.a_stream ? kABlockBytes : 0u)));
              if (p.a_stream)
                tma_load_2d_cg2(sB + (size_t)st * stage_stride + kStageBytes, &tmA, kb * kTcKB, a_row, &full[st], pol_keep);
              const uint32_t h = rank / kHalfC;
              uint8_t* dst = sB + (size_t)st * stage_stride + (size_t)(rank % kHalfC) * kPartRows * 128;
              tma_load_2d_mc_cg2(dst, &tmB, kb * kTcKB, row, &full[st], (uint16_t)(kEvenMask << h), pol_stream);
            } else {
              // the whole B tile (C parts from C CTAs) + this CTA's own A k-block when A is streamed
              mbar_expect_tx(&full[st], kStageBytes + (p.a_stream ? kABlockBytes : 0u));
              if (p.a_stream)   // re-read per tile from L2 (the A block is 128 x dim bf16 <= 256 KB, L2 resident)
                tma_load_2d(sB + (size_t)st * stage_stride + kStageBytes, &tmA, kb * kTcKB, a_row, &full[st], pol_keep);
              uint8_t* dst = sB + (size_t)st * stage_stride + (size_t)rank * kPartRows * 128;
              if (C > 1)
                tma_load_2d_mc(dst, &tmB, kb * kTcKB, row, &full[st], kMask, pol_stream);
              else
                tma_load_2d(dst, &tmB, kb * kTcKB, row, &full[st], pol_stream);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (pair mode: the leader CTA only) ==============================
    if (lane == 0 && leader) {
      uint32_t it = 0, n_item = 0, tile_ctr = 0;
      for (int w = cluster_id; w < p.n_items; w += n_clusters, ++n_item) {
        int ablock;
        uint32_t t0, t1;
        item_range(w, ablock, t0, t1);
        if (!p.a_stream) {
          mbar_wait(a_full, n_item & 1);
          tc_fence_after();
        }
        for (uint32_t t = t0; t < t1; ++t, ++tile_ctr) {
          const uint32_t acc = tile_ctr % ACC;
          const uint32_t use = tile_ctr / ACC;
          if (use > 0) {
            if (CG == 2) mbar_wait_cluster(&tempty[acc], (use - 1) & 1); else mbar_wait(&tempty[acc], (use - 1) & 1);
          }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = 0; kb < p.kb_count; ++kb, ++it) {
            const int st = it % p.stages;
            mbar_wait(&full[st], (it / p.stages) & 1);
            tc_fence_after();
            const uint64_t da = make_smem_desc(smem_u32(p.a_stream ? sB + (size_t)st * stage_stride + kStageBytes
                                                                   : sA + (size_t)kb * kABlockBytes));
            const uint64_t db = make_smem_desc(smem_u32(sB + (size_t)st * stage_stride));
#pragma unroll
            for (int k = 0; k < kTcKB / 16; ++k) {  // UMMA_K = 16 bf16 = 32 B: advance start address by 2 (16-B units)
              if (CG == 2) umma_bf16_cg2(d_tmem, da + 2 * k, db + 2 * k, kIdesc, (kb | k) != 0);
              else umma_bf16(d_tmem, da + 2 * k, db + 2 * k, kIdesc, (kb | k) != 0);
            }
            // frees the stage (in every CTA of the cluster) when these MMAs have read it
            if (CG == 2) umma_commit_mc_cg2(&empty[st], kMask);
            else if (C > 1) umma_commit_mc(&empty[st], kMask);
            else umma_commit(&empty[st]);
          }
          // accumulator ready for the epilogue (pair mode: of both CTAs)
          if (CG == 2) umma_commit_mc_cg2(&tfull[acc], (uint16_t)(3u << rank)); else umma_commit(&tfull[acc]);
        }
        if (!p.a_stream) {                         // A block may be overwritten (pair mode: in both CTAs)
          if (CG == 2) umma_commit_mc_cg2(a_empty, (uint16_t)(3u << rank)); else umma_commit(a_empty);
        }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps (2..9): TMEM lane quarter = warp % 4 ==========================
    // Two warps per lane quarter (one per scheduler pair) split the 32-column groups of every tile
    // even/odd, so TMEM/global latencies of one overlap the list work of the other.  Each thread
    // owns ONE query and keeps its own list; the two lists of a query are merged with the slices'.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;              // 0: even groups, 1: odd groups
    const int m_local = quarter * 32 + lane;       // A row (query) owned by this thread
    ThreadTopK<KL> top;
    uint32_t tile_ctr = 0;
    for (int w = cluster_id; w < p.n_items; w += n_clusters) {
      int ablock;
      uint32_t t0, t1;
      item_range(w, ablock, t0, t1);
      if (MODE == kModeTopK) top.init();
      const uint32_t a_global = (MODE == kModeDedup ? (uint32_t)p.a_row_lo : 0u) + (uint32_t)ablock * kTcM + m_local;
      float inv_a = 1.f;
      bool a_ok = true;
      if (MODE == kModeDedup) {
        a_ok = a_global < (uint32_t)p.a_row_hi && a_global >= (uint32_t)p.a_row_min;
        inv_a = a_ok ? __ldg(p.inv_norm + a_global) : 0.f;
      }
      // top-k: a lower bound on this query's global k-th score, shared by all threads (slices,
      // halves) that work on the same query through gbound[] (atomicMax on orderable keys).  A row
      // scoring strictly below it cannot be in the global top-k.
      uint32_t* gb_ptr = nullptr;
      constexpr int KLP = pool_stride(KL);
      bool has_ub = false;
      float ub_s = 0.f;
      int64_t ub_r = 0;
      if (MODE == kModeTopK) {
        gb_ptr = p.gbound + ((size_t)ablock * kTcM + m_local) * KLP;
        if (p.ub_s != nullptr && ablock * kTcM + m_local < p.F) {
          has_ub = true;
          ub_s = __ldg(p.ub_s + (size_t)(ablock * kTcM + m_local) * p.ub_stride);
          ub_r = __ldg(p.ub_r + (size_t)(ablock * kTcM + m_local) * p.ub_stride);
          if (ub_r < 0) ub_s = VS_NEG_INF;   // the previous round ran out of rows: nothing is left for this query
        }
      }
      for (uint32_t t = t0; t < t1; ++t, ++tile_ctr) {
        const uint32_t acc = tile_ctr % ACC;
        float gb = VS_NEG_INF;
        if (MODE == kModeTopK) {
          int unused;
          gb = key_score(pool_min<KLP>(gb_ptr, p.k_real, unused));
        }
        mbar_wait_long(&tfull[acc], (tile_ctr / ACC) & 1);
        tc_fence_after();
        const uint32_t ib = tile_ctr % kInvRing;
        mbar_wait(&ifull[ib], (tile_ctr / kInvRing) & 1);
        const uint32_t row0 = t * BN;
#ifdef VS_TC_NOEPI   // profiling build: the epilogue only hands the accumulator back (mainloop ceiling)
        constexpr int n_groups = 0;
#else
        constexpr int n_groups = BN / 32;
#endif
        if constexpr (MODE == kModeFilter) {
          // Filter sweep: the two warps of a lane quarter take the LOWER / UPPER half of the tile's 32-column groups, so
          // every thread ends up with BN/64 CONSECUTIVE bit words of its prompt's row and writes them with one vector
          // store (16 B for BN = 256) instead of BN/64 scattered 4-byte stores: 4x fewer L2 write transactions.
          // The groups are read from TMEM TWO at a time (one wait for both loads), and the accumulator buffer is handed
          // back to the MMA issuer as soon as the LAST load has landed in registers -- before the arithmetic on it --
          // so the next-but-one tile's MMAs are not held up by this tile's thresholding and stores.
          constexpr int kWords = BN / 64;
          static_assert(kWords % 2 == 0, "two groups per TMEM wait");
          uint32_t words[kWords];
          bool released = false;
#pragma unroll
          for (int gi = 0; gi < (n_groups ? kWords : 0); gi += 2) {
            const int g = half * kWords + gi;
            uint32_t v0[32], v1[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + g * 32, v0);
            tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + (g + 1) * 32, v1);
            tmem_ld_wait();
            if (gi + 2 >= kWords) {   // both of this warp's last groups are in registers: release the accumulator now
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (CG == 2 && !leader) mbar_arrive_cluster(&tempty[acc], rank - 1u);
                else mbar_arrive(&tempty[acc]);
              }
              released = true;
            }
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const uint32_t rg = row0 + (g + h2) * 32;
              const uint32_t nvalid = rg < p.n_rows ? min(32u, p.n_rows - rg) : 0u;
              const float4* ip = reinterpret_cast<const float4*>(s_invt + ib * kInvStride + (g + h2) * 32);
              float inv[32];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 f = ip[j];
                inv[4 * j] = f.x;
                inv[4 * j + 1] = f.y;
                inv[4 * j + 2] = f.z;
                inv[4 * j + 3] = f.w;
              }
              uint32_t bits = 0;
              if (h2 == 0) threshold_bits<0>(bits, v0, inv, p.tau); else threshold_bits<0>(bits, v1, inv, p.tau);
              words[gi + h2] = bits & (nvalid >= 32u ? 0xFFFFFFFFu : ((1u << nvalid) - 1u));
            }
          }
          const uint32_t f = (uint32_t)ablock * kTcM + m_local;
          if (n_groups && f < (uint32_t)p.F) {
            uint32_t* dst = p.out_bits + (size_t)f * p.words_per_filter + row0 / 32 + half * kWords;
            if constexpr (kWords == 4) *reinterpret_cast<uint4*>(dst) = make_uint4(words[0], words[1], words[2], words[3]);
            else *reinterpret_cast<uint2*>(dst) = make_uint2(words[0], words[1]);
          }
          // the inverse-norm stage is released below; the accumulator already was (unless the profiling build skipped the loop)
          __syncwarp();
          if (lane == 0) {
            if (!released) {
              tc_fence_before();
              if (CG == 2 && !leader) mbar_arrive_cluster(&tempty[acc], rank - 1u);
              else mbar_arrive(&tempty[acc]);
            }
            mbar_arrive(&iempty[ib]);
          }
        }
        if constexpr (MODE != kModeFilter) {
        // (Reading two groups per tcgen05.wait::ld and releasing the accumulator before the list work -- what the filter
        //  epilogue does -- was measured here as well: K2 lost 1 % at 10M rows and 5 % on a 1.25M-row shard under
        //  sustained load, profiles/r02_tensor_path.md; the compact single-group loop stays.)
#pragma unroll 1
        for (int g = half; g < n_groups; g += 2) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + g * 32, v);
          const uint32_t rg = row0 + g * 32;
          if (MODE == kModeTopK && t - t0 < 2u) {
            // warm-up: the pool fills within the first groups of an item; pick the bound up per group
            int unused;
            gb = fmaxf(gb, key_score(pool_min<pool_stride(KL)>(gb_ptr, p.k_real, unused)));
          }
          const uint32_t nvalid = rg < p.n_rows ? min(32u, p.n_rows - rg) : 0u;
          const float4* ip = reinterpret_cast<const float4*>(s_invt + ib * kInvStride + g * 32);   // staged by the producer
          {
            // Fast reject on RAW accumulators: score_j = acc_j * inv_j <= max_j(acc_j) / min_j(norm_j).
            // gmin = (1 - 2^-20) * min norm of the 32 rows, so `max acc <= bound * gmin` proves that no
            // row of the group reaches `bound`; one max tree + one warp vote per 32 rows.  The exact
            // arithmetic (acc * inv_norm, as the oracle) only runs in the rare slow path.
            const float gmn = s_invt[ib * kInvStride + BN + g];
            tmem_ld_wait();
            float m = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
#pragma unroll
            for (int j = 2; j < 32; j += 2) m = fmaxf(m, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
            bool hit;
            if (MODE == kModeTopK) {
              const float tb = fmaxf(top.threshold(), gb);
              hit = tb > 0.f ? (m > tb * gmn) : (nvalid > 0);
            } else {
              hit = p.tau > 0.f ? (m * inv_a >= p.tau * gmn) : (nvalid > 0);
            }
            if (__any_sync(0xffffffffu, hit)) {
              // ---- slow path (compact on purpose: one insert site, no per-column code copies) ----
              // exact scores + a per-lane bitmask of candidate columns, then candidates are pulled
              // out in ascending column (= row) order with a select tree (no dynamic register index)
              float sc[32];
              uint32_t cm = 0;
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 f = ip[j4];
                sc[4 * j4] = __uint_as_float(v[4 * j4]) * f.x;
                sc[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) * f.y;
                sc[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) * f.z;
                sc[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) * f.w;
              }
              const float thr0 = top.threshold();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                bool c;
                if (MODE == kModeTopK) {
                  c = sc[j] > thr0 && sc[j] >= gb;
                } else {
                  sc[j] *= inv_a;
                  c = sc[j] >= p.tau && a_ok && a_global < rg + j;
                }
                cm |= (c && (uint32_t)j < nvalid) ? (1u << j) : 0u;
              }
#pragma unroll 1
              while (__any_sync(0xffffffffu, cm != 0u)) {
                if (cm != 0u) {
                  const int j = __ffs(cm) - 1;
                  cm &= cm - 1u;
                  // select tree: sc[j] without indexing registers dynamically
                  float t16[16], t8[8], t4[4], t2[2];
#pragma unroll
                  for (int i = 0; i < 16; ++i) t16[i] = (j & 16) ? sc[i + 16] : sc[i];
#pragma unroll
                  for (int i = 0; i < 8; ++i) t8[i] = (j & 8) ? t16[i + 8] : t16[i];
#pragma unroll
                  for (int i = 0; i < 4; ++i) t4[i] = (j & 4) ? t8[i + 4] : t8[i];
#pragma unroll
                  for (int i = 0; i < 2; ++i) t2[i] = (j & 2) ? t4[i + 2] : t4[i];
                  const float s = (j & 1) ? t2[1] : t2[0];
                  const uint32_t row = rg + j;
                  if (MODE == kModeTopK) {
                    // (a later round of 32 < k <= 128 only admits rows ranking strictly after the previous round's last entry)
                    if (s > top.threshold() &&
                        (!has_ub || s < ub_s || (s == ub_s && (int64_t)row * p.row_stride + p.row_base > ub_r))) {
                      if (!p.use_mask || tc_mask_ok(p.mask, row, p.req)) {
                        top.insert(s, row);
                        atomicMax(gb_ptr + row % (uint32_t)p.k_real, score_key(s));
                      }
                    }
                  } else {
                    const unsigned long long slot = atomicAdd(p.out_count, 1ull);
                    if ((int64_t)slot < p.cap) {
                      p.out_i[slot] = (int64_t)a_global * p.row_stride + p.row_base;
                      p.out_j[slot] = (int64_t)row * p.row_stride + p.row_base;
                      p.out_score[slot] = s;
                    }
                  }
                }
              }
            }
          }
        }
        // this warp is done reading the accumulator buffer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2 && !leader) mbar_arrive_cluster(&tempty[acc], rank - 1u);   // the leader's MMA thread waits for both CTAs
          else mbar_arrive(&tempty[acc]);
          mbar_arrive(&iempty[ib]);
        }
        }   // MODE != kModeFilter
      }
      if (MODE == kModeTopK) {
        // partial lists: [slice][half][Bp][KL]
        const size_t base = (((size_t)(w / p.n_agroups) * 2 + half) * p.Bp + (size_t)ablock * kTcM + m_local) * KL;
#pragma unroll
        for (int j = 0; j < KL; ++j) {
          p.part_s[base + j] = top.s[j];
          p.part_r[base + j] = top.r[j] == kEmptyRow ? -1 : (int64_t)top.r[j] * p.row_stride + p.row_base;
        }
      }
    }
  }

  tc_fence_before();
  // no CTA may exit while a peer can still multicast into its smem or arrive on its barriers
  if (C > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// query preparation: normalise (f32), round to bf16, zero-pad to [Bp][Dp]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_queries_kernel(const float* __restrict__ q, int B, int dim, int Bp, int Dp,
                                                           __nv_bfloat16* __restrict__ out, uint32_t* __restrict__ gbound,
                                                           int pool_stride_words) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= Bp) return;
  if (gbound && lane < pool_stride_words) gbound[(size_t)b * pool_stride_words + lane] = score_key(VS_NEG_INF);
  if (q == nullptr) return;   // a later round of 32 < k <= 128: the queries are already in place, only the pool is reset
  __nv_bfloat16* o = out + (size_t)b * Dp;
  if (b >= B) {
    for (int e = lane; e < Dp; e += 32) o[e] = __float2bfloat16_rn(0.f);
    return;
  }
  const float* s = q + (size_t)b * dim;
  float ss = 0.f;
  for (int e = lane; e < dim; e += 32) ss = fmaf(s[e], s[e], ss);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  const float inv = 1.0f / (sqrtf(ss) + 1e-30f);
  for (int e = lane; e < Dp; e += 32) o[e] = __float2bfloat16_rn(e < dim ? s[e] * inv : 0.f);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// bf16 [rows][ld] row-major, box = 64 k-elements x box_rows, 128B swizzle, zero fill out of bounds
static cudaError_t make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t dim, uint64_t ld_elems,
                            uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorNotSupported;
  cuuint64_t gdim[2] = {dim, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)kTcKB, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// Tuning knobs.  The PRODUCT library has none: every value below is a constant.  A tuning build
// (`VS_BUILD_TUNING=1 python build.py` -> -DVS_TUNING, a separate .so loaded through VS_LIB_PATH by
// tools/bench_tensor.py) reads VS_TC_* from the environment for A/B runs on the box.
static int env_int(const char* name, int dflt) {
#ifdef VS_TUNING
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
#else
  (void)name;
  return dflt;
#endif
}
static int tc_block_n() {
  static int bn = env_int("VS_TC_BN", 256) == 128 ? 128 : 256;
  return bn;
}
// largest cluster size allowed (1, 2, 4 or 8)
static int tc_max_cluster(int dflt) {
  const int v = env_int("VS_TC_CLUSTER", dflt);
  return v >= 8 ? 8 : v >= 4 ? 4 : v >= 2 ? 2 : v >= 1 ? 1 : 0;
}

struct TcPlan {
  int BN, stages, kb_count, Dp, a_stream, cg;
  size_t smem;
  bool ok;
};
// cg_request: 0 = auto, 1 / 2 = force the issue mode
static TcPlan plan_for_cg(int dim, int cg_request) {
  TcPlan pl;
  pl.kb_count = (dim + kTcKB - 1) / kTcKB;
  pl.Dp = pl.kb_count * kTcKB;
  pl.BN = tc_block_n();
  pl.cg = 1;
  const size_t a_bytes = (size_t)pl.kb_count * kTcM * 128;
  const size_t fixed = 1024 /*alignment slack*/ + kTcBarrierBytes + (size_t)kInvRing * (256 + 8) * 4 /*inverse-norm ring*/;
  auto stages_for = [&](int bn) { return (int)(((size_t)kTcSmemMax - fixed - a_bytes) / ((size_t)bn * 128)); };
  // Resident A needs kb_count x 16 KB of smem.  At dim 768 that left 2 x 16 KB of B stages and the
  // kernel was TMA-latency bound (dedup 200k x 768: 0.45 PFLOP/s).  Streaming the A k-block with every
  // stage instead (16 KB more L2->SM traffic per stage, A block stays L2 resident) gives a 4-deep
  // ring of BN = 256 tiles at ANY dim (profiles/r01_tensor_path.md); VS_TC_ASTREAM=0 keeps A resident.
  static const int force_stream = env_int("VS_TC_ASTREAM", -1);
  const bool resident_ok = a_bytes + fixed + 2 * 128 * 128 <= (size_t)kTcSmemMax;
  pl.a_stream = force_stream >= 0 ? (force_stream != 0 || !resident_ok || stages_for(pl.BN) < 2) : 1;
  // CTA pairs (tcgen05 cta_group::2): needs BN = 256 and an even cluster; VS_TC_CG=1 = single-CTA MMAs.
  static const int want_cg = env_int("VS_TC_CG", 2);
  if (pl.a_stream && pl.BN == 256 && want_cg == 2 && cg_request != 1) pl.cg = 2;
  // Pair mode with a RESIDENT A block (when >= 4 stages of 16 KB fit beside it, i.e. dim <= 512): each CTA
  // then receives only its half B tile per stage (16 KB instead of 16 + 16 KB).  Measured at dim 512:
  // K2 1.148 -> 1.172, K3 1.059 -> 1.114, K4 1.11 -> 1.19 PFLOP/s.
  static const int ares2 = env_int("VS_TC_ARES2", 1);
  if (pl.cg == 2 && ares2 != 0) {
    const size_t stride = (size_t)pl.BN * 128 / 2;
    const int st2 = a_bytes + fixed <= (size_t)kTcSmemMax ? (int)(((size_t)kTcSmemMax - fixed - a_bytes) / stride) : 0;
    if (st2 >= 4) {
      pl.a_stream = 0;
      pl.stages = st2 > kTcMaxStages ? kTcMaxStages : st2;
      pl.ok = true;
      pl.smem = a_bytes + fixed + (size_t)pl.stages * stride;
      return pl;
    }
  }
  if (pl.a_stream) {
    const size_t stride = (size_t)pl.BN * 128 / pl.cg + (size_t)kTcM * 128;
    pl.stages = (int)(((size_t)kTcSmemMax - fixed) / stride);
    if (pl.stages > kTcMaxStages) pl.stages = kTcMaxStages;
    pl.ok = pl.stages >= 2;
    pl.smem = fixed + (size_t)pl.stages * stride;
    return pl;
  }
  pl.ok = resident_ok;
  if (pl.ok && stages_for(pl.BN) < 3) pl.BN = 128;
  pl.stages = pl.ok ? stages_for(pl.BN) : 0;
  if (pl.stages > kTcMaxStages) pl.stages = kTcMaxStages;
  if (pl.stages < 2) pl.ok = false;
  pl.smem = a_bytes + fixed + (size_t)pl.stages * pl.BN * 128;
  return pl;
}

static TcPlan plan_for(int dim) { return plan_for_cg(dim, 0); }

static int kl_for(int k) { return k <= 10 ? 10 : 32; }
// cluster size for `chunks` A blocks: the largest power of two <= min(chunks, cap)
static int cluster_for(int chunks, int cap) {
  int c = 1;
  while (c * 2 <= chunks && c * 2 <= cap) c *= 2;
  return c;
}

// workspace layout: [gbound u32 x Bp x 32][queries bf16 Bp x Dp][partial lists]
struct TcWorkspace {
  uint32_t* gbound;
  __nv_bfloat16* q;
  float* parts;
  size_t total;
};
static TcWorkspace carve_workspace(void* base, int B, int dim, int k, int sm_count) {
  const TcPlan pl = plan_for(dim);
  const int Bp = (B + kTcM - 1) / kTcM * kTcM;
  const int KL = kl_for(k);
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  TcWorkspace w;
  char* p = static_cast<char*>(base);
  size_t off = 0;
  w.gbound = reinterpret_cast<uint32_t*>(p + off);
  off += up((size_t)Bp * 32 * 4);   // union pool: up to 32 keys per query
  w.q = reinterpret_cast<__nv_bfloat16*>(p + off);
  off += up((size_t)Bp * pl.Dp * 2);
  w.parts = reinterpret_cast<float*>(p + off);
  // partial lists [n_slices][2 halves][Bp][KL] (f32 score + i64 row); n_slices <= sm_count
  off += up((size_t)2 * sm_count * Bp * KL * 4) + up((size_t)2 * sm_count * Bp * KL * 8) + 1024;
  w.total = off;
  return w;
}
size_t tensor_workspace_bytes(int B, int dim, int k, int sm_count) {
  // one launch covers at most kTcMaxBatch queries (launch_tensor_topk loops beyond that)
  return carve_workspace(nullptr, B < kTcMaxBatch ? B : kTcMaxBatch, dim, k, sm_count).total;
}
static int tc_prefetch() {
  static int v = env_int("VS_TC_PREFETCH", 0);   // measured on B200: no gain for top-k, a loss for the sweep
  return v < 0 ? 0 : v > 64 ? 64 : v;
}

// Per (kernel instantiation, smem size): the dynamic-smem attribute is set once and the number of co-resident
// clusters is asked once -- neither belongs on the per-query path.
template <int MODE, int BN, int KL, int C, int CG>
static int prepared_clusters(const TcPlan& pl, int sm_count) {
  // function attributes are per DEVICE: a single process may drive several GPUs (vs_group_t)
  static std::mutex mu;
  static size_t smem_of[kMaxDevices] = {};
  static int n_of[kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  std::lock_guard<std::mutex> lk(mu);
  size_t& cached_smem = smem_of[dev];
  int& cached_n = n_of[dev];
  if (cached_smem == pl.smem && cached_n > 0) return cached_n;
  auto kern = tc_kernel<MODE, BN, KL, C, CG>;
  int n = sm_count / C;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(sm_count / C * C), 1, 1);
  cfg.blockDim = dim3(kTcThreads, 1, 1);
  cfg.dynamicSmemBytes = pl.smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int q = 0;
  if (cudaOccupancyMaxActiveClusters(&q, kern, &cfg) == cudaSuccess && q > 0) n = q;
  else cudaGetLastError();
  cached_smem = pl.smem;
  cached_n = n;
  return n;
}

template <int MODE, int BN, int KL, int C, int CG>
static cudaError_t launch_tc_c(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, const TcPlan& pl,
                               int n_clusters_wanted, cudaStream_t st) {
  auto kern = tc_kernel<MODE, BN, KL, C, CG>;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_clusters_wanted * C), 1, 1);
  cfg.blockDim = dim3(kTcThreads, 1, 1);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p);
  count_launch();
  return e;
}

// CG = 2 (CTA pairs) exists for BN = 256 and even clusters only; `pl.cg` selects it at run time
#define VS_TC_DISPATCH_C1(FN, MODE, BN, KL, C, ...)           \
  ((C) == 8   ? FN<MODE, BN, KL, 8, 1>(__VA_ARGS__)           \
   : (C) == 4 ? FN<MODE, BN, KL, 4, 1>(__VA_ARGS__)           \
   : (C) == 2 ? FN<MODE, BN, KL, 2, 1>(__VA_ARGS__)           \
              : FN<MODE, BN, KL, 1, 1>(__VA_ARGS__))
#define VS_TC_DISPATCH_C2(FN, MODE, KL, C, ...)               \
  ((C) == 8   ? FN<MODE, 256, KL, 8, 2>(__VA_ARGS__)          \
   : (C) == 4 ? FN<MODE, 256, KL, 4, 2>(__VA_ARGS__)          \
              : FN<MODE, 256, KL, 2, 2>(__VA_ARGS__))
#define VS_TC_DISPATCH_KL(FN, MODE, BN, KL, C, CG, ...)                                            \
  ((BN) == 256 ? (((CG) == 2 && (C) >= 2) ? VS_TC_DISPATCH_C2(FN, MODE, KL, C, __VA_ARGS__)        \
                                          : VS_TC_DISPATCH_C1(FN, MODE, 256, KL, C, __VA_ARGS__))  \
               : VS_TC_DISPATCH_C1(FN, MODE, 128, KL, C, __VA_ARGS__))
#define VS_TC_DISPATCH(FN, MODE, BN, KL, C, CG, ...)                                   \
  ((KL) == 10 ? VS_TC_DISPATCH_KL(FN, MODE, BN, 10, C, CG, __VA_ARGS__)                \
              : VS_TC_DISPATCH_KL(FN, MODE, BN, 32, C, CG, __VA_ARGS__))

// filter / dedup epilogues keep no list: only the KL = 10 instantiation exists for them
#define VS_TC_DISPATCH10(FN, MODE, BN, C, CG, ...) VS_TC_DISPATCH_KL(FN, MODE, BN, 10, C, CG, __VA_ARGS__)

static void fill_common(TcParams& p, const TensorArgs& a, const TcPlan& pl) {
  memset(&p, 0, sizeof(p));
  p.inv_norm = a.inv_norm;
  p.mask = a.mask;
  p.use_mask = 0;
  for (int w = 0; w < kMaskWords; ++w) {
    p.req[w] = a.req[w];
    if (a.req[w]) p.use_mask = 1;
  }
  if (!a.mask) p.use_mask = 0;
  p.row_base = a.row_base;
  p.row_stride = a.row_stride;
  p.n_rows = (uint32_t)a.n_rows;
  p.n_tiles = (uint32_t)((a.n_rows + pl.BN - 1) / pl.BN);
  p.kb_count = pl.kb_count;
  p.stages = pl.stages;
  p.a_stream = pl.a_stream;
  p.prefetch = tc_prefetch();
  p.n_agroups = 1;
}

// n_clusters resident clusters, n_agroups A groups: (slices x A groups) work items, the A groups of one slice on
// neighbouring clusters so that they stream the same tiles together
static void plan_slices(TcParams& p, int n_clusters, int n_agroups) {
  p.n_agroups = n_agroups;
  int slices = n_clusters / n_agroups;
  if (slices < 1) slices = 1;
  if (slices > (int)p.n_tiles) slices = (int)p.n_tiles;
  p.tiles_per_slice = (int)((p.n_tiles + slices - 1) / slices);
  p.n_slices = (int)((p.n_tiles + p.tiles_per_slice - 1) / p.tiles_per_slice);
  p.n_items = p.n_slices * n_agroups;
}

static bool dims_ok(const TensorArgs& a) { return tensor_dim_ok(a.dim) && a.ld_elems % 8 == 0; }
bool tensor_dim_ok(int dim) { return dim >= 8 && dim % 8 == 0 && dim <= 4096; }

// ub extraction for the rounds of 32 < k <= 128: nothing to launch -- round r reads the (score, row) at
// out[b][32 r - 1] of the result buffer itself (written by round r-1's merge, stream ordered).
cudaError_t launch_tensor_topk(const TensorArgs& a, const float* q, int B, int k, void* workspace, float* out_s,
                               int64_t* out_r, int sm_count, cudaStream_t st) {
  const TcPlan pl = plan_for(a.dim);
  if (!pl.ok || !dims_ok(a) || k > kMaxTensorK || k <= 0 || B <= 0 || !a.gmin) return cudaErrorNotSupported;
  // Cluster choice for K2 (profiles/r02_tensor_path.md, same-box A/B under sustained load):
  //  * clusters of 8 (one corpus stream multicast to 8 query blocks) move every tile L2->SM once per 1024 queries: the
  //    most energy-efficient form, best on long shards where the card sits at its power cap (10M rows: 9.45 ms vs 9.99);
  //  * CTA PAIRS with several A groups per corpus slice (the pairs (slice, 0..n_agroups-1) sit on neighbouring SMs and
  //    stream the same tiles in lockstep: HBM once, L2 for the other groups) fill all 148 SMs where clusters of 8 only fit
  //    15 times (120 SMs): best on shorter shards (sustained, same box: 1.25M rows 1.191 vs 1.256 ms, 2.5M 2.332 vs 2.460,
  //    5M 4.691 vs 4.809; at 10M rows the A groups of a slice drift apart and the re-reads reach HBM: 10.30 vs 9.53 ms).
  static const int cap_env = tc_max_cluster(0);
  const int cap = cap_env > 0 ? cap_env : (a.n_rows >= kTopkPairRowsBelow ? 8 : 2);
  for (int b0 = 0; b0 < B; b0 += kTcMaxBatch) {
    const int nb = B - b0 < kTcMaxBatch ? B - b0 : kTcMaxBatch;
    const int Bp = (nb + kTcM - 1) / kTcM * kTcM;
    const int chunks = Bp / kTcM;
    const int C = cluster_for(chunks, cap);
    const int n_agroups = (chunks + C - 1) / C;
    const int Bpp = n_agroups * C * kTcM;             // padded to whole A groups
    const int cg = C >= 2 ? pl.cg : 1;
    const TcPlan plc = plan_for_cg(a.dim, cg);
    const int k_list = k <= 32 ? k : 32;              // per round
    const int KL = kl_for(k_list);
    const int KLP = pool_stride(KL);
    const TcWorkspace ws = carve_workspace(workspace, Bpp, a.dim, k_list, sm_count);
    const int ncl = VS_TC_DISPATCH(prepared_clusters, kModeTopK, plc.BN, KL, C, cg, plc, sm_count);
    if (ncl <= 0) return cudaErrorLaunchOutOfResources;
    CUtensorMap tmA, tmB;
    cudaError_t e = make_map(&tmA, ws.q, (uint64_t)Bpp, pl.Dp, pl.Dp, kTcM);
    if (e != cudaSuccess) return e;
    e = make_map(&tmB, a.rows, a.n_rows, a.dim, a.ld_elems, plc.BN / C);
    if (e != cudaSuccess) return e;
    float* o_s = out_s + (size_t)b0 * k;
    int64_t* o_r = out_r + (size_t)b0 * k;
    for (int k0 = 0; k0 < k; k0 += 32) {              // rounds: ranks [k0, k0 + 32) of every query
      const int kr = k - k0 < 32 ? k - k0 : 32;
      // (re)normalise + round the queries only once; the pool of class maxima is reset every round
      prep_queries_kernel<<<(Bpp + 7) / 8, 256, 0, st>>>(k0 == 0 ? q + (size_t)b0 * a.dim : nullptr, nb, a.dim, Bpp, pl.Dp,
                                                         ws.q, ws.gbound, KLP);
      count_launch();
      TcParams p;
      fill_common(p, a, plc);
      p.gmin = a.gmin;
      p.gbound = ws.gbound;
      p.k_real = kr;
      p.F = nb;
      if (k0 > 0) {
        p.ub_s = o_s + (k0 - 1);
        p.ub_r = o_r + (k0 - 1);
        p.ub_stride = k;
      }
      plan_slices(p, ncl, n_agroups);
      p.Bp = Bpp;
      p.part_s = ws.parts;
      p.part_r = reinterpret_cast<int64_t*>(((uintptr_t)(ws.parts + (size_t)2 * p.n_slices * p.Bp * KL) + 255) & ~(uintptr_t)255);
      const int grid_clusters = p.n_items < ncl ? p.n_items : ncl;
      e = VS_TC_DISPATCH(launch_tc_c, kModeTopK, plc.BN, KL, C, cg, tmA, tmB, p, plc, grid_clusters, st);
      if (e != cudaSuccess) return e;
      e = launch_merge_ex(p.part_s, p.part_r, 2 * p.n_slices, p.Bp, nb, KL, kr, o_s + k0, o_r + k0, st, k);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

cudaError_t launch_tensor_filter(const TensorArgs& a, const float* prompts, int F, float tau, void* workspace,
                                 uint32_t* out_bits, int64_t words_per_filter, int sm_count, cudaStream_t st) {
  const TcPlan pl = plan_for(a.dim);
  if (!pl.ok || !dims_ok(a) || F <= 0) return cudaErrorNotSupported;
  // K3 default cluster cap (tuned on B200, profiles/r02_tensor_path.md)
  static const int cap = tc_max_cluster(kFilterClusterCap);
  for (int f0 = 0; f0 < F; f0 += kTcMaxBatch) {
    const int nf = F - f0 < kTcMaxBatch ? F - f0 : kTcMaxBatch;
    const int Bp = (nf + kTcM - 1) / kTcM * kTcM;
    const int chunks = Bp / kTcM;
    const int C = cluster_for(chunks, cap);
    const int n_agroups = (chunks + C - 1) / C;
    const int Bpp = n_agroups * C * kTcM;
    const int cg = C >= 2 ? pl.cg : 1;
    const TcPlan plc = plan_for_cg(a.dim, cg);
    const TcWorkspace ws = carve_workspace(workspace, Bpp, a.dim, 1, sm_count);
    prep_queries_kernel<<<(Bpp + 7) / 8, 256, 0, st>>>(prompts + (size_t)f0 * a.dim, nf, a.dim, Bpp, pl.Dp, ws.q, nullptr, 0);
    count_launch();
    TcParams p;
    fill_common(p, a, plc);
    const int ncl = VS_TC_DISPATCH10(prepared_clusters, kModeFilter, plc.BN, C, cg, plc, sm_count);
    if (ncl <= 0) return cudaErrorLaunchOutOfResources;
    plan_slices(p, ncl, n_agroups);
    p.out_bits = out_bits + (size_t)f0 * words_per_filter;
    p.words_per_filter = words_per_filter;
    p.tau = tau;
    p.F = nf;
    CUtensorMap tmA, tmB;
    cudaError_t e = make_map(&tmA, ws.q, (uint64_t)Bpp, pl.Dp, pl.Dp, kTcM);
    if (e != cudaSuccess) return e;
    e = make_map(&tmB, a.rows, a.n_rows, a.dim, a.ld_elems, plc.BN / C);
    if (e != cudaSuccess) return e;
    const int grid_clusters = p.n_items < ncl ? p.n_items : ncl;
    e = VS_TC_DISPATCH10(launch_tc_c, kModeFilter, plc.BN, C, cg, tmA, tmB, p, plc, grid_clusters, st);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_tensor_dedup(const TensorArgs& a, int64_t row_lo, int64_t row_hi, float tau, int64_t cap,
                                int64_t* out_i, int64_t* out_j, float* out_score, unsigned long long* out_count,
                                void* workspace, int sm_count, cudaStream_t st) {
  const TcPlan pl = plan_for(a.dim);
  if (!pl.ok || !dims_ok(a) || !a.gmin) return cudaErrorNotSupported;
  (void)workspace;
  static const int ccap = tc_max_cluster(8);
  const int64_t a_row_min = row_lo;
  row_lo = row_lo / kTcM * kTcM;   // A blocks are 128-row aligned; rows below the caller's row_lo are filtered out
  const int n_blocks = (int)((row_hi - row_lo + kTcM - 1) / kTcM);
  const int C = cluster_for(n_blocks, ccap);
  const int cg = C >= 2 ? pl.cg : 1;
  const TcPlan plc = plan_for_cg(a.dim, cg);
  TcParams p;
  fill_common(p, a, plc);
  p.gmin = a.gmin;
  p.a_row_min = a_row_min;
  p.a_row_lo = row_lo;
  p.a_row_hi = row_hi;
  p.n_items = (n_blocks + C - 1) / C;
  p.tau = tau;
  p.cap = cap;
  p.out_i = out_i;
  p.out_j = out_j;
  p.out_score = out_score;
  p.out_count = out_count;
  CUtensorMap tmA, tmB;
  cudaError_t e = make_map(&tmA, a.rows, a.n_rows, a.dim, a.ld_elems, kTcM);
  if (e != cudaSuccess) return e;
  e = make_map(&tmB, a.rows, a.n_rows, a.dim, a.ld_elems, plc.BN / C);
  if (e != cudaSuccess) return e;
  int ncl = VS_TC_DISPATCH10(prepared_clusters, kModeDedup, plc.BN, C, cg, plc, sm_count);
  if (ncl <= 0) return cudaErrorLaunchOutOfResources;
  if (ncl > p.n_items) ncl = p.n_items;
  return VS_TC_DISPATCH10(launch_tc_c, kModeDedup, plc.BN, C, cg, tmA, tmB, p, plc, ncl, st);
}

}  // namespace vs
