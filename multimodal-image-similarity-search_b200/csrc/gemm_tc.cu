// gemm_tc.cu -- K2/K3/K4: batched cosine scoring on the 5th-gen tensor cores (sm_100a).
//
//   K2  batched top-k      : batched form of Collection.query (backend/app/main.py:761-765)
//   K3  filter sweep       : F prompt embeddings x N rows, thresholded to a bit mask
//   K4  all-pairs dedup    : rows x rows, pairs with cos >= tau
//
// One warp-specialised kernel, three epilogues.  D[m, n] = sum_k A[m, k] * B[n, k] with
//   A = a block of 128 "query-side" vectors (queries / prompts / a block of corpus rows), bf16,
//       RESIDENT in shared memory for the whole work item (loaded once by TMA, 128B swizzle),
//   B = corpus rows, streamed from HBM in tiles of BN rows x 64 k-elements by TMA into a ring of
//       stages (the corpus is read once per A block; CTAs that share a corpus slice run in
//       lock-step so the re-reads are L2 hits),
//   D = fp32 accumulators in TMEM: lane = query, column = corpus row, 512/BN buffers so the MMA
//       of tile t+1 overlaps the epilogue of tile t.
// Roles: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected thread), warps 2..5 =
// epilogue (tcgen05.ld 32x32b: each thread owns ONE query and sees that query's scores against
// 32 consecutive rows per load -> a private register top-k list, no cross-thread traffic).
// Epilogue arithmetic: score = acc * inv_norm[row] (queries are L2-normalised then rounded to
// bf16 by prep_queries_kernel; the oracle does the same), filter bits are only fetched for rows
// that would enter a list, nothing but the final candidates is written.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace vs {

constexpr int kTcM = 128;          // A rows per block (UMMA M)
constexpr int kTcKB = 64;          // k elements per stage (= 128 B = one swizzle span)
constexpr int kTcThreads = 192;
constexpr int kTcMaxStages = 8;
constexpr int kTcSmemMax = 232448; // 227 KB

enum { kModeTopK = 0, kModeFilter = 1, kModeDedup = 2 };

struct TcParams {
  const float* inv_norm;
  const uint64_t* mask;
  uint64_t req[kMaskWords];
  int use_mask;
  int64_t row_base;
  uint32_t n_rows;
  uint32_t n_tiles;        // ceil(n_rows / BN)
  int kb_count;            // ceil(dim / 64)
  int stages;
  int n_chunks;            // A blocks
  int n_slices;            // corpus slices (top-k / filter)
  int tiles_per_slice;
  int n_items;
  // top-k
  float* part_s;           // [n_slices][Bp][KL]
  int64_t* part_r;
  int Bp;
  // filter
  uint32_t* out_bits;
  int64_t words_per_filter;
  float tau;
  int F;
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
// PTX wrappers: TMA tensor loads, tcgen05 (alloc / mma / commit / ld / fences)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "l"(policy)
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

__device__ __forceinline__ bool tc_mask_ok(const uint64_t* mask, uint32_t row, const uint64_t* req) {
  const uint64_t* m = mask + (size_t)row * kMaskWords;
  bool ok = true;
#pragma unroll
  for (int w = 0; w < kMaskWords; ++w) ok = ok && ((__ldg(m + w) & req[w]) == req[w]);
  return ok;
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int MODE, int BN, int KL>
__global__ void __launch_bounds__(kTcThreads, 1)
    tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  constexpr int ACC = 512 / BN;              // accumulator buffers in TMEM
  constexpr uint32_t kStageBytes = BN * 128; // BN rows x 64 bf16
  constexpr uint32_t kABlockBytes = kTcM * 128;
  constexpr uint32_t kIdesc = make_idesc(kTcM, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 128B swizzle: 1024-B aligned
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.kb_count * kABlockBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + (size_t)p.stages * kStageBytes);
  uint64_t* empty = full + kTcMaxStages;
  uint64_t* tfull = empty + kTcMaxStages;
  uint64_t* tempty = tfull + 4;
  uint64_t* a_full = tempty + 4;
  uint64_t* a_empty = a_full + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(a_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // work item -> (A block, tile range); identical sequence in every role
  auto item_range = [&](int w, int& chunk, uint32_t& t0, uint32_t& t1) {
    if (MODE == kModeDedup) {
      chunk = w;
      const uint32_t a_row0 = (uint32_t)p.a_row_lo + (uint32_t)w * kTcM;
      t0 = a_row0 / BN;       // only columns j > i can pair with row i
      t1 = p.n_tiles;
    } else {
      chunk = w % p.n_chunks;
      const int slice = w / p.n_chunks;
      t0 = (uint32_t)slice * p.tiles_per_slice;
      t1 = min(t0 + (uint32_t)p.tiles_per_slice, p.n_tiles);
    }
  };

  if (warp == 0) {
    // ================= TMA producer =========================================================
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();
      const uint64_t pol_keep = policy_evict_last();
      uint32_t it = 0, n_item = 0;
      for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++n_item) {
        int chunk;
        uint32_t t0, t1;
        item_range(w, chunk, t0, t1);
        // A block: resident for the whole item
        if (n_item > 0) mbar_wait(a_empty, (n_item - 1) & 1);
        mbar_expect_tx(a_full, (uint32_t)p.kb_count * kABlockBytes);
        const int a_row = MODE == kModeDedup ? (int)(p.a_row_lo + (int64_t)chunk * kTcM) : chunk * kTcM;
        for (int kb = 0; kb < p.kb_count; ++kb)
          tma_load_2d(sA + (size_t)kb * kABlockBytes, &tmA, kb * kTcKB, a_row, a_full, pol_keep);
        for (uint32_t t = t0; t < t1; ++t) {
          for (int kb = 0; kb < p.kb_count; ++kb, ++it) {
            const int st = it % p.stages;
            const uint32_t use = it / p.stages;
            if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
            mbar_expect_tx(&full[st], kStageBytes);
            tma_load_2d(sB + (size_t)st * kStageBytes, &tmB, kb * kTcKB, (int)(t * BN), &full[st], pol_stream);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer ===============================================================
    if (lane == 0) {
      uint32_t it = 0, n_item = 0, tile_ctr = 0;
      for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++n_item) {
        int chunk;
        uint32_t t0, t1;
        item_range(w, chunk, t0, t1);
        mbar_wait(a_full, n_item & 1);
        tc_fence_after();
        for (uint32_t t = t0; t < t1; ++t, ++tile_ctr) {
          const uint32_t acc = tile_ctr % ACC;
          const uint32_t use = tile_ctr / ACC;
          if (use > 0) mbar_wait(&tempty[acc], (use - 1) & 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = 0; kb < p.kb_count; ++kb, ++it) {
            const int st = it % p.stages;
            mbar_wait(&full[st], (it / p.stages) & 1);
            tc_fence_after();
            const uint64_t da = make_smem_desc(smem_u32(sA + (size_t)kb * kABlockBytes));
            const uint64_t db = make_smem_desc(smem_u32(sB + (size_t)st * kStageBytes));
#pragma unroll
            for (int k = 0; k < kTcKB / 16; ++k)   // UMMA_K = 16 bf16 = 32 B: advance start address by 2 (16-B units)
              umma_bf16(d_tmem, da + 2 * k, db + 2 * k, kIdesc, (kb | k) != 0);
            umma_commit(&empty[st]);               // frees the stage when these MMAs have read it
          }
          umma_commit(&tfull[acc]);                // accumulator ready for the epilogue
        }
        umma_commit(a_empty);                      // A block may be overwritten
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps (2..5): TMEM lane quarter = warp % 4 ========================
    const int quarter = warp & 3;
    const int m_local = quarter * 32 + lane;       // A row (query) owned by this thread
    ThreadTopK<KL> top;
    uint32_t tile_ctr = 0;
    for (int w = blockIdx.x; w < p.n_items; w += gridDim.x) {
      int chunk;
      uint32_t t0, t1;
      item_range(w, chunk, t0, t1);
      if (MODE == kModeTopK) top.init();
      const uint32_t a_global = (MODE == kModeDedup ? (uint32_t)p.a_row_lo : 0u) + (uint32_t)chunk * kTcM + m_local;
      float inv_a = 1.f;
      if (MODE == kModeDedup) inv_a = a_global < p.n_rows ? __ldg(p.inv_norm + a_global) : 0.f;
      for (uint32_t t = t0; t < t1; ++t, ++tile_ctr) {
        const uint32_t acc = tile_ctr % ACC;
        mbar_wait(&tfull[acc], (tile_ctr / ACC) & 1);
        tc_fence_after();
        const uint32_t row0 = t * BN;
#pragma unroll 1
        for (int g = 0; g < BN / 32; ++g) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + g * 32, v);
          const uint32_t rg = row0 + g * 32;
          float inv[32];
          {
            const float4* ip = reinterpret_cast<const float4*>(p.inv_norm + rg);   // padded past n_rows
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 f = __ldg(ip + j);
              inv[4 * j] = f.x;
              inv[4 * j + 1] = f.y;
              inv[4 * j + 2] = f.z;
              inv[4 * j + 3] = f.w;
            }
          }
          tmem_ld_wait();
          const uint32_t nvalid = rg < p.n_rows ? min(32u, p.n_rows - rg) : 0u;
          if (MODE == kModeTopK) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = __uint_as_float(v[j]) * inv[j];
              if (s > top.threshold() && (uint32_t)j < nvalid) {
                const uint32_t row = rg + j;
                if (!p.use_mask || tc_mask_ok(p.mask, row, p.req)) top.insert(s, row);
              }
            }
          } else if (MODE == kModeFilter) {
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = __uint_as_float(v[j]) * inv[j];
              bits |= (s >= p.tau && (uint32_t)j < nvalid) ? (1u << j) : 0u;
            }
            const uint32_t f = (uint32_t)chunk * kTcM + m_local;
            if (f < (uint32_t)p.F) p.out_bits[(size_t)f * p.words_per_filter + rg / 32] = bits;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = __uint_as_float(v[j]) * inv_a * inv[j];
              const uint32_t row = rg + j;
              if (s >= p.tau && (uint32_t)j < nvalid && a_global < row && a_global < (uint32_t)p.a_row_hi &&
                  a_global >= (uint32_t)p.a_row_min) {
                const unsigned long long slot = atomicAdd(p.out_count, 1ull);
                if ((int64_t)slot < p.cap) {
                  p.out_i[slot] = (int64_t)a_global + p.row_base;
                  p.out_j[slot] = (int64_t)row + p.row_base;
                  p.out_score[slot] = s;
                }
              }
            }
          }
        }
        // this warp is done reading the accumulator buffer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      }
      if (MODE == kModeTopK) {
        const int slice = w / p.n_chunks;
        const size_t base = ((size_t)slice * p.Bp + (size_t)chunk * kTcM + m_local) * KL;
#pragma unroll
        for (int j = 0; j < KL; ++j) {
          p.part_s[base + j] = top.s[j];
          p.part_r[base + j] = top.r[j] == kEmptyRow ? -1 : (int64_t)top.r[j] + p.row_base;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// query preparation: normalise (f32), round to bf16, zero-pad to [Bp][Dp]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_queries_kernel(const float* __restrict__ q, int B, int dim, int Bp, int Dp,
                                                           __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= Bp) return;
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

bool tensor_path_available() { return true; }

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

static int tc_block_n() {
  static int bn = [] {
    const char* e = getenv("VS_TC_BN");
    return (e && atoi(e) == 128) ? 128 : 256;
  }();
  return bn;
}

struct TcPlan {
  int BN, stages, kb_count, Dp;
  size_t smem;
  bool ok;
};
static TcPlan plan_for(int dim) {
  TcPlan pl;
  pl.kb_count = (dim + kTcKB - 1) / kTcKB;
  pl.Dp = pl.kb_count * kTcKB;
  pl.BN = tc_block_n();
  const size_t a_bytes = (size_t)pl.kb_count * kTcM * 128;
  const size_t fixed = 1024 /*alignment slack*/ + 256 /*barriers*/;
  auto stages_for = [&](int bn) { return (int)(((size_t)kTcSmemMax - fixed - a_bytes) / ((size_t)bn * 128)); };
  pl.ok = a_bytes + fixed + 2 * 128 * 128 <= (size_t)kTcSmemMax;
  if (pl.ok && stages_for(pl.BN) < 3) pl.BN = 128;
  pl.stages = pl.ok ? stages_for(pl.BN) : 0;
  if (pl.stages > kTcMaxStages) pl.stages = kTcMaxStages;
  if (pl.stages < 2) pl.ok = false;
  pl.smem = a_bytes + fixed + (size_t)pl.stages * pl.BN * 128;
  return pl;
}

static int kl_for(int k) { return k <= 10 ? 10 : 32; }

size_t tensor_workspace_bytes(int B, int dim, int k, int sm_count) {
  const TcPlan pl = plan_for(dim);
  const int Bp = (B + kTcM - 1) / kTcM * kTcM;
  const int KL = kl_for(k);
  size_t q = (size_t)Bp * pl.Dp * 2;
  q = (q + 255) & ~(size_t)255;
  const size_t parts = (size_t)sm_count * Bp * KL * 12 + 512;
  return q + parts;
}

template <int MODE, int BN, int KL>
static cudaError_t launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, const TcPlan& pl, int grid,
                             cudaStream_t st) {
  auto kern = tc_kernel<MODE, BN, KL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, kTcThreads, pl.smem, st>>>(tmA, tmB, p);
  count_launch();
  return cudaGetLastError();
}

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
  p.n_rows = (uint32_t)a.n_rows;
  p.n_tiles = (uint32_t)((a.n_rows + pl.BN - 1) / pl.BN);
  p.kb_count = pl.kb_count;
  p.stages = pl.stages;
}

static bool dims_ok(const TensorArgs& a) { return a.dim >= 8 && a.dim % 8 == 0 && a.ld_elems % 8 == 0; }

cudaError_t launch_tensor_topk(const TensorArgs& a, const float* q, int B, int k, void* workspace, float* out_s,
                               int64_t* out_r, int sm_count, cudaStream_t st) {
  const TcPlan pl = plan_for(a.dim);
  if (!pl.ok || !dims_ok(a) || k > 32 || B <= 0) return cudaErrorNotSupported;
  const int Bp = (B + kTcM - 1) / kTcM * kTcM;
  const int KL = kl_for(k);
  __nv_bfloat16* qb = static_cast<__nv_bfloat16*>(workspace);
  size_t qbytes = ((size_t)Bp * pl.Dp * 2 + 255) & ~(size_t)255;
  float* part_s = reinterpret_cast<float*>(static_cast<char*>(workspace) + qbytes);

  prep_queries_kernel<<<(Bp + 7) / 8, 256, 0, st>>>(q, B, a.dim, Bp, pl.Dp, qb);
  count_launch();

  const int chunks_total = Bp / kTcM;
  // process at most `sm_count` query chunks per launch (each chunk needs >= 1 CTA)
  for (int c0 = 0; c0 < chunks_total; c0 += sm_count) {
    const int nch = min(sm_count, chunks_total - c0);
    TcParams p;
    fill_common(p, a, pl);
    p.n_chunks = nch;
    p.n_slices = max(1, min(sm_count / nch, (int)p.n_tiles));
    p.tiles_per_slice = (int)((p.n_tiles + p.n_slices - 1) / p.n_slices);
    p.n_slices = (int)((p.n_tiles + p.tiles_per_slice - 1) / p.tiles_per_slice);
    p.n_items = p.n_chunks * p.n_slices;
    p.Bp = nch * kTcM;
    p.part_s = part_s;
    p.part_r = reinterpret_cast<int64_t*>(part_s + (size_t)p.n_slices * p.Bp * KL + 64);
    p.part_r = reinterpret_cast<int64_t*>(((uintptr_t)p.part_r + 15) & ~(uintptr_t)15);
    CUtensorMap tmA, tmB;
    cudaError_t e = make_map(&tmA, qb + (size_t)c0 * kTcM * pl.Dp, (uint64_t)nch * kTcM, pl.Dp, pl.Dp, kTcM);
    if (e != cudaSuccess) return e;
    e = make_map(&tmB, a.rows, a.n_rows, a.dim, a.ld_elems, pl.BN);
    if (e != cudaSuccess) return e;
    const int grid = p.n_items;
    if (pl.BN == 256)
      e = KL == 10 ? launch_tc<kModeTopK, 256, 10>(tmA, tmB, p, pl, grid, st)
                   : launch_tc<kModeTopK, 256, 32>(tmA, tmB, p, pl, grid, st);
    else
      e = KL == 10 ? launch_tc<kModeTopK, 128, 10>(tmA, tmB, p, pl, grid, st)
                   : launch_tc<kModeTopK, 128, 32>(tmA, tmB, p, pl, grid, st);
    if (e != cudaSuccess) return e;
    const int nb = min(B - c0 * kTcM, nch * kTcM);
    e = launch_merge_ex(p.part_s, p.part_r, p.n_slices, p.Bp, nb, KL, k, out_s + (size_t)c0 * kTcM * k,
                        out_r + (size_t)c0 * kTcM * k, st);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_tensor_filter(const TensorArgs& a, const float* prompts, int F, float tau, void* workspace,
                                 uint32_t* out_bits, int64_t words_per_filter, int sm_count, cudaStream_t st) {
  const TcPlan pl = plan_for(a.dim);
  if (!pl.ok || !dims_ok(a) || F <= 0) return cudaErrorNotSupported;
  const int Bp = (F + kTcM - 1) / kTcM * kTcM;
  __nv_bfloat16* qb = static_cast<__nv_bfloat16*>(workspace);
  prep_queries_kernel<<<(Bp + 7) / 8, 256, 0, st>>>(prompts, F, a.dim, Bp, pl.Dp, qb);
  count_launch();
  const int chunks_total = Bp / kTcM;
  for (int c0 = 0; c0 < chunks_total; c0 += sm_count) {
    const int nch = min(sm_count, chunks_total - c0);
    TcParams p;
    fill_common(p, a, pl);
    p.n_chunks = nch;
    p.n_slices = max(1, min(sm_count / nch, (int)p.n_tiles));
    p.tiles_per_slice = (int)((p.n_tiles + p.n_slices - 1) / p.n_slices);
    p.n_slices = (int)((p.n_tiles + p.tiles_per_slice - 1) / p.tiles_per_slice);
    p.n_items = p.n_chunks * p.n_slices;
    p.out_bits = out_bits + (size_t)c0 * kTcM * words_per_filter;
    p.words_per_filter = words_per_filter;
    p.tau = tau;
    p.F = F - c0 * kTcM;
    CUtensorMap tmA, tmB;
    cudaError_t e = make_map(&tmA, qb + (size_t)c0 * kTcM * pl.Dp, (uint64_t)nch * kTcM, pl.Dp, pl.Dp, kTcM);
    if (e != cudaSuccess) return e;
    e = make_map(&tmB, a.rows, a.n_rows, a.dim, a.ld_elems, pl.BN);
    if (e != cudaSuccess) return e;
    e = pl.BN == 256 ? launch_tc<kModeFilter, 256, 10>(tmA, tmB, p, pl, p.n_items, st)
                     : launch_tc<kModeFilter, 128, 10>(tmA, tmB, p, pl, p.n_items, st);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_tensor_dedup(const TensorArgs& a, int64_t row_lo, int64_t row_hi, float tau, int64_t cap,
                                int64_t* out_i, int64_t* out_j, float* out_score, unsigned long long* out_count,
                                void* workspace, int sm_count, cudaStream_t st) {
  (void)workspace;
  const TcPlan pl = plan_for(a.dim);
  if (!pl.ok || !dims_ok(a)) return cudaErrorNotSupported;
  TcParams p;
  fill_common(p, a, pl);
  p.a_row_min = row_lo;
  row_lo = row_lo / kTcM * kTcM;   // A blocks are 128-row aligned; rows below the caller's row_lo are filtered out
  p.a_row_lo = row_lo;
  p.a_row_hi = row_hi;
  p.n_chunks = (int)((row_hi - row_lo + kTcM - 1) / kTcM);
  p.n_items = p.n_chunks;
  p.tau = tau;
  p.cap = cap;
  p.out_i = out_i;
  p.out_j = out_j;
  p.out_score = out_score;
  p.out_count = out_count;
  CUtensorMap tmA, tmB;
  cudaError_t e = make_map(&tmA, a.rows, a.n_rows, a.dim, a.ld_elems, kTcM);
  if (e != cudaSuccess) return e;
  e = make_map(&tmB, a.rows, a.n_rows, a.dim, a.ld_elems, pl.BN);
  if (e != cudaSuccess) return e;
  const int grid = min(sm_count, p.n_items);
  return pl.BN == 256 ? launch_tc<kModeDedup, 256, 10>(tmA, tmB, p, pl, grid, st)
                      : launch_tc<kModeDedup, 128, 10>(tmA, tmB, p, pl, grid, st);
}

}  // namespace vs
