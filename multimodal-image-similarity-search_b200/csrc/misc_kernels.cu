// misc_kernels.cu -- K6 ingest (cast + inverse norm), export, multimodal blend, K5 candidate
// merge (block bitonic), and the exact radix select used for k > 128.
#include "common.cuh"
#include "exchange.cuh"
#include "kernels.h"

namespace vs {

// ------------------------------------------------------------------------------------------
// K6 ingest.  Collection.add (backend/app/main.py:735-740): one warp per row; the row is cast
// to the storage dtype, the pad up to the pitch is zeroed, and 1/(||stored row|| + 1e-30) is
// written beside it (hnswlib's cosine space normalises at insert the same way).
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float store_elem(T* dst, float v);
template <>
__device__ __forceinline__ float store_elem<float>(float* dst, float v) {
  *dst = v;
  return v;
}
template <>
__device__ __forceinline__ float store_elem<__nv_bfloat16>(__nv_bfloat16* dst, float v) {
  const __nv_bfloat16 b = __float2bfloat16_rn(v);
  *dst = b;
  return __bfloat162float(b);
}

template <typename T>
__global__ void __launch_bounds__(256) ingest_kernel(const float* __restrict__ src, int64_t n, int dim, T* __restrict__ dst,
                                                     int64_t ld, float* __restrict__ inv) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < n; row += nwarps) {
    const float* s = src + row * dim;
    T* d = dst + row * ld;
    float ss = 0.f;
    for (int e = lane; e < (int)ld; e += 32) {
      const float v = e < dim ? s[e] : 0.f;
      const float r = store_elem<T>(d + e, v);
      ss = fmaf(r, r, ss);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if (lane == 0) inv[row] = 1.0f / (sqrtf(ss) + 1e-30f);
  }
}

cudaError_t launch_ingest(const float* src, int64_t n, int dim, int dtype, void* dst_rows, int64_t ld_elems,
                          float* dst_inv, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t blocks = min((int64_t)148 * 8, (n + 7) / 8);
  if (dtype == 0)
    ingest_kernel<float><<<(int)blocks, 256, 0, st>>>(src, n, dim, static_cast<float*>(dst_rows), ld_elems, dst_inv);
  else
    ingest_kernel<__nv_bfloat16>
        <<<(int)blocks, 256, 0, st>>>(src, n, dim, static_cast<__nv_bfloat16*>(dst_rows), ld_elems, dst_inv);
  count_launch();
  return cudaGetLastError();
}

// inverse norms of rows that are ALREADY in the storage dtype (persistence slab reload, vs_add_raw_host):
// the same arithmetic as ingest_kernel applied to the stored values, so a reloaded collection scores
// bit-identically to the one that was saved.
template <typename T>
__global__ void __launch_bounds__(256) renorm_kernel(const T* __restrict__ rows, int64_t first, int64_t n, int dim, int64_t ld,
                                                     float* __restrict__ inv) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp0; i < n; i += nwarps) {
    const T* r = rows + (first + i) * ld;
    float ss = 0.f;
    for (int e = lane; e < (int)ld; e += 32) {
      const float v = e < dim ? (float)r[e] : 0.f;
      ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if (lane == 0) inv[first + i] = 1.0f / (sqrtf(ss) + 1e-30f);
  }
}
cudaError_t launch_renorm(const void* rows, int64_t first, int64_t n, int dim, int dtype, int64_t ld_elems, float* inv,
                          cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t blocks = min((int64_t)148 * 8, (n + 7) / 8);
  if (dtype == 0)
    renorm_kernel<float><<<(int)blocks, 256, 0, st>>>(static_cast<const float*>(rows), first, n, dim, ld_elems, inv);
  else
    renorm_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(rows), first, n, dim, ld_elems, inv);
  count_launch();
  return cudaGetLastError();
}

template <typename T>
__global__ void __launch_bounds__(256) export_kernel(const T* __restrict__ rows, int64_t n, int dim, int64_t ld,
                                                     float* __restrict__ dst) {
  const int64_t total = n * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dim;
    const int e = (int)(i - r * dim);
    dst[i] = (float)rows[r * ld + e];
  }
}

cudaError_t launch_export(const void* rows, int64_t n, int dim, int dtype, int64_t ld_elems, float* dst,
                          cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t blocks = min((int64_t)148 * 8, (n * dim + 255) / 256);
  if (dtype == 0)
    export_kernel<float><<<(int)blocks, 256, 0, st>>>(static_cast<const float*>(rows), n, dim, ld_elems, dst);
  else
    export_kernel<__nv_bfloat16>
        <<<(int)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(rows), n, dim, ld_elems, dst);
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// multimodal blend, search_multimodal (backend/app/main.py:850-860): one warp per query pair.
//   i^ = i/||i||, t^ = t/||t||, c = w*i^ + (1-w)*t^, out = c/||c||      (all float32)
// Same statement order as the reference's numpy code so the result matches it to rounding of
// the norm reductions.  No zero-norm guard there either; here a zero norm yields zeros.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) blend_kernel(const float* __restrict__ img, const float* __restrict__ txt,
                                                    const double* __restrict__ w, int B, int dim, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const float* ip = img + (size_t)b * dim;
  const float* tp = txt + (size_t)b * dim;
  float si = 0.f, stt = 0.f;
  for (int e = lane; e < dim; e += 32) {
    si = fmaf(ip[e], ip[e], si);
    stt = fmaf(tp[e], tp[e], stt);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    si += __shfl_xor_sync(0xffffffffu, si, off);
    stt += __shfl_xor_sync(0xffffffffu, stt, off);
  }
  const float ni = sqrtf(si), nt = sqrtf(stt);
  // numpy semantics of the reference: the python-float weight and (1 - weight) are each
  // rounded to float32 once, then multiplied in float32.
  const float wi = (float)w[b], wt = (float)(1.0 - w[b]);
  float sc = 0.f;
  for (int e = lane; e < dim; e += 32) {
    const float c = wi * (ni > 0.f ? ip[e] / ni : 0.f) + wt * (nt > 0.f ? tp[e] / nt : 0.f);
    out[(size_t)b * dim + e] = c;
    sc = fmaf(c, c, sc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, off);
  const float nc = sqrtf(sc);
  __syncwarp();
  for (int e = lane; e < dim; e += 32) {
    const float c = out[(size_t)b * dim + e];
    out[(size_t)b * dim + e] = nc > 0.f ? c / nc : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// Store maintenance (Collection.delete, backend/app/main.py:1069; reset :1058-1098).
// move_rows: for every (src, dst) pair copy the stored row, its inverse norm and its filter bits
// src -> dst.  The host guarantees {src} and {dst} are disjoint (sources are surviving tail rows,
// destinations are holes below the new count), so ONE launch compacts any number of deletions.
// One warp per pair, 16-byte chunks.  `src_*` may be peer-mapped memory of another GPU (vs_copy_row).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) move_rows_kernel(const uint4* src_rows, const float* src_inv, const uint64_t* src_mask,
                                                        uint4* dst_rows, float* dst_inv, uint64_t* dst_mask,
                                                        const int64_t* __restrict__ pairs, int64_t n_pairs, int chunks) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w0; i < n_pairs; i += nw) {
    const int64_t s = pairs[2 * i], d = pairs[2 * i + 1];
    for (int c = lane; c < chunks; c += 32) dst_rows[d * chunks + c] = src_rows[s * chunks + c];
    if (lane == 0) dst_inv[d] = src_inv[s];
    if (dst_mask && lane < kMaskWords) dst_mask[d * kMaskWords + lane] = src_mask ? src_mask[s * kMaskWords + lane] : 0ull;
  }
}
cudaError_t launch_move_rows(const void* src_rows, const float* src_inv, const uint64_t* src_mask, void* dst_rows,
                             float* dst_inv, uint64_t* dst_mask, const int64_t* pairs_dev, int64_t n_pairs, int64_t ld_bytes,
                             cudaStream_t st) {
  if (n_pairs <= 0) return cudaSuccess;
  const int64_t blocks = min((int64_t)148 * 8, (n_pairs + 7) / 8);
  move_rows_kernel<<<(int)blocks, 256, 0, st>>>(static_cast<const uint4*>(src_rows), src_inv, src_mask,
                                                 static_cast<uint4*>(dst_rows), dst_inv, dst_mask, pairs_dev, n_pairs,
                                                 (int)(ld_bytes / 16));
  count_launch();
  return cudaGetLastError();
}

// dst row (dst_first + l * dst_stride) <- src row l for l in [0, n): raw storage (rows, inverse norms).
// Runs on the DESTINATION device and reads the source shard through NVLink peer memory: this is how the
// row-striped shards of a vs_group are replicated into one full index per GPU for the all-pairs pass.
__global__ void __launch_bounds__(256) strided_copy_kernel(const uint4* __restrict__ src_rows, const float* __restrict__ src_inv,
                                                           uint4* __restrict__ dst_rows, float* __restrict__ dst_inv, int64_t n,
                                                           int64_t dst_first, int64_t dst_stride, int chunks) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t l = w0; l < n; l += nw) {
    const int64_t d = dst_first + l * dst_stride;
    for (int c = lane; c < chunks; c += 32) dst_rows[d * chunks + c] = src_rows[l * chunks + c];
    if (lane == 0) dst_inv[d] = src_inv[l];
  }
}
cudaError_t launch_strided_copy(const void* src_rows, const float* src_inv, void* dst_rows, float* dst_inv, int64_t n,
                                int64_t dst_first, int64_t dst_stride, int64_t ld_bytes, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t blocks = min((int64_t)148 * 16, (n + 7) / 8);
  strided_copy_kernel<<<(int)blocks, 256, 0, st>>>(static_cast<const uint4*>(src_rows), src_inv, static_cast<uint4*>(dst_rows),
                                                    dst_inv, n, dst_first, dst_stride, (int)(ld_bytes / 16));
  count_launch();
  return cudaGetLastError();
}

// Filter sweep -> stored filter bits without leaving the device (the CLIP-side analogue of the answers
// written at backend/app/main.py:1010-1033): bit `bit` of row r := bit r of the sweep's word array.
__global__ void __launch_bounds__(256) apply_sweep_bits_kernel(const uint32_t* __restrict__ words, int64_t n, int bit,
                                                               uint64_t* __restrict__ mask) {
  const uint64_t m = 1ull << (bit & 63);
  const int w = bit >> 6;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const bool yes = (words[r >> 5] >> (r & 31)) & 1u;
    uint64_t* p = mask + r * kMaskWords + w;
    *p = yes ? (*p | m) : (*p & ~m);
  }
}
cudaError_t launch_apply_sweep_bits(const uint32_t* words, int64_t n, int bit, uint64_t* mask, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  apply_sweep_bits_kernel<<<(int)min((int64_t)148 * 8, (n + 255) / 256), 256, 0, st>>>(words, n, bit, mask);
  count_launch();
  return cudaGetLastError();
}

// gmin[g] = (1 - 2^-20) * min over the valid rows of 32-row group g of the row norm 1/inv_norm
// (+inf for a group with no valid row); the tcgen05 epilogues' fast-reject bound (gemm_tc.cu).
// Kept current by every mutation of the index (api.cu: refresh_gmin), so no query pays for it.
__global__ void __launch_bounds__(256) group_min_norm_kernel(const float* __restrict__ inv, uint32_t n_rows, uint32_t g_lo,
                                                             uint32_t g_hi, float* __restrict__ gmin) {
  const uint32_t g = g_lo + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= g_hi) return;
  const uint32_t row = g * 32 + (threadIdx.x & 31);
  float nrm = __int_as_float(0x7f800000);
  if (row < n_rows) nrm = 1.0f / __ldg(inv + row);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) nrm = fminf(nrm, __shfl_xor_sync(0xffffffffu, nrm, off));
  if ((threadIdx.x & 31) == 0) gmin[g] = nrm * (1.0f - 9.5367431640625e-07f);
}
cudaError_t launch_group_min(const float* inv, int64_t n_rows, int64_t g_lo, int64_t g_hi, float* gmin, cudaStream_t st) {
  if (g_hi <= g_lo) return cudaSuccess;
  group_min_norm_kernel<<<(unsigned)((g_hi - g_lo + 7) / 8), 256, 0, st>>>(inv, (uint32_t)n_rows, (uint32_t)g_lo, (uint32_t)g_hi, gmin);
  count_launch();
  return cudaGetLastError();
}

__global__ void fill_empty_kernel(float* s, int64_t* r, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    s[i] = VS_NEG_INF;
    r[i] = -1;
  }
}
cudaError_t launch_fill_empty(float* s, int64_t* r, int64_t n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  fill_empty_kernel<<<(int)min((int64_t)592, (n + 255) / 256), 256, 0, st>>>(s, r, n);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_blend(const float* img, const float* txt, const double* w, int B, int dim, float* out,
                         cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  blend_kernel<<<(B + 7) / 8, 256, 0, st>>>(img, txt, w, B, dim, out);
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K5: merge of per-shard candidates after the all-gather.  One CTA per query; candidates get a
// unique 64-bit key (orderable score << 32 | ~position); positions are (shard, rank) ordered,
// which for equal scores is the same as global-row order because shards are contiguous row
// ranges in rank order.  Block bitonic sort in shared memory, descending.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* keys, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t a = keys[lo], b = keys[hi];
        if ((a < b) == desc) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

// candidates: [G][Bstride][kin]; output [B][kout] (kout <= kin)
__global__ void __launch_bounds__(1024) merge_kernel(const float* __restrict__ cs, const int64_t* __restrict__ cr, int G,
                                                     int Bstride, int kin, int kout, int P, float* __restrict__ out_s,
                                                     int64_t* __restrict__ out_r, int out_stride) {
  extern __shared__ __align__(16) uint64_t keys[];
  const int b = blockIdx.x;
  const int C = G * kin;
  for (int c = threadIdx.x; c < P; c += blockDim.x) {
    uint64_t key = 0;
    if (c < C) {
      const int g = c / kin, e = c - g * kin;
      const size_t src = ((size_t)g * Bstride + b) * kin + e;
      const float s = cs[src];
      // ties in score rank by ascending ROW (lists may interleave row ranges), so the row itself is
      // the low half of the key; global rows are < 2^32 (documented collection limit)
      if (cr[src] >= 0 && s == s) key = ((uint64_t)score_key(s) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)cr[src]);
    }
    keys[c] = key;
  }
  bitonic_sort_desc(keys, P);
  for (int e = threadIdx.x; e < kout; e += blockDim.x) {
    const uint64_t key = e < P ? keys[e] : 0;
    float s = VS_NEG_INF;
    int64_t r = -1;
    if (key != 0) {
      s = key_score((uint32_t)(key >> 32));
      r = (int64_t)(0xFFFFFFFFu - (uint32_t)key);
    }
    out_s[(size_t)b * out_stride + e] = s;
    out_r[(size_t)b * out_stride + e] = r;
  }
}

cudaError_t launch_merge_ex(const float* cs, const int64_t* cr, int G, int Bstride, int B, int kin, int kout,
                            float* out_s, int64_t* out_r, cudaStream_t st, int out_stride) {
  if (G <= 0 || B <= 0 || kin <= 0 || kout <= 0 || kout > kin || Bstride < B) return cudaErrorInvalidValue;
  if (out_stride <= 0) out_stride = kout;
  int P = 2;
  while (P < G * kin) P <<= 1;
  if (P > 16384) return cudaErrorInvalidValue;
  const size_t smem = (size_t)P * 8;
  static bool attr_set[64] = {};      // per DEVICE: a single process may drive several GPUs (vs_group_t)
  int dev = 0;
  cudaError_t de = cudaGetDevice(&dev);
  if (de != cudaSuccess) return de;
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int threads = P / 2 < 1024 ? (P / 2 < 32 ? 32 : P / 2) : 1024;
  merge_kernel<<<B, threads, smem, st>>>(cs, cr, G, Bstride, kin, kout, P, out_s, out_r, out_stride);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_merge(const float* cs, const int64_t* cr, int G, int B, int k, float* out_s, int64_t* out_r,
                         cudaStream_t st) {
  return launch_merge_ex(cs, cr, G, B, B, k, k, out_s, out_r, st);
}

// ------------------------------------------------------------------------------------------
// The peer exchange as a kernel of its own (protocol: exchange.cuh), for candidates produced by
// the tcgen05 path or by batches larger than one fused scan launch.  One warp per query slot.
// The grid never exceeds the SM count, so every CTA is resident; each warp first pushes ALL of its
// queries and only then starts waiting, so no rank can wait on a push that has not been issued.
// ------------------------------------------------------------------------------------------
size_t exchange_bytes(int Bmax, int kmax, int G) { return 2 * xchg_layout(Bmax, kmax, G).half_bytes; }

constexpr int kXchgWarps = 8;

template <int M>
__device__ __forceinline__ void load_candidates(WarpTopK<M>& top, const float* cs, const int64_t* cr, int k, int lane) {
  top.init();
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const int p = m * 32 + lane;
    if (p < k && cr[p] >= 0) {
      top.s[m] = cs[p];
      top.r[m] = (uint32_t)cr[p];
    }
  }
}

// PUSH: candidates of this rank -> every peer's buffer + flags.  COLLECT: wait + merge + write.
template <int M, bool PUSH, bool COLLECT>
__global__ void __launch_bounds__(kXchgWarps * 32) exchange_merge_kernel(const float* __restrict__ cs,
                                                                         const int64_t* __restrict__ cr, XchgParams x, int B,
                                                                         int k, float* __restrict__ out_s,
                                                                         int64_t* __restrict__ out_r) {
  const int lane = threadIdx.x & 31;
  const int w0 = blockIdx.x * kXchgWarps + (threadIdx.x >> 5);
  const int nw = gridDim.x * kXchgWarps;
  WarpTopK<M> top;
  if constexpr (PUSH) {
    for (int b = w0; b < B; b += nw) {
      load_candidates(top, cs + (size_t)b * k, cr + (size_t)b * k, k, lane);
      xchg_push(x, top, x.slot0 + b, k, lane);
    }
  }
  if constexpr (COLLECT) {
    for (int b = w0; b < B; b += nw) {
      xchg_wait_merge(x, top, x.slot0 + b, k, lane);
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const int p = m * 32 + lane;
        if (p < k) {
          out_s[(size_t)b * k + p] = top.s[m];
          out_r[(size_t)b * k + p] = top.r[m] == kEmptyRow ? -1 : (int64_t)top.r[m];
        }
      }
    }
  }
}

// CUDA loads kernels lazily and a first launch may synchronise the context.  When several shards
// live in ONE process (tests, single-GPU multi-handle use) a peer's kernel may already be spinning
// on this shard's push at that moment, so everything the sharded path can launch besides the scan
// kernel itself is loaded up front (called by vs_exchange_create).
cudaError_t preload_exchange_kernels() {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, exchange_merge_kernel<1, true, true>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, exchange_merge_kernel<4, true, true>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, exchange_merge_kernel<1, true, false>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, exchange_merge_kernel<4, true, false>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, exchange_merge_kernel<1, false, true>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, exchange_merge_kernel<4, false, true>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, fill_empty_kernel);
  return e;
}

// what: 3 = push + collect (one exchange), 1 = push only, 2 = collect only
cudaError_t launch_exchange_merge(const float* cs, const int64_t* cr, const XchgParams& x, int B, int k, float* out_s,
                                  int64_t* out_r, int sm_count, cudaStream_t st, int what) {
  if (x.G <= 0 || x.G > kMaxPeers || x.rank < 0 || x.rank >= x.G || B <= 0 || x.slot0 < 0 || x.slot0 + B > x.Bmax ||
      k <= 0 || k > x.kmax || k > kMaxFusedK || what < 1 || what > 3)
    return cudaErrorInvalidValue;
  int grid = (B + kXchgWarps - 1) / kXchgWarps;
  if (grid > sm_count) grid = sm_count;
  const dim3 g(grid), b(kXchgWarps * 32);
  if (k <= 32) {
    if (what == 3) exchange_merge_kernel<1, true, true><<<g, b, 0, st>>>(cs, cr, x, B, k, out_s, out_r);
    else if (what == 1) exchange_merge_kernel<1, true, false><<<g, b, 0, st>>>(cs, cr, x, B, k, out_s, out_r);
    else exchange_merge_kernel<1, false, true><<<g, b, 0, st>>>(cs, cr, x, B, k, out_s, out_r);
  } else {
    if (what == 3) exchange_merge_kernel<4, true, true><<<g, b, 0, st>>>(cs, cr, x, B, k, out_s, out_r);
    else if (what == 1) exchange_merge_kernel<4, true, false><<<g, b, 0, st>>>(cs, cr, x, B, k, out_s, out_r);
    else exchange_merge_kernel<4, false, true><<<g, b, 0, st>>>(cs, cr, x, B, k, out_s, out_r);
  }
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Exact select for k > 128 (the UI's "All" = 1000, backend/app/main.py:757) over materialised
// scores.  Every row gets the unique key (orderable score << 32 | ~row); six MSB-first radix
// passes (11,11,10,11,11,10 bits) pin down the k-th largest key exactly -- ties in score are
// resolved by row index inside the key, so exactly k rows satisfy key >= threshold.  The last
// CTA of each pass (ticket) picks the digit; the last CTA of the compaction pass sorts.
// ------------------------------------------------------------------------------------------
constexpr int kSelBins = 2048;
struct SelectState {
  unsigned long long prefix;   // known high bits of the threshold key
  unsigned long long known;    // mask of known bits
  unsigned int k_rem;          // how many keys still to take from the matching set
  unsigned int ticket;
  unsigned int out_count;
  unsigned int k_eff;
  unsigned int hist[kSelBins];
  unsigned long long buf[kMaxK];
};
size_t select_workspace_bytes(int B) { return sizeof(SelectState) * (size_t)B; }

__device__ __forceinline__ unsigned long long select_key(float s, uint32_t row) {
  if (!(s == s)) s = VS_NEG_INF;
  return ((unsigned long long)score_key(s) << 32) | (unsigned long long)(0xFFFFFFFFu - row);
}

__global__ void __launch_bounds__(256) select_init_kernel(SelectState* st, int B, int k, int64_t n) {
  const int b = blockIdx.x;
  if (b >= B) return;
  SelectState* s = st + b;
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) s->hist[i] = 0;
  if (threadIdx.x == 0) {
    s->prefix = 0;
    s->known = 0;
    s->k_eff = (unsigned int)min((int64_t)k, n);
    s->k_rem = s->k_eff;
    s->ticket = 0;
    s->out_count = 0;
  }
}

__global__ void __launch_bounds__(256) select_hist_kernel(const float* __restrict__ scores, int64_t n, SelectState* st,
                                                          int shift, int bits) {
  __shared__ unsigned int sh[kSelBins];
  __shared__ int s_last;
  SelectState* s = st + blockIdx.y;
  const float* sc = scores + (size_t)blockIdx.y * n;
  const int nb = 1 << bits;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const unsigned long long prefix = s->prefix, known = s->known;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = select_key(sc[i], (uint32_t)i);
    if (((key ^ prefix) & known) == 0) atomicAdd(&sh[(unsigned int)(key >> shift) & (nb - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x)
    if (sh[i]) atomicAdd(&s->hist[i], sh[i]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&s->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // pick the digit: walk bins from the top until the cumulative count reaches k_rem.
  // (2048 bins, one thread -- a few microseconds, once per pass.)
  if (threadIdx.x == 0) {
    volatile unsigned int* h = s->hist;
    unsigned int rem = s->k_rem;
    int d = nb - 1;
    for (; d > 0; --d) {
      const unsigned int c = h[d];
      if (c >= rem) break;
      rem -= c;
    }
    s->k_rem = rem;
    s->prefix = prefix | ((unsigned long long)d << shift);
    s->known = known | ((unsigned long long)(nb - 1) << shift);
    s->ticket = 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) s->hist[i] = 0;
}

__global__ void __launch_bounds__(256) select_compact_kernel(const float* __restrict__ scores, int64_t n, SelectState* st,
                                                             int k, int64_t row_base, int64_t row_stride, float* __restrict__ out_s,
                                                             int64_t* __restrict__ out_r) {
  __shared__ int s_last;
  __shared__ unsigned long long keys[kMaxK];
  SelectState* s = st + blockIdx.y;
  const float* sc = scores + (size_t)blockIdx.y * n;
  const unsigned long long thr = s->prefix;
  const unsigned int k_eff = s->k_eff;
  if (k_eff > 0) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const unsigned long long key = select_key(sc[i], (uint32_t)i);
      if (key >= thr) {
        const unsigned int slot = atomicAdd(&s->out_count, 1u);
        if (slot < (unsigned int)kMaxK) s->buf[slot] = key;
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&s->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  volatile unsigned long long* gb = s->buf;
  for (int i = threadIdx.x; i < kMaxK; i += blockDim.x) keys[i] = i < (int)k_eff ? gb[i] : 0ull;
  bitonic_sort_desc(reinterpret_cast<uint64_t*>(keys), kMaxK);
  for (int e = threadIdx.x; e < k; e += blockDim.x) {
    float v = VS_NEG_INF;
    int64_t r = -1;
    if (e < (int)k_eff) {
      const unsigned long long key = keys[e];
      v = key_score((uint32_t)(key >> 32));
      if (v != VS_NEG_INF) r = (int64_t)(0xFFFFFFFFu - (uint32_t)key) * row_stride + row_base;
    }
    out_s[(size_t)blockIdx.y * k + e] = v;
    out_r[(size_t)blockIdx.y * k + e] = r;
  }
  if (threadIdx.x == 0) {
    s->ticket = 0;
    s->out_count = 0;
  }
}

cudaError_t launch_select(const float* scores, int64_t n, int B, int k, int64_t row_base, int64_t row_stride, void* workspace,
                          float* out_s, int64_t* out_r, cudaStream_t st) {
  if (k > kMaxK || k <= 0 || n <= 0 || n > 0xFFFFFFF0LL) return cudaErrorInvalidValue;
  SelectState* ss = static_cast<SelectState*>(workspace);
  select_init_kernel<<<B, 256, 0, st>>>(ss, B, k, n);
  int gx = (int)min((int64_t)148 * 4, (n + 255) / 256);
  if (gx < 1) gx = 1;
  const int shifts[6] = {53, 42, 32, 21, 10, 0};
  const int bits[6] = {11, 11, 10, 11, 11, 10};
  for (int pass = 0; pass < 6; ++pass) select_hist_kernel<<<dim3(gx, B), 256, 0, st>>>(scores, n, ss, shifts[pass], bits[pass]);
  select_compact_kernel<<<dim3(gx, B), 256, 0, st>>>(scores, n, ss, k, row_base, row_stride, out_s, out_r);
  count_launch(8);
  return cudaGetLastError();
}

}  // namespace vs
