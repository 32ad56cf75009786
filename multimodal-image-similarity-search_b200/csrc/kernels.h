// kernels.h -- host-callable launchers of the sm_100a kernels (internal to the library).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "exchange.cuh"

namespace vs {

constexpr int kMaskWords = 4;
constexpr int kMaxFusedK = 128;   // register-list top-k limit of the fused scan kernel
constexpr int kMaxTensorK = 128;  // tcgen05 top-k: per-thread register lists of 32, ceil(k / 32) exact rounds
constexpr int kMaxK = 1024;
constexpr int kScanMaxCtasPerSm = 4;   // the partial-list workspace is sized for this many CTAs per SM

void count_launch(int n = 1);

// ---- K6 ingest: f32 rows -> storage dtype + inverse norms -------------------------------
cudaError_t launch_ingest(const float* src, int64_t n, int dim, int dtype, void* dst_rows, int64_t ld_elems,
                          float* dst_inv, cudaStream_t st);
// inverse norms of rows [first, first+n) that are already stored (slab reload)
cudaError_t launch_renorm(const void* rows, int64_t first, int64_t n, int dim, int dtype, int64_t ld_elems, float* inv,
                          cudaStream_t st);
// storage -> f32 (Collection.get(include=["embeddings"]), persistence)
cudaError_t launch_export(const void* rows, int64_t n, int dim, int dtype, int64_t ld_elems, float* dst,
                          cudaStream_t st);
// multimodal blend (backend/app/main.py:850-860) for B triples
cudaError_t launch_blend(const float* img, const float* txt, const double* w, int B, int dim, float* out,
                         cudaStream_t st);

// store maintenance: batched row moves (compaction after deletes), strided replication, sweep bits, group bounds
cudaError_t launch_move_rows(const void* src_rows, const float* src_inv, const uint64_t* src_mask, void* dst_rows,
                             float* dst_inv, uint64_t* dst_mask, const int64_t* pairs_dev, int64_t n_pairs, int64_t ld_bytes,
                             cudaStream_t st);
cudaError_t launch_strided_copy(const void* src_rows, const float* src_inv, void* dst_rows, float* dst_inv, int64_t n,
                                int64_t dst_first, int64_t dst_stride, int64_t ld_bytes, cudaStream_t st);
cudaError_t launch_apply_sweep_bits(const uint32_t* words, int64_t n, int bit, uint64_t* mask, cudaStream_t st);
// gmin[g] for 32-row groups [g_lo, g_hi): (1 - 2^-20) * min row norm over rows < n_rows (+inf if none)
cudaError_t launch_group_min(const float* inv, int64_t n_rows, int64_t g_lo, int64_t g_hi, float* gmin, cudaStream_t st);

// empty result slots: score = -inf, row = -1
cudaError_t launch_fill_empty(float* s, int64_t* r, int64_t n, cudaStream_t st);

// ---- K1 scan: fused normalise + GEMV + top-k (HBM-bound) ---------------------------------
struct ScanArgs {
  const void* rows;
  const float* inv_norm;
  const uint64_t* mask;      // [n][kMaskWords] or nullptr
  uint64_t req[kMaskWords];  // required bits (all zero = no filter)
  const float* q;            // [B][dim] raw f32 queries (device)
  int B, dim, dtype, k;
  int64_t ld_bytes;          // row pitch in bytes (multiple of 16)
  int64_t n_rows;
  int64_t row_base;
  int64_t row_stride = 1;    // reported row = row_base + local row * row_stride (striped shards)
  // workspace (device): partial lists [B][grid][k], tickets [B]
  float* part_s;
  uint32_t* part_r;
  unsigned int* tickets;
  int grid_x;                // CTAs per query (<= SM count); workspace is sized for it
  // outputs (device)
  float* out_s;              // [B][k]
  int64_t* out_r;            // [B][k]
  float* scores_full;        // [B][n_rows] when materialising for the large-k path, else nullptr
  XchgParams xg = {};        // xg.G > 0: fused peer exchange of the result (B <= 64, k <= 128)
  int early_wait = 1;        // 0 only when the previous launch on the stream was a scan and q / the rows are older than it
  unsigned int* done_flag = nullptr;   // host-mapped completion flags [B] (request/response without a stream sync)
  unsigned int done_seq = 0;
  const float* q_host = nullptr;       // B == 1, dim <= kMaxInlineQ: the query (HOST memory) rides in the kernel parameters
};
constexpr int kMaxInlineQ = 1024;
// returns cudaErrorInvalidValue when (dtype, ld_bytes) has no instantiation
cudaError_t launch_scan(const ScanArgs& a, int sm_count, cudaStream_t st);
int scan_rows_per_tile(int dtype, int64_t ld_bytes);

// ---- K5 merge / select --------------------------------------------------------------------
// [G][B][k] candidates (global rows, <0 empty) -> [B][k]
cudaError_t launch_merge(const float* cs, const int64_t* cr, int G, int B, int k, float* out_s, int64_t* out_r,
                         cudaStream_t st);
// the exchange kernel alone: [B][k] candidates (global rows, <0 empty) of this rank -> global [B][k]
size_t exchange_bytes(int Bmax, int kmax, int G);
cudaError_t preload_exchange_kernels();
// what: 3 = push + wait + merge (one exchange), 1 = push only, 2 = wait + merge only (deferred collect)
cudaError_t launch_exchange_merge(const float* cs, const int64_t* cr, const XchgParams& x, int B, int k, float* out_s,
                                  int64_t* out_r, int sm_count, cudaStream_t st, int what = 3);
// general form: candidates [G][Bstride][kin] -> out[b * out_stride + 0..kout) (out_stride <= 0: kout)
cudaError_t launch_merge_ex(const float* cs, const int64_t* cr, int G, int Bstride, int B, int kin, int kout,
                            float* out_s, int64_t* out_r, cudaStream_t st, int out_stride = 0);
// exact top-k (k <= 1024) of materialised scores [B][n] -> [B][k]; workspace: see select_workspace_bytes
size_t select_workspace_bytes(int B);
cudaError_t launch_select(const float* scores, int64_t n, int B, int k, int64_t row_base, int64_t row_stride, void* workspace,
                          float* out_s, int64_t* out_r, cudaStream_t st);

// ---- K2/K3/K4 tcgen05 kernels (bf16 storage) -------------------------------------------------
struct TensorArgs {
  const void* rows;          // bf16 [n][ld]
  const float* inv_norm;
  const float* gmin;         // [8 * ceil(n / 256)] group bounds (launch_group_min), maintained by the index
  const uint64_t* mask;
  uint64_t req[kMaskWords];
  int dim;
  int64_t ld_elems;
  int64_t n_rows;
  int64_t row_base;
  int64_t row_stride = 1;
};
// queries: f32 [B][dim] raw (normalised + rounded to bf16 on device into q_bf16 workspace)
size_t tensor_workspace_bytes(int B, int dim, int k, int sm_count);
cudaError_t launch_tensor_topk(const TensorArgs& a, const float* q, int B, int k, void* workspace, float* out_s,
                               int64_t* out_r, int sm_count, cudaStream_t st);
cudaError_t launch_tensor_filter(const TensorArgs& a, const float* prompts, int F, float tau, void* workspace,
                                 uint32_t* out_bits, int64_t words_per_filter, int sm_count, cudaStream_t st);
cudaError_t launch_tensor_dedup(const TensorArgs& a, int64_t row_lo, int64_t row_hi, float tau, int64_t cap,
                                int64_t* out_i, int64_t* out_j, float* out_score, unsigned long long* out_count,
                                void* workspace, int sm_count, cudaStream_t st);
bool tensor_dim_ok(int dim);   // dim % 8 == 0, dim <= 4096 (A block resident up to 512, streamed beyond)

}  // namespace vs
