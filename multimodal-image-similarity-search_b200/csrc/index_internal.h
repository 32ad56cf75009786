// index_internal.h -- the object behind vs_index_t and the helpers api.cu and group.cu share.
// Internal to the library: nothing here is part of the C ABI (include/vecsearch_b200.h).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "../../include/vecsearch_b200.h"
#include "kernels.h"

namespace vs {

int fail(int code, const char* fmt, ...);   // sets the calling thread's vs_last_error() text, returns code

#define VS_CU(call)                                                                                          \
  do {                                                                                                       \
    cudaError_t e__ = (call);                                                                                \
    if (e__ != cudaSuccess) {                                                                                \
      cudaGetLastError();                                                                                    \
      return vs::fail(e__ == cudaErrorMemoryAllocation ? VS_ERR_OOM : VS_ERR_CUDA, "%s: %s (%s:%d)", #call,  \
                      cudaGetErrorString(e__), __FILE__, __LINE__);                                          \
    }                                                                                                        \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    size_t want = need + need / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return e;
    bytes = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};
struct HostBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
    size_t want = need + need / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) return e;
    bytes = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
  }
};

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
    ok = (cudaSetDevice(dev) == cudaSuccess);
    if (!ok) cudaGetLastError();
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// queries per scan launch (bounds the partial-list workspace)
constexpr int kScanBatch = 64;

}  // namespace vs

struct vs_index {
  int device = 0;
  int dim = 0;
  int dtype = 0;
  int esize = 4;
  int64_t ld = 0;        // row pitch in elements
  int64_t n = 0;
  int64_t cap = 0;
  int64_t row_base = 0;
  int64_t row_stride = 1;
  int sm_count = 148;
  int last_path = 0;
  void* rows = nullptr;
  float* inv = nullptr;
  float* gmin = nullptr;      // bf16 only: per 32-row group (1 - 2^-20) * min row norm, kept current by every mutation
  uint64_t* mask = nullptr;   // lazily allocated [cap][4]
  cudaStream_t stream = nullptr;
  cudaEvent_t ev = nullptr;
  cudaStream_t last_stream = nullptr;
  bool last_stream_valid = false;
  // PDL chain: true while the last kernel this library enqueued on `chain_stream` was a scan launch and
  // no row / inverse norm / filter bit has been modified since (see ScanArgs::early_wait)
  bool chain_ok = false;
  cudaStream_t chain_stream = nullptr;
  std::mutex mu;
  // scratch
  vs::DevBuf d_q, d_out_s, d_out_r, d_part_s, d_part_r, d_tickets, d_scores, d_select, d_tensor, d_stage, d_misc;
  vs::HostBuf h_in, h_out;
  size_t tickets_n = 0;
  // peer exchange (row-sharded collection): local buffer + the peers' mappings
  struct Exchange {
    int G = 0, rank = 0, Bmax = 0, kmax = 0;
    size_t bytes = 0;
    void* local = nullptr;
    void* peers[vs::kMaxPeers] = {};
    bool ipc_opened[vs::kMaxPeers] = {};
    bool attached = false;
    uint32_t epoch = 0;
    unsigned int* h_err = nullptr;   // sticky error word: host-mapped pinned memory, written by the kernels
  } xc;
  vs::DevBuf d_xs, d_xr;
};

namespace vs {

// Optional behaviours of one query call.
struct QueryOpts {
  // in: ask for the peer exchange to be fused into the scan kernel (row-sharded collection).  push_slot0 >= 0:
  // only PUSH into slots [push_slot0, +B) of the CURRENT epoch; else a whole new exchange (push + wait + merge).
  bool want_fused = false;
  int push_slot0 = -1;
  // in: host-mapped completion flags [B]; armed only on the single-launch fused scan path
  unsigned int* done_flag = nullptr;
  unsigned int done_seq = 0;
  // in: B == 1 on the fused scan path: the query in HOST memory, shipped inside the launch packet (no H2D copy)
  const float* q_host = nullptr;
  // out
  bool fused = false;        // the exchange was done by the scan kernel
  bool done_armed = false;   // the kernel will raise done_flag[0..B)
};

// The bodies behind the extern "C" entry points; the caller holds ix->mu and has selected ix->device.
int grow_locked(vs_index* ix, int64_t need_rows);
int add_dev_locked(vs_index* ix, const float* rows_dev, int64_t n, int64_t* first_row, cudaStream_t st);
int query_dev_locked(vs_index* ix, const float* q_dev, int B, int k, const uint64_t* req, int mode, float* out_s,
                     int64_t* out_r, cudaStream_t st, QueryOpts* opts = nullptr);
int sharded_query_dev_locked(vs_index* ix, const float* q_dev, int B, int k, const uint64_t* req, int mode, float* out_s,
                             int64_t* out_r, cudaStream_t st, QueryOpts* opts = nullptr);
int exchange_create_locked(vs_index* ix, int G, int rank, int B_max, int k_max);
int exchange_error_locked(vs_index* ix, bool clear);   // VS_OK or VS_ERR_EXCHANGE (message set)
inline cudaStream_t pick_stream(vs_index* ix, void* stream) { return stream ? (cudaStream_t)stream : ix->stream; }
inline void touch(vs_index* ix) { ix->chain_ok = false; }
cudaError_t order_after_last(vs_index* ix, cudaStream_t st);
cudaError_t refresh_gmin(vs_index* ix, int64_t row_lo, int64_t row_hi, cudaStream_t st);

}  // namespace vs
