// api.cu -- the C ABI declared in include/vecsearch_b200.h: index handle, slab management,
// host<->device staging, path selection.  No compute happens on the host.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/vecsearch_b200.h"
#include "kernels.h"

namespace vs {
static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
}  // namespace vs

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      cudaGetLastError();                                                                          \
      return fail(e__ == cudaErrorMemoryAllocation ? VS_ERR_OOM : VS_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                                    \
    }                                                                                              \
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
  uint64_t* mask = nullptr;   // lazily allocated [cap][4]
  bool any_mask = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev = nullptr;
  cudaStream_t last_stream = nullptr;
  bool last_stream_valid = false;
  std::mutex mu;
  // scratch
  DevBuf d_q, d_out_s, d_out_r, d_part_s, d_part_r, d_tickets, d_scores, d_select, d_tensor, d_stage, d_misc, d_err;
  HostBuf h_in, h_out;
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
  } xc;
  DevBuf d_xs, d_xr;
};

namespace {

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

int64_t pitch_elems(int dim, int esize) {
  const int per16 = 16 / esize;
  return ((int64_t)dim + per16 - 1) / per16 * per16;
}

int grow(vs_index* ix, int64_t need_rows) {
  if (need_rows <= ix->cap) return VS_OK;
  int64_t ncap = ix->cap > 0 ? ix->cap : 1024;
  while (ncap < need_rows) ncap += ncap < (1 << 20) ? ncap : ncap / 2;
  const size_t row_bytes = (size_t)ix->ld * ix->esize;
  void* nrows = nullptr;
  float* ninv = nullptr;
  // +64 floats: the scan's bulk copy of inverse norms rounds the tail up to 4 entries
  cudaError_t e = cudaMalloc(&nrows, (size_t)ncap * row_bytes + 256);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(VS_ERR_OOM, "cudaMalloc(%zu) for %lld rows failed: %s", (size_t)ncap * row_bytes, (long long)ncap,
                cudaGetErrorString(e));
  }
  e = cudaMalloc((void**)&ninv, ((size_t)ncap + 512) * sizeof(float));
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaFree(nrows);
    return fail(VS_ERR_OOM, "cudaMalloc inverse norms failed: %s", cudaGetErrorString(e));
  }
  CU(cudaMemsetAsync(ninv, 0, ((size_t)ncap + 512) * sizeof(float), ix->stream));
  uint64_t* nmask = nullptr;
  if (ix->mask) {
    e = cudaMalloc((void**)&nmask, (size_t)ncap * vs::kMaskWords * 8);
    if (e != cudaSuccess) {
      cudaGetLastError();
      cudaFree(nrows);
      cudaFree(ninv);
      return fail(VS_ERR_OOM, "cudaMalloc filter bits failed: %s", cudaGetErrorString(e));
    }
    CU(cudaMemsetAsync(nmask, 0, (size_t)ncap * vs::kMaskWords * 8, ix->stream));
  }
  if (ix->n > 0) {
    CU(cudaMemcpyAsync(nrows, ix->rows, (size_t)ix->n * row_bytes, cudaMemcpyDeviceToDevice, ix->stream));
    CU(cudaMemcpyAsync(ninv, ix->inv, (size_t)ix->n * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
    if (nmask)
      CU(cudaMemcpyAsync(nmask, ix->mask, (size_t)ix->n * vs::kMaskWords * 8, cudaMemcpyDeviceToDevice, ix->stream));
  }
  // rare path: wait for everything (adds/queries may be in flight on caller streams)
  CU(cudaDeviceSynchronize());
  if (ix->rows) cudaFree(ix->rows);
  if (ix->inv) cudaFree(ix->inv);
  if (ix->mask) cudaFree(ix->mask);
  ix->rows = nrows;
  ix->inv = ninv;
  ix->mask = nmask;
  ix->cap = ncap;
  return VS_OK;
}

int ensure_mask(vs_index* ix) {
  if (ix->mask) return VS_OK;
  if (ix->cap == 0) {
    int rc = grow(ix, 1);
    if (rc) return rc;
  }
  CU(cudaMalloc((void**)&ix->mask, (size_t)ix->cap * vs::kMaskWords * 8));
  CU(cudaMemsetAsync(ix->mask, 0, (size_t)ix->cap * vs::kMaskWords * 8, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

cudaStream_t pick_stream(vs_index* ix, void* stream) { return stream ? (cudaStream_t)stream : ix->stream; }

// queries per scan launch (bounds the partial-list workspace)
constexpr int kScanBatch = 64;

// kernel-side exchange descriptor for the CURRENT epoch; `advance` starts a new exchange first
vs::XchgParams make_exchange(vs_index* ix, int slot0, bool advance, bool push_only) {
  if (advance && ++ix->xc.epoch == 0) ix->xc.epoch = 2;   // 0 is the "never written" flag value; keep the parity sequence
  vs::XchgParams x = {};
  for (int g = 0; g < ix->xc.G; ++g) x.peers[g] = static_cast<unsigned char*>(ix->xc.peers[g]);
  x.err = (unsigned int*)ix->d_err.p;
  x.G = ix->xc.G;
  x.rank = ix->xc.rank;
  x.Bmax = ix->xc.Bmax;
  x.kmax = ix->xc.kmax;
  x.slot0 = slot0;
  x.epoch = ix->xc.epoch;
  x.push_only = push_only ? 1 : 0;
  return x;
}

// `fused_exchange`: when non-null and the query takes the fused scan path in ONE launch, the
// exchange (or, with push_slot0 >= 0, only the push into slots [push_slot0, +B) of the current
// epoch) is done by the scan kernel itself and *fused_exchange is set to true.
int query_dev_locked(vs_index* ix, const float* q_dev, int B, int k, const uint64_t* req, int mode, float* out_s,
                     int64_t* out_r, cudaStream_t st, bool* fused_exchange = nullptr, int push_slot0 = -1) {
  if (fused_exchange) *fused_exchange = false;
  if (B <= 0) return VS_OK;
  if (k <= 0 || k > vs::kMaxK) return fail(VS_ERR_ARG, "k=%d out of range [1,%d]", k, vs::kMaxK);
  // the scratch buffers (partial lists, tickets, select state) are shared by all queries of this
  // index: serialise against the stream that used them last
  if (ix->last_stream_valid && ix->last_stream != st) {
    CU(cudaEventRecord(ix->ev, ix->last_stream));
    CU(cudaStreamWaitEvent(st, ix->ev, 0));
  }
  ix->last_stream = st;
  ix->last_stream_valid = true;
  if (ix->n == 0) {
    // empty collection: all slots empty
    CU(vs::launch_fill_empty(out_s, out_r, (int64_t)B * k, st));
    return VS_OK;
  }
  bool use_mask = false;
  uint64_t reqw[vs::kMaskWords] = {0, 0, 0, 0};
  if (req)
    for (int w = 0; w < vs::kMaskWords; ++w) {
      reqw[w] = req[w];
      if (req[w]) use_mask = true;
    }
  if (use_mask && !ix->mask) {
    // filter requested but no row carries any bit: nothing can match
    CU(vs::launch_fill_empty(out_s, out_r, (int64_t)B * k, st));
    return VS_OK;
  }
  int path = mode;
  if (path == VS_Q_AUTO)
    path = (ix->dtype == VS_BF16 && B >= 16 && k <= vs::kMaxTensorK && vs::tensor_dim_ok(ix->dim) &&
            vs::tensor_path_available()) ? VS_Q_TENSOR : VS_Q_SCAN;
  if (path == VS_Q_TENSOR) {
    if (ix->dtype != VS_BF16) return fail(VS_ERR_UNSUPPORTED, "tensor path needs bf16 storage");
    if (k > vs::kMaxTensorK) return fail(VS_ERR_UNSUPPORTED, "tensor path supports k <= %d", vs::kMaxTensorK);
    if (!vs::tensor_path_available()) return fail(VS_ERR_UNSUPPORTED, "tensor path not built");
    vs::TensorArgs ta;
    ta.rows = ix->rows;
    ta.inv_norm = ix->inv;
    ta.mask = use_mask ? ix->mask : nullptr;
    memcpy(ta.req, reqw, sizeof(reqw));
    ta.dim = ix->dim;
    ta.ld_elems = ix->ld;
    ta.n_rows = ix->n;
    ta.row_base = ix->row_base;
    ta.row_stride = ix->row_stride;
    CU(ix->d_tensor.reserve(vs::tensor_workspace_bytes(B, ix->dim, k, ix->sm_count, ix->n)));
    CU(vs::launch_tensor_topk(ta, q_dev, B, k, ix->d_tensor.p, out_s, out_r, ix->sm_count, st));
    ix->last_path = VS_Q_TENSOR;
    return VS_OK;
  }
  // ---- scan path ----
  const int64_t ld_bytes = ix->ld * ix->esize;
  if (vs::scan_rows_per_tile(ix->dtype, ld_bytes) < 0)
    return fail(VS_ERR_UNSUPPORTED, "row pitch %lld bytes exceeds the scan kernel's 4096-byte limit", (long long)ld_bytes);
  const bool large_k = k > vs::kMaxFusedK;
  const int step = large_k ? 4 : kScanBatch;
  if (!large_k) {
    CU(ix->d_part_s.reserve((size_t)step * ix->sm_count * vs::kScanMaxCtasPerSm * k * sizeof(float)));
    CU(ix->d_part_r.reserve((size_t)step * ix->sm_count * vs::kScanMaxCtasPerSm * k * sizeof(uint32_t)));
  } else {
    CU(ix->d_scores.reserve((size_t)step * ix->n * sizeof(float)));
    CU(ix->d_select.reserve(vs::select_workspace_bytes(step)));
  }
  if (ix->tickets_n < (size_t)kScanBatch) {
    CU(ix->d_tickets.reserve(kScanBatch * sizeof(unsigned int)));
    CU(cudaMemsetAsync(ix->d_tickets.p, 0, ix->d_tickets.bytes, st));
    ix->tickets_n = kScanBatch;
  }
  for (int b0 = 0; b0 < B; b0 += step) {
    const int nb = B - b0 < step ? B - b0 : step;
    vs::ScanArgs a;
    a.rows = ix->rows;
    a.inv_norm = ix->inv;
    a.mask = use_mask ? ix->mask : nullptr;
    memcpy(a.req, reqw, sizeof(reqw));
    a.q = q_dev + (size_t)b0 * ix->dim;
    a.B = nb;
    a.dim = ix->dim;
    a.dtype = ix->dtype;
    a.k = k;
    a.ld_bytes = ld_bytes;
    a.n_rows = ix->n;
    a.row_base = ix->row_base;
    a.row_stride = ix->row_stride;
    a.part_s = (float*)ix->d_part_s.p;
    a.part_r = (uint32_t*)ix->d_part_r.p;
    a.tickets = (unsigned int*)ix->d_tickets.p;
    a.grid_x = ix->sm_count;
    a.out_s = out_s + (size_t)b0 * k;
    a.out_r = out_r + (size_t)b0 * k;
    a.scores_full = large_k ? (float*)ix->d_scores.p : nullptr;
    if (fused_exchange && !large_k && B <= step) {
      a.xg = push_slot0 >= 0 ? make_exchange(ix, push_slot0, false, true) : make_exchange(ix, 0, true, false);
      *fused_exchange = true;
    }
    CU(vs::launch_scan(a, ix->sm_count, st));
    if (large_k)
      CU(vs::launch_select((const float*)ix->d_scores.p, ix->n, nb, k, ix->row_base, ix->row_stride, ix->d_select.p, a.out_s, a.out_r, st));
  }
  ix->last_path = VS_Q_SCAN;
  return VS_OK;
}

}  // namespace

extern "C" {

const char* vs_last_error(void) { return g_err; }
int vs_abi_version(void) { return 1; }
uint64_t vs_launch_count(void) { return vs::g_launches.load(); }

int vs_create(int device, int dim, int dtype, int64_t capacity_rows, vs_index_t** out) {
  if (!out) return fail(VS_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (dim <= 0 || dim > 65536) return fail(VS_ERR_ARG, "dim=%d out of range", dim);
  if (dtype != VS_F32 && dtype != VS_BF16) return fail(VS_ERR_ARG, "dtype must be VS_F32 or VS_BF16");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(VS_ERR_CUDA, "no CUDA device (%s); this engine has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= ndev) return fail(VS_ERR_ARG, "device %d not in [0,%d)", device, ndev);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(VS_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  DeviceGuard g(device);
  if (!g.ok) return fail(VS_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  vs_index* ix = new (std::nothrow) vs_index();
  if (!ix) return fail(VS_ERR_OOM, "host allocation failed");
  ix->device = device;
  ix->dim = dim;
  ix->dtype = dtype;
  ix->esize = dtype == VS_F32 ? 4 : 2;
  ix->ld = pitch_elems(dim, ix->esize);
  ix->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ix->ev, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    delete ix;
    return fail(VS_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  if (capacity_rows > 0) {
    int rc = grow(ix, capacity_rows);
    if (rc) {
      cudaStreamDestroy(ix->stream);
      delete ix;
      return rc;
    }
  }
  *out = ix;
  return VS_OK;
}

int vs_destroy(vs_index_t* ix) {
  if (!ix) return VS_OK;
  DeviceGuard g(ix->device);
  cudaStreamSynchronize(ix->stream);
  if (ix->rows) cudaFree(ix->rows);
  if (ix->inv) cudaFree(ix->inv);
  if (ix->mask) cudaFree(ix->mask);
  DevBuf* bufs[] = {&ix->d_q,       &ix->d_out_s,  &ix->d_out_r,  &ix->d_part_s, &ix->d_part_r, &ix->d_tickets,
                    &ix->d_scores,  &ix->d_select, &ix->d_tensor, &ix->d_stage,  &ix->d_misc,   &ix->d_err};
  for (DevBuf* b : bufs) b->release();
  ix->h_in.release();
  ix->h_out.release();
  for (int g = 0; g < vs::kMaxPeers; ++g)
    if (ix->xc.ipc_opened[g]) cudaIpcCloseMemHandle(ix->xc.peers[g]);
  if (ix->xc.local) cudaFree(ix->xc.local);
  ix->d_xs.release();
  ix->d_xr.release();
  if (ix->ev) cudaEventDestroy(ix->ev);
  cudaStreamDestroy(ix->stream);
  cudaGetLastError();
  delete ix;
  return VS_OK;
}

int64_t vs_count(const vs_index_t* ix) { return ix ? ix->n : 0; }
int vs_dim(const vs_index_t* ix) { return ix ? ix->dim : 0; }
int vs_dtype(const vs_index_t* ix) { return ix ? ix->dtype : -1; }
int vs_last_query_path(const vs_index_t* ix) { return ix ? ix->last_path : 0; }
int vs_device_sm_count(const vs_index_t* ix) { return ix ? ix->sm_count : 0; }

int vs_set_row_base(vs_index_t* ix, int64_t row_base) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  std::lock_guard<std::mutex> lk(ix->mu);
  ix->row_base = row_base;
  return VS_OK;
}

int vs_set_row_map(vs_index_t* ix, int64_t row_base, int64_t row_stride) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (row_base < 0 || row_stride < 1 || row_stride > 65536) return fail(VS_ERR_ARG, "need row_base >= 0 and 1 <= row_stride <= 65536");
  std::lock_guard<std::mutex> lk(ix->mu);
  ix->row_base = row_base;
  ix->row_stride = row_stride;
  return VS_OK;
}

int vs_add_dev(vs_index_t* ix, const float* rows_dev, int64_t n, int64_t* first_row, void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (n < 0 || (n > 0 && !rows_dev)) return fail(VS_ERR_ARG, "bad rows/n");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  if (first_row) *first_row = ix->n;
  if (n == 0) return VS_OK;
  if (ix->n + n > 0xFFFFFFF0LL) return fail(VS_ERR_ARG, "shard would exceed 2^32 rows");
  int rc = grow(ix, ix->n + n);
  if (rc) return rc;
  cudaStream_t st = pick_stream(ix, stream);
  char* dst = (char*)ix->rows + (size_t)ix->n * ix->ld * ix->esize;
  CU(vs::launch_ingest(rows_dev, n, ix->dim, ix->dtype, dst, ix->ld, ix->inv + ix->n, st));
  if (ix->mask) CU(cudaMemsetAsync(ix->mask + (size_t)ix->n * vs::kMaskWords, 0, (size_t)n * vs::kMaskWords * 8, st));
  if (st != ix->stream) {
    // later work on the index's own stream (host-buffer queries) must see these rows
    CU(cudaEventRecord(ix->ev, st));
    CU(cudaStreamWaitEvent(ix->stream, ix->ev, 0));
  }
  ix->n += n;
  return VS_OK;
}

int vs_add_host(vs_index_t* ix, const float* rows, int64_t n, int64_t* first_row) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (n < 0 || (n > 0 && !rows)) return fail(VS_ERR_ARG, "bad rows/n");
  if (n == 0) {
    if (first_row) *first_row = ix->n;
    return VS_OK;
  }
  // stage through a device buffer in chunks of <= 64 MiB
  const int64_t chunk_rows = (64LL << 20) / ((int64_t)ix->dim * 4) > 0 ? (64LL << 20) / ((int64_t)ix->dim * 4) : 1;
  int64_t first = -1;
  for (int64_t r0 = 0; r0 < n; r0 += chunk_rows) {
    const int64_t nr = n - r0 < chunk_rows ? n - r0 : chunk_rows;
    {
      std::lock_guard<std::mutex> lk(ix->mu);
      DeviceGuard g(ix->device);
      CU(ix->d_stage.reserve((size_t)nr * ix->dim * 4));
      CU(cudaMemcpyAsync(ix->d_stage.p, rows + (size_t)r0 * ix->dim, (size_t)nr * ix->dim * 4, cudaMemcpyHostToDevice,
                         ix->stream));
    }
    int64_t fr = 0;
    int rc = vs_add_dev(ix, (const float*)ix->d_stage.p, nr, &fr, nullptr);
    if (rc) return rc;
    {
      DeviceGuard g(ix->device);
      CU(cudaStreamSynchronize(ix->stream));
    }
    if (first < 0) first = fr;
  }
  if (first_row) *first_row = first;
  return VS_OK;
}

int vs_remove(vs_index_t* ix, int64_t row, int64_t* moved_from) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (row < 0 || row >= ix->n) return fail(VS_ERR_ARG, "row %lld out of range [0,%lld)", (long long)row, (long long)ix->n);
  DeviceGuard g(ix->device);
  const int64_t last = ix->n - 1;
  if (moved_from) *moved_from = row == last ? -1 : last;
  if (row != last) {
    const size_t rb = (size_t)ix->ld * ix->esize;
    CU(cudaMemcpyAsync((char*)ix->rows + row * rb, (char*)ix->rows + last * rb, rb, cudaMemcpyDeviceToDevice, ix->stream));
    CU(cudaMemcpyAsync(ix->inv + row, ix->inv + last, sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
    if (ix->mask)
      CU(cudaMemcpyAsync(ix->mask + row * vs::kMaskWords, ix->mask + last * vs::kMaskWords, vs::kMaskWords * 8,
                         cudaMemcpyDeviceToDevice, ix->stream));
  }
  if (ix->mask) CU(cudaMemsetAsync(ix->mask + last * vs::kMaskWords, 0, vs::kMaskWords * 8, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  ix->n = last;
  return VS_OK;
}

int vs_set_row_host(vs_index_t* ix, int64_t row, const float* vec) {
  if (!ix || !vec) return fail(VS_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (row < 0 || row >= ix->n) return fail(VS_ERR_ARG, "row %lld out of range [0,%lld)", (long long)row, (long long)ix->n);
  DeviceGuard g(ix->device);
  CU(ix->d_stage.reserve((size_t)ix->dim * 4));
  CU(cudaMemcpyAsync(ix->d_stage.p, vec, (size_t)ix->dim * 4, cudaMemcpyHostToDevice, ix->stream));
  char* dst = (char*)ix->rows + (size_t)row * ix->ld * ix->esize;
  CU(vs::launch_ingest((const float*)ix->d_stage.p, 1, ix->dim, ix->dtype, dst, ix->ld, ix->inv + row, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_clear(vs_index_t* ix) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  if (ix->mask && ix->n > 0) CU(cudaMemsetAsync(ix->mask, 0, (size_t)ix->n * vs::kMaskWords * 8, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  ix->n = 0;
  return VS_OK;
}

int vs_set_mask_bits(vs_index_t* ix, int64_t row, const uint64_t bits[VS_MASK_WORDS]) {
  if (!ix || !bits) return fail(VS_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (row < 0 || row >= ix->n) return fail(VS_ERR_ARG, "row %lld out of range", (long long)row);
  DeviceGuard g(ix->device);
  int rc = ensure_mask(ix);
  if (rc) return rc;
  CU(cudaMemcpyAsync(ix->mask + row * vs::kMaskWords, bits, vs::kMaskWords * 8, cudaMemcpyHostToDevice, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_set_mask_bits_range(vs_index_t* ix, int64_t first_row, int64_t n, const uint64_t* bits) {
  if (!ix || (!bits && n > 0)) return fail(VS_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (first_row < 0 || n < 0 || first_row + n > ix->n) return fail(VS_ERR_ARG, "row range out of bounds");
  if (n == 0) return VS_OK;
  DeviceGuard g(ix->device);
  int rc = ensure_mask(ix);
  if (rc) return rc;
  CU(cudaMemcpyAsync(ix->mask + first_row * vs::kMaskWords, bits, (size_t)n * vs::kMaskWords * 8, cudaMemcpyHostToDevice,
                     ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_get_mask_bits(const vs_index_t* cix, int64_t row, uint64_t bits[VS_MASK_WORDS]) {
  vs_index* ix = const_cast<vs_index*>(cix);
  if (!ix || !bits) return fail(VS_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (row < 0 || row >= ix->n) return fail(VS_ERR_ARG, "row %lld out of range", (long long)row);
  memset(bits, 0, vs::kMaskWords * 8);
  if (!ix->mask) return VS_OK;
  DeviceGuard g(ix->device);
  CU(cudaMemcpyAsync(bits, ix->mask + row * vs::kMaskWords, vs::kMaskWords * 8, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_get_rows_host(const vs_index_t* cix, int64_t first_row, int64_t n, float* out) {
  vs_index* ix = const_cast<vs_index*>(cix);
  if (!ix || (!out && n > 0)) return fail(VS_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (first_row < 0 || n < 0 || first_row + n > ix->n) return fail(VS_ERR_ARG, "row range out of bounds");
  if (n == 0) return VS_OK;
  DeviceGuard g(ix->device);
  CU(ix->d_stage.reserve((size_t)n * ix->dim * 4));
  const char* src = (const char*)ix->rows + (size_t)first_row * ix->ld * ix->esize;
  CU(vs::launch_export(src, n, ix->dim, ix->dtype, ix->ld, (float*)ix->d_stage.p, ix->stream));
  CU(cudaMemcpyAsync(out, ix->d_stage.p, (size_t)n * ix->dim * 4, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_get_rows_dev(const vs_index_t* cix, int64_t first_row, int64_t n, float* out_dev, void* stream) {
  vs_index* ix = const_cast<vs_index*>(cix);
  if (!ix || (!out_dev && n > 0)) return fail(VS_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (first_row < 0 || n < 0 || first_row + n > ix->n) return fail(VS_ERR_ARG, "row range out of bounds");
  if (n == 0) return VS_OK;
  DeviceGuard g(ix->device);
  const char* src = (const char*)ix->rows + (size_t)first_row * ix->ld * ix->esize;
  CU(vs::launch_export(src, n, ix->dim, ix->dtype, ix->ld, out_dev, pick_stream(ix, stream)));
  return VS_OK;
}

int vs_query_topk_dev(vs_index_t* ix, const float* q_dev, int B, int k, const uint64_t* require_bits, int mode,
                      float* out_scores_dev, int64_t* out_rows_dev, void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B < 0 || (B > 0 && (!q_dev || !out_scores_dev || !out_rows_dev))) return fail(VS_ERR_ARG, "NULL buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  return query_dev_locked(ix, q_dev, B, k, require_bits, mode, out_scores_dev, out_rows_dev, pick_stream(ix, stream));
}

int vs_query_topk_host(vs_index_t* ix, const float* q, int B, int k, const uint64_t* require_bits, int mode,
                       float* out_scores, int64_t* out_rows) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B < 0 || (B > 0 && (!q || !out_scores || !out_rows))) return fail(VS_ERR_ARG, "NULL buffer");
  if (B == 0) return VS_OK;
  if (k <= 0 || k > vs::kMaxK) return fail(VS_ERR_ARG, "k=%d out of range [1,%d]", k, vs::kMaxK);
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  const size_t qbytes = (size_t)B * ix->dim * 4, sbytes = (size_t)B * k * 4, rbytes = (size_t)B * k * 8;
  CU(ix->d_q.reserve(qbytes));
  CU(ix->d_out_s.reserve(sbytes));
  CU(ix->d_out_r.reserve(rbytes));
  CU(ix->h_in.reserve(qbytes));
  CU(ix->h_out.reserve(sbytes + rbytes));
  memcpy(ix->h_in.p, q, qbytes);  // pinned staging so the H2D copy is a true async DMA
  CU(cudaMemcpyAsync(ix->d_q.p, ix->h_in.p, qbytes, cudaMemcpyHostToDevice, ix->stream));
  int rc = query_dev_locked(ix, (const float*)ix->d_q.p, B, k, require_bits, mode, (float*)ix->d_out_s.p,
                            (int64_t*)ix->d_out_r.p, ix->stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(ix->h_out.p, ix->d_out_s.p, sbytes, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaMemcpyAsync((char*)ix->h_out.p + sbytes, ix->d_out_r.p, rbytes, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  memcpy(out_scores, ix->h_out.p, sbytes);
  memcpy(out_rows, (char*)ix->h_out.p + sbytes, rbytes);
  return VS_OK;
}

int vs_blend_dev(vs_index_t* ix, const float* img_dev, const float* txt_dev, const double* w_dev, int B, float* out_dev,
                 void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B < 0 || (B > 0 && (!img_dev || !txt_dev || !w_dev || !out_dev))) return fail(VS_ERR_ARG, "NULL buffer");
  DeviceGuard g(ix->device);
  CU(vs::launch_blend(img_dev, txt_dev, w_dev, B, ix->dim, out_dev, pick_stream(ix, stream)));
  return VS_OK;
}

int vs_query_multimodal_host(vs_index_t* ix, const float* img, const float* txt, const double* w, int B, int k,
                             const uint64_t* require_bits, int mode, float* out_scores, int64_t* out_rows) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B < 0 || (B > 0 && (!img || !txt || !w || !out_scores || !out_rows))) return fail(VS_ERR_ARG, "NULL buffer");
  if (B == 0) return VS_OK;
  if (k <= 0 || k > vs::kMaxK) return fail(VS_ERR_ARG, "k=%d out of range [1,%d]", k, vs::kMaxK);
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  const size_t qbytes = (size_t)B * ix->dim * 4, wbytes = (size_t)B * 8;
  const size_t sbytes = (size_t)B * k * 4, rbytes = (size_t)B * k * 8;
  // device staging: [img | txt | blended] in d_stage, weights in d_misc
  CU(ix->d_stage.reserve(3 * qbytes));
  CU(ix->d_misc.reserve(wbytes));
  CU(ix->d_out_s.reserve(sbytes));
  CU(ix->d_out_r.reserve(rbytes));
  CU(ix->h_in.reserve(2 * qbytes + wbytes));
  CU(ix->h_out.reserve(sbytes + rbytes));
  char* hin = (char*)ix->h_in.p;
  memcpy(hin, img, qbytes);
  memcpy(hin + qbytes, txt, qbytes);
  memcpy(hin + 2 * qbytes, w, wbytes);
  char* dst = (char*)ix->d_stage.p;
  CU(cudaMemcpyAsync(dst, hin, 2 * qbytes, cudaMemcpyHostToDevice, ix->stream));
  CU(cudaMemcpyAsync(ix->d_misc.p, hin + 2 * qbytes, wbytes, cudaMemcpyHostToDevice, ix->stream));
  float* blended = (float*)(dst + 2 * qbytes);
  CU(vs::launch_blend((const float*)dst, (const float*)(dst + qbytes), (const double*)ix->d_misc.p, B, ix->dim, blended,
                      ix->stream));
  int rc = query_dev_locked(ix, blended, B, k, require_bits, mode, (float*)ix->d_out_s.p, (int64_t*)ix->d_out_r.p,
                            ix->stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(ix->h_out.p, ix->d_out_s.p, sbytes, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaMemcpyAsync((char*)ix->h_out.p + sbytes, ix->d_out_r.p, rbytes, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  memcpy(out_scores, ix->h_out.p, sbytes);
  memcpy(out_rows, (char*)ix->h_out.p + sbytes, rbytes);
  return VS_OK;
}

int vs_merge_topk_dev(vs_index_t* ix, const float* cand_scores_dev, const int64_t* cand_rows_dev, int G, int B, int k,
                      float* out_scores_dev, int64_t* out_rows_dev, void* stream) {
  if (!cand_scores_dev || !cand_rows_dev || !out_scores_dev || !out_rows_dev) return fail(VS_ERR_ARG, "NULL buffer");
  if (G <= 0 || B <= 0 || k <= 0 || (int64_t)G * k > 16384) return fail(VS_ERR_ARG, "G*k must be in [1,16384]");
  cudaStream_t st = stream ? (cudaStream_t)stream : (ix ? ix->stream : (cudaStream_t) nullptr);
  if (ix) {
    DeviceGuard g(ix->device);
    CU(vs::launch_merge(cand_scores_dev, cand_rows_dev, G, B, k, out_scores_dev, out_rows_dev, st));
  } else {
    CU(vs::launch_merge(cand_scores_dev, cand_rows_dev, G, B, k, out_scores_dev, out_rows_dev, st));
  }
  return VS_OK;
}

size_t vs_exchange_bytes(int B_max, int k_max, int G) {
  if (B_max <= 0 || k_max <= 0 || G <= 0 || G > vs::kMaxPeers) return 0;
  return vs::exchange_bytes(B_max, k_max, G);
}

static int exchange_create_locked(vs_index* ix, int G, int rank, int B_max, int k_max);

int vs_exchange_create(vs_index_t* ix, int G, int rank, int B_max, int k_max) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (G <= 0 || G > vs::kMaxPeers || rank < 0 || rank >= G)
    return fail(VS_ERR_ARG, "need 1 <= G <= %d and 0 <= rank < G", vs::kMaxPeers);
  if (B_max <= 0 || B_max > 65536 || k_max <= 0 || k_max > vs::kMaxFusedK)
    return fail(VS_ERR_ARG, "need 1 <= B_max <= 65536 and 1 <= k_max <= %d", vs::kMaxFusedK);
  std::lock_guard<std::mutex> lk(ix->mu);
  if (ix->xc.local) return fail(VS_ERR_ARG, "exchange already created for this index");
  DeviceGuard g(ix->device);
  const int rc = exchange_create_locked(ix, G, rank, B_max, k_max);
  if (rc != VS_OK && ix->xc.local) {   // roll back: a later retry (or the NCCL arm) starts clean
    cudaFree(ix->xc.local);
    cudaGetLastError();
    ix->xc.local = nullptr;
  }
  return rc;
}

static int exchange_create_locked(vs_index* ix, int G, int rank, int B_max, int k_max) {
  const size_t bytes = vs::exchange_bytes(B_max, k_max, G);
  // plain cudaMalloc (not a pool allocation): required for cudaIpcGetMemHandle
  CU(cudaMalloc(&ix->xc.local, bytes));
  CU(cudaMemset(ix->xc.local, 0, bytes));
  CU(ix->d_err.reserve(16));
  CU(cudaMemset(ix->d_err.p, 0, 16));
  // no allocation on the sharded query path: a cudaFree there would synchronise the device while a
  // peer may be spinning on this rank's push.  Reserve every scratch buffer for (B_max, k_max) now.
  CU(ix->d_xs.reserve((size_t)B_max * k_max * sizeof(float)));
  CU(ix->d_xr.reserve((size_t)B_max * k_max * sizeof(int64_t)));
  CU(ix->d_part_s.reserve((size_t)kScanBatch * ix->sm_count * vs::kScanMaxCtasPerSm * k_max * sizeof(float)));
  CU(ix->d_part_r.reserve((size_t)kScanBatch * ix->sm_count * vs::kScanMaxCtasPerSm * k_max * sizeof(uint32_t)));
  if (ix->tickets_n < (size_t)kScanBatch) {
    CU(ix->d_tickets.reserve(kScanBatch * sizeof(unsigned int)));
    CU(cudaMemset(ix->d_tickets.p, 0, ix->d_tickets.bytes));
    ix->tickets_n = kScanBatch;
  }
  if (ix->dtype == VS_BF16 && vs::tensor_dim_ok(ix->dim) && vs::tensor_path_available()) {
    const int64_t rows_hint = ix->cap > ix->n ? ix->cap : ix->n;
    const int kt = k_max < vs::kMaxTensorK ? k_max : vs::kMaxTensorK;
    CU(ix->d_tensor.reserve(vs::tensor_workspace_bytes(B_max, ix->dim, kt, ix->sm_count, rows_hint > 0 ? rows_hint : 1)));
  }
  CU(vs::preload_exchange_kernels());
  CU(cudaDeviceSynchronize());
  ix->xc.G = G;
  ix->xc.rank = rank;
  ix->xc.Bmax = B_max;
  ix->xc.kmax = k_max;
  ix->xc.bytes = bytes;
  ix->xc.peers[rank] = ix->xc.local;
  ix->xc.attached = (G == 1);
  return VS_OK;
}

int vs_exchange_ipc_handle(vs_index_t* ix, unsigned char handle_out[64]) {
  if (!ix || !handle_out) return fail(VS_ERR_ARG, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (!ix->xc.local) return fail(VS_ERR_ARG, "vs_exchange_create has not been called");
  DeviceGuard g(ix->device);
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, ix->xc.local));
  memcpy(handle_out, &h, 64);
  return VS_OK;
}

void* vs_exchange_local_ptr(vs_index_t* ix) { return ix ? ix->xc.local : nullptr; }

int vs_exchange_attach(vs_index_t* ix, const unsigned char* ipc_handles, void* const* peer_ptrs) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (!ix->xc.local) return fail(VS_ERR_ARG, "vs_exchange_create has not been called");
  if (ix->xc.attached && ix->xc.G > 1) return fail(VS_ERR_ARG, "exchange already attached");
  DeviceGuard g(ix->device);
  for (int p = 0; p < ix->xc.G; ++p) {
    if (p == ix->xc.rank) continue;
    if (peer_ptrs && peer_ptrs[p]) {
      // a pointer of this process: make sure this device may dereference it
      cudaPointerAttributes at;
      CU(cudaPointerGetAttributes(&at, peer_ptrs[p]));
      if (at.type != cudaMemoryTypeDevice) return fail(VS_ERR_ARG, "peer pointer %d is not device memory", p);
      if (at.device != ix->device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, ix->device, at.device));
        if (!can) return fail(VS_ERR_UNSUPPORTED, "device %d cannot access peer device %d", ix->device, at.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
        cudaGetLastError();
      }
      ix->xc.peers[p] = peer_ptrs[p];
    } else if (ipc_handles) {
      cudaIpcMemHandle_t h;
      memcpy(&h, ipc_handles + (size_t)64 * p, 64);
      void* ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(VS_ERR_CUDA, "cudaIpcOpenMemHandle(peer %d): %s", p, cudaGetErrorString(e));
      }
      ix->xc.peers[p] = ptr;
      ix->xc.ipc_opened[p] = true;
    } else {
      return fail(VS_ERR_ARG, "no pointer or IPC handle for peer %d", p);
    }
  }
  ix->xc.attached = true;
  return VS_OK;
}

namespace {
int exchange_ready(vs_index* ix, int k) {
  if (!ix->xc.local || !ix->xc.attached) return fail(VS_ERR_ARG, "exchange not created/attached");
  if (k <= 0 || k > ix->xc.kmax) return fail(VS_ERR_UNSUPPORTED, "k=%d exceeds the exchange's k_max=%d", k, ix->xc.kmax);
  if (ix->row_base < 0 || ix->row_base + ix->n * ix->row_stride > 0xFFFFFFF0LL)
    return fail(VS_ERR_UNSUPPORTED, "sharded queries need global rows < 2^32");
  return VS_OK;
}
// the exchange kernel on [B,k] candidates, in chunks of B_max slots
int exchange_chunks(vs_index* ix, const float* cs, const int64_t* cr, int B, int k, float* out_s, int64_t* out_r,
                    cudaStream_t st) {
  for (int b0 = 0; b0 < B; b0 += ix->xc.Bmax) {
    const int nb = B - b0 < ix->xc.Bmax ? B - b0 : ix->xc.Bmax;
    const vs::XchgParams x = make_exchange(ix, 0, true, false);
    CU(vs::launch_exchange_merge(cs + (size_t)b0 * k, cr + (size_t)b0 * k, x, nb, k, out_s + (size_t)b0 * k,
                                 out_r + (size_t)b0 * k, ix->sm_count, st));
  }
  return VS_OK;
}
}  // namespace

int vs_query_topk_sharded_dev(vs_index_t* ix, const float* q_dev, int B, int k, const uint64_t* require_bits, int mode,
                              float* out_scores_dev, int64_t* out_rows_dev, void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B <= 0 || !q_dev || !out_scores_dev || !out_rows_dev) return fail(VS_ERR_ARG, "bad B or NULL buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  int rc = exchange_ready(ix, k);
  if (rc) return rc;
  cudaStream_t st = pick_stream(ix, stream);
  if (ix->xc.G == 1) return query_dev_locked(ix, q_dev, B, k, require_bits, mode, out_scores_dev, out_rows_dev, st);
  int path = mode;
  if (path == VS_Q_AUTO)
    path = (ix->dtype == VS_BF16 && B >= 16 && k <= vs::kMaxTensorK && vs::tensor_dim_ok(ix->dim) &&
            vs::tensor_path_available()) ? VS_Q_TENSOR : VS_Q_SCAN;
  // fused form: the shard's result never leaves the scan kernel (one launch, exchange inside)
  if (path == VS_Q_SCAN && ix->n > 0 && B <= kScanBatch && B <= ix->xc.Bmax) {
    bool fused = false;
    rc = query_dev_locked(ix, q_dev, B, k, require_bits, VS_Q_SCAN, out_scores_dev, out_rows_dev, st, &fused);
    if (rc) return rc;
    if (fused) return VS_OK;
    // (a filter nobody carries bits for produced an all-empty local result in out_*: exchange it)
    CU(cudaMemcpyAsync(ix->d_xs.p, out_scores_dev, (size_t)B * k * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(ix->d_xr.p, out_rows_dev, (size_t)B * k * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    return exchange_chunks(ix, (const float*)ix->d_xs.p, (const int64_t*)ix->d_xr.p, B, k, out_scores_dev, out_rows_dev, st);
  }
  // general form, in chunks of B_max queries: local query (K1 batches or K2) into scratch, then the
  // exchange kernel; the scratch was reserved by vs_exchange_create
  for (int b0 = 0; b0 < B; b0 += ix->xc.Bmax) {
    const int nb = B - b0 < ix->xc.Bmax ? B - b0 : ix->xc.Bmax;
    rc = query_dev_locked(ix, q_dev + (size_t)b0 * ix->dim, nb, k, require_bits, path, (float*)ix->d_xs.p,
                          (int64_t*)ix->d_xr.p, st);
    if (rc) return rc;
    rc = exchange_chunks(ix, (const float*)ix->d_xs.p, (const int64_t*)ix->d_xr.p, nb, k, out_scores_dev + (size_t)b0 * k,
                         out_rows_dev + (size_t)b0 * k, st);
    if (rc) return rc;
  }
  return VS_OK;
}

int vs_query_topk_sharded_host(vs_index_t* ix, const float* q, int B, int k, const uint64_t* require_bits, int mode,
                               float* out_scores, int64_t* out_rows) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B <= 0 || !q || !out_scores || !out_rows) return fail(VS_ERR_ARG, "bad B or NULL buffer");
  const size_t qbytes = (size_t)B * ix->dim * 4, sbytes = (size_t)B * k * 4, rbytes = (size_t)B * k * 8;
  {
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    CU(ix->d_q.reserve(qbytes));
    CU(ix->d_out_s.reserve(sbytes + rbytes + 64));   // scores | rows in ONE buffer: one D2H
    CU(ix->h_in.reserve(qbytes));
    CU(ix->h_out.reserve(sbytes + rbytes + 64));
    memcpy(ix->h_in.p, q, qbytes);
    CU(cudaMemcpyAsync(ix->d_q.p, ix->h_in.p, qbytes, cudaMemcpyHostToDevice, ix->stream));
  }
  float* d_s = (float*)ix->d_out_s.p;
  int64_t* d_r = (int64_t*)((char*)ix->d_out_s.p + ((sbytes + 15) & ~(size_t)15));
  int rc = vs_query_topk_sharded_dev(ix, (const float*)ix->d_q.p, B, k, require_bits, mode, d_s, d_r, nullptr);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  const size_t roff = (sbytes + 15) & ~(size_t)15;
  CU(cudaMemcpyAsync(ix->h_out.p, ix->d_out_s.p, roff + rbytes, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  memcpy(out_scores, ix->h_out.p, sbytes);
  memcpy(out_rows, (char*)ix->h_out.p + roff, rbytes);
  return VS_OK;
}

int vs_exchange_begin(vs_index_t* ix) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (!ix->xc.local || !ix->xc.attached) return fail(VS_ERR_ARG, "exchange not created/attached");
  if (++ix->xc.epoch == 0) ix->xc.epoch = 2;
  return VS_OK;
}

int vs_query_topk_push_dev(vs_index_t* ix, const float* q_dev, int B, int k, const uint64_t* require_bits, int mode,
                           int slot0, void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B <= 0 || !q_dev) return fail(VS_ERR_ARG, "bad B or NULL buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  int rc = exchange_ready(ix, k);
  if (rc) return rc;
  if (ix->xc.epoch == 0) return fail(VS_ERR_ARG, "vs_exchange_begin has not been called");
  if (slot0 < 0 || slot0 + B > ix->xc.Bmax) return fail(VS_ERR_ARG, "slots [%d,%d) exceed B_max=%d", slot0, slot0 + B, ix->xc.Bmax);
  cudaStream_t st = pick_stream(ix, stream);
  int path = mode;
  if (path == VS_Q_AUTO)
    path = (ix->dtype == VS_BF16 && B >= 16 && k <= vs::kMaxTensorK && vs::tensor_dim_ok(ix->dim) &&
            vs::tensor_path_available()) ? VS_Q_TENSOR : VS_Q_SCAN;
  float* xs = (float*)ix->d_xs.p;
  int64_t* xr = (int64_t*)ix->d_xr.p;
  if (path == VS_Q_SCAN && ix->n > 0 && B <= kScanBatch) {
    bool fused = false;
    rc = query_dev_locked(ix, q_dev, B, k, require_bits, VS_Q_SCAN, xs, xr, st, &fused, slot0);
    if (rc) return rc;
    if (fused) return VS_OK;          // the scan kernel's last CTA pushed the candidates
  } else {
    rc = query_dev_locked(ix, q_dev, B, k, require_bits, path, xs, xr, st);
    if (rc) return rc;
  }
  CU(vs::launch_exchange_merge(xs, xr, make_exchange(ix, slot0, false, true), B, k, nullptr, nullptr, ix->sm_count, st, 1));
  return VS_OK;
}

int vs_exchange_collect_dev(vs_index_t* ix, int B, int k, float* out_scores_dev, int64_t* out_rows_dev, void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B <= 0 || !out_scores_dev || !out_rows_dev) return fail(VS_ERR_ARG, "bad B or NULL buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  int rc = exchange_ready(ix, k);
  if (rc) return rc;
  if (B > ix->xc.Bmax) return fail(VS_ERR_ARG, "B=%d exceeds B_max=%d", B, ix->xc.Bmax);
  cudaStream_t st = pick_stream(ix, stream);
  if (ix->last_stream_valid && ix->last_stream != st) {
    CU(cudaEventRecord(ix->ev, ix->last_stream));
    CU(cudaStreamWaitEvent(st, ix->ev, 0));
  }
  ix->last_stream = st;
  ix->last_stream_valid = true;
  CU(vs::launch_exchange_merge(nullptr, nullptr, make_exchange(ix, 0, false, false), B, k, out_scores_dev, out_rows_dev,
                               ix->sm_count, st, 2));
  return VS_OK;
}

int vs_exchange_merge_dev(vs_index_t* ix, const float* cand_scores_dev, const int64_t* cand_rows_dev, int B, int k,
                          float* out_scores_dev, int64_t* out_rows_dev, void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B <= 0 || !cand_scores_dev || !cand_rows_dev || !out_scores_dev || !out_rows_dev)
    return fail(VS_ERR_ARG, "bad B or NULL buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  int rc = exchange_ready(ix, k);
  if (rc) return rc;
  cudaStream_t st = pick_stream(ix, stream);
  // exchanges of one index are ordered: serialise against the stream that ran the previous one
  if (ix->last_stream_valid && ix->last_stream != st) {
    CU(cudaEventRecord(ix->ev, ix->last_stream));
    CU(cudaStreamWaitEvent(st, ix->ev, 0));
  }
  ix->last_stream = st;
  ix->last_stream_valid = true;
  return exchange_chunks(ix, cand_scores_dev, cand_rows_dev, B, k, out_scores_dev, out_rows_dev, st);
}

int vs_exchange_error(vs_index_t* ix) {
  if (!ix || !ix->d_err.p) return 0;
  DeviceGuard g(ix->device);
  unsigned int v = 0;
  if (cudaMemcpy(&v, ix->d_err.p, 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return (int)v;
}

int64_t vs_filter_words(const vs_index_t* ix) {
  if (!ix) return 0;
  return (ix->n + 255) / 256 * 8;   // whole 256-row tiles: the sweep writes full tiles
}

int vs_filter_sweep_dev(vs_index_t* ix, const float* prompts_dev, int F, float tau, uint32_t* out_bits_dev, void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (F <= 0 || !prompts_dev || !out_bits_dev) return fail(VS_ERR_ARG, "bad arguments");
  if (ix->dtype != VS_BF16) return fail(VS_ERR_UNSUPPORTED, "filter sweep needs bf16 storage (tcgen05 path)");
  if (!vs::tensor_path_available()) return fail(VS_ERR_UNSUPPORTED, "tensor path not built");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  if (ix->n == 0) return VS_OK;
  vs::TensorArgs ta;
  ta.rows = ix->rows;
  ta.inv_norm = ix->inv;
  ta.mask = nullptr;
  memset(ta.req, 0, sizeof(ta.req));
  ta.dim = ix->dim;
  ta.ld_elems = ix->ld;
  ta.n_rows = ix->n;
  ta.row_base = ix->row_base;
  ta.row_stride = ix->row_stride;
  CU(ix->d_tensor.reserve(vs::tensor_workspace_bytes(F, ix->dim, 1, ix->sm_count, ix->n)));
  CU(vs::launch_tensor_filter(ta, prompts_dev, F, tau, ix->d_tensor.p, out_bits_dev, vs_filter_words(ix), ix->sm_count,
                              pick_stream(ix, stream)));
  return VS_OK;
}

int vs_filter_sweep_host(vs_index_t* ix, const float* prompts, int F, float tau, uint32_t* out_bits) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (F <= 0 || !prompts || !out_bits) return fail(VS_ERR_ARG, "bad arguments");
  const size_t pbytes = (size_t)F * ix->dim * 4;
  const size_t obytes = (size_t)F * vs_filter_words(ix) * 4;
  {
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    CU(ix->d_q.reserve(pbytes));
    CU(ix->d_misc.reserve(obytes + 16));
    CU(cudaMemcpyAsync(ix->d_q.p, prompts, pbytes, cudaMemcpyHostToDevice, ix->stream));
  }
  int rc = vs_filter_sweep_dev(ix, (const float*)ix->d_q.p, F, tau, (uint32_t*)ix->d_misc.p, nullptr);
  if (rc) return rc;
  DeviceGuard g(ix->device);
  CU(cudaMemcpyAsync(out_bits, ix->d_misc.p, obytes, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_dedup_dev(vs_index_t* ix, int64_t row_lo, int64_t row_hi, float tau, int64_t cap, int64_t* out_i_dev,
                 int64_t* out_j_dev, float* out_score_dev, unsigned long long* out_count_dev, void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (!out_i_dev || !out_j_dev || !out_score_dev || !out_count_dev || cap < 0) return fail(VS_ERR_ARG, "bad arguments");
  if (ix->dtype != VS_BF16) return fail(VS_ERR_UNSUPPORTED, "dedup needs bf16 storage (tcgen05 path)");
  if (!vs::tensor_path_available()) return fail(VS_ERR_UNSUPPORTED, "tensor path not built");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  if (row_lo < 0) row_lo = 0;
  if (row_hi > ix->n) row_hi = ix->n;
  cudaStream_t st = pick_stream(ix, stream);
  CU(cudaMemsetAsync(out_count_dev, 0, sizeof(unsigned long long), st));
  if (row_lo >= row_hi) return VS_OK;
  vs::TensorArgs ta;
  ta.rows = ix->rows;
  ta.inv_norm = ix->inv;
  ta.mask = nullptr;
  memset(ta.req, 0, sizeof(ta.req));
  ta.dim = ix->dim;
  ta.ld_elems = ix->ld;
  ta.n_rows = ix->n;
  ta.row_base = ix->row_base;
  ta.row_stride = ix->row_stride;
  CU(ix->d_tensor.reserve(vs::tensor_workspace_bytes(128, ix->dim, 1, ix->sm_count, ix->n)));
  CU(vs::launch_tensor_dedup(ta, row_lo, row_hi, tau, cap, out_i_dev, out_j_dev, out_score_dev, out_count_dev,
                             ix->d_tensor.p, ix->sm_count, st));
  return VS_OK;
}

int vs_dedup_host(vs_index_t* ix, int64_t row_lo, int64_t row_hi, float tau, int64_t cap, int64_t* out_i, int64_t* out_j,
                  float* out_score, int64_t* out_count) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (!out_i || !out_j || !out_score || !out_count || cap < 0) return fail(VS_ERR_ARG, "bad arguments");
  void* buf = nullptr;
  {
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    CU(ix->d_misc.reserve((size_t)cap * 20 + 64));
    buf = ix->d_misc.p;
  }
  unsigned long long* d_count = (unsigned long long*)buf;
  int64_t* d_i = (int64_t*)((char*)buf + 16);
  int64_t* d_j = d_i + cap;
  float* d_s = (float*)(d_j + cap);
  int rc = vs_dedup_dev(ix, row_lo, row_hi, tau, cap, d_i, d_j, d_s, d_count, nullptr);
  if (rc) return rc;
  DeviceGuard g(ix->device);
  unsigned long long cnt = 0;
  CU(cudaMemcpyAsync(&cnt, d_count, 8, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  *out_count = (int64_t)cnt;
  const int64_t m = (int64_t)cnt < cap ? (int64_t)cnt : cap;
  if (m > 0) {
    CU(cudaMemcpyAsync(out_i, d_i, m * 8, cudaMemcpyDeviceToHost, ix->stream));
    CU(cudaMemcpyAsync(out_j, d_j, m * 8, cudaMemcpyDeviceToHost, ix->stream));
    CU(cudaMemcpyAsync(out_score, d_s, m * 4, cudaMemcpyDeviceToHost, ix->stream));
    CU(cudaStreamSynchronize(ix->stream));
  }
  if ((int64_t)cnt > cap) return fail(VS_ERR_OVERFLOW, "found %lld pairs, capacity %lld", (long long)cnt, (long long)cap);
  return VS_OK;
}

}  // extern "C"
