// api.cu -- the C ABI declared in include/vecsearch_b200.h for ONE row shard: index handle, slab
// management, host<->device staging, path selection.  No compute happens on the host.
// (group.cu builds the single-process multi-GPU collection on top of these bodies.)
//
// Thread safety: every entry point takes the index's mutex for its WHOLE body (staging buffers are
// per index), so concurrent callers on one handle serialise; "_dev" entry points return once the work
// is enqueued.  Mutations order themselves after queries still in flight on caller streams.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <new>
#include <vector>

#include "index_internal.h"

namespace vs {
static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace vs

using vs::DeviceGuard;
using vs::fail;
using vs::kScanBatch;
using vs::pick_stream;
using vs::touch;
#define CU(call) VS_CU(call)

namespace vs {

static int64_t pitch_elems(int dim, int esize) {
  const int per16 = 16 / esize;
  return ((int64_t)dim + per16 - 1) / per16 * per16;
}
// groups of 32 rows covered by whole 256-row tiles (what the tcgen05 kernels read)
static int64_t gmin_groups(int64_t rows) { return (rows + 255) / 256 * 8; }

int grow_locked(vs_index* ix, int64_t need_rows) {
  if (need_rows <= ix->cap) return VS_OK;
  int64_t ncap = ix->cap > 0 ? ix->cap : 1024;
  while (ncap < need_rows) ncap += ncap < (1 << 20) ? ncap : ncap / 2;
  const size_t row_bytes = (size_t)ix->ld * ix->esize;
  void* nrows = nullptr;
  float* ninv = nullptr;
  float* ngmin = nullptr;
  uint64_t* nmask = nullptr;
  auto bail = [&](const char* what, cudaError_t e) {
    cudaGetLastError();
    if (nrows) cudaFree(nrows);
    if (ninv) cudaFree(ninv);
    if (ngmin) cudaFree(ngmin);
    if (nmask) cudaFree(nmask);
    return fail(VS_ERR_OOM, "cudaMalloc %s for %lld rows failed: %s", what, (long long)ncap, cudaGetErrorString(e));
  };
  cudaError_t e = cudaMalloc(&nrows, (size_t)ncap * row_bytes + 256);
  if (e != cudaSuccess) return bail("rows", e);
  // +512 floats: tile-granular bulk copies of inverse norms never leave the allocation
  e = cudaMalloc((void**)&ninv, ((size_t)ncap + 512) * sizeof(float));
  if (e != cudaSuccess) return bail("inverse norms", e);
  CU(cudaMemsetAsync(ninv, 0, ((size_t)ncap + 512) * sizeof(float), ix->stream));
  const size_t gmin_bytes = ((size_t)gmin_groups(ncap) + 64) * sizeof(float);
  if (ix->dtype == VS_BF16) {
    e = cudaMalloc((void**)&ngmin, gmin_bytes);
    if (e != cudaSuccess) return bail("group bounds", e);
    CU(cudaMemsetAsync(ngmin, 0x7f, gmin_bytes, ix->stream));   // 0x7f7f7f7f: a huge finite norm (groups without rows)
  }
  if (ix->mask) {
    e = cudaMalloc((void**)&nmask, (size_t)ncap * kMaskWords * 8);
    if (e != cudaSuccess) return bail("filter bits", e);
    CU(cudaMemsetAsync(nmask, 0, (size_t)ncap * kMaskWords * 8, ix->stream));
  }
  if (ix->n > 0) {
    CU(cudaMemcpyAsync(nrows, ix->rows, (size_t)ix->n * row_bytes, cudaMemcpyDeviceToDevice, ix->stream));
    CU(cudaMemcpyAsync(ninv, ix->inv, (size_t)ix->n * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
    if (ngmin) CU(cudaMemcpyAsync(ngmin, ix->gmin, (size_t)gmin_groups(ix->n) * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
    if (nmask) CU(cudaMemcpyAsync(nmask, ix->mask, (size_t)ix->n * kMaskWords * 8, cudaMemcpyDeviceToDevice, ix->stream));
  }
  // rare path: wait for everything (adds/queries may be in flight on caller streams)
  CU(cudaDeviceSynchronize());
  if (ix->rows) cudaFree(ix->rows);
  if (ix->inv) cudaFree(ix->inv);
  if (ix->gmin) cudaFree(ix->gmin);
  if (ix->mask) cudaFree(ix->mask);
  ix->rows = nrows;
  ix->inv = ninv;
  ix->gmin = ngmin;
  ix->mask = nmask;
  ix->cap = ncap;
  touch(ix);
  return VS_OK;
}

static int ensure_mask(vs_index* ix) {
  if (ix->mask) return VS_OK;
  if (ix->cap == 0) {
    int rc = grow_locked(ix, 1);
    if (rc) return rc;
  }
  CU(cudaMalloc((void**)&ix->mask, (size_t)ix->cap * kMaskWords * 8));
  CU(cudaMemsetAsync(ix->mask, 0, (size_t)ix->cap * kMaskWords * 8, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  touch(ix);
  return VS_OK;
}

// order `st` after whatever stream used the index's shared scratch / rows last
cudaError_t order_after_last(vs_index* ix, cudaStream_t st) {
  if (ix->last_stream_valid && ix->last_stream != st) {
    cudaError_t e = cudaEventRecord(ix->ev, ix->last_stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamWaitEvent(st, ix->ev, 0);
    if (e != cudaSuccess) return e;
  }
  ix->last_stream = st;
  ix->last_stream_valid = true;
  return cudaSuccess;
}

// recompute the group bounds of rows [row_lo, row_hi) (and of the empty groups up to the tile end)
cudaError_t refresh_gmin(vs_index* ix, int64_t row_lo, int64_t row_hi, cudaStream_t st) {
  if (!ix->gmin || row_hi <= row_lo) return cudaSuccess;
  int64_t g_hi = (row_hi + 31) / 32;
  if (row_hi >= ix->n) g_hi = gmin_groups(ix->n);   // the tail: up to the end of the last 256-row tile
  return launch_group_min(ix->inv, ix->n, row_lo / 32, g_hi, ix->gmin, st);
}

static uint32_t next_epoch(uint32_t e) { return e + 1 == 0 ? 2 : e + 1; }   // 0 is the "never written" flag value; keep the parity sequence

// kernel-side exchange descriptor.  `advance`: for the NEXT epoch (a new exchange) -- the caller commits
// it with `ix->xc.epoch = x.epoch` only once the launch that carries it has succeeded, so a failed call
// leaves this rank's epoch aligned with its peers'.
static XchgParams make_exchange(vs_index* ix, int slot0, bool advance, bool push_only) {
  XchgParams x = {};
  for (int g = 0; g < ix->xc.G; ++g) x.peers[g] = static_cast<unsigned char*>(ix->xc.peers[g]);
  x.err = ix->xc.h_err;
  x.G = ix->xc.G;
  x.rank = ix->xc.rank;
  x.Bmax = ix->xc.Bmax;
  x.kmax = ix->xc.kmax;
  x.slot0 = slot0;
  x.epoch = advance ? next_epoch(ix->xc.epoch) : ix->xc.epoch;
  x.push_only = push_only ? 1 : 0;
  return x;
}

static int pick_path(const vs_index* ix, int mode, int B, int k) {
  if (mode != VS_Q_AUTO) return mode;
  return (ix->dtype == VS_BF16 && B >= 16 && k <= kMaxTensorK && tensor_dim_ok(ix->dim)) ? VS_Q_TENSOR : VS_Q_SCAN;
}

static TensorArgs tensor_args(vs_index* ix, const uint64_t* reqw, bool use_mask) {
  TensorArgs ta;
  ta.rows = ix->rows;
  ta.inv_norm = ix->inv;
  ta.gmin = ix->gmin;
  ta.mask = use_mask ? ix->mask : nullptr;
  if (reqw) memcpy(ta.req, reqw, sizeof(ta.req)); else memset(ta.req, 0, sizeof(ta.req));
  ta.dim = ix->dim;
  ta.ld_elems = ix->ld;
  ta.n_rows = ix->n;
  ta.row_base = ix->row_base;
  ta.row_stride = ix->row_stride;
  return ta;
}

int query_dev_locked(vs_index* ix, const float* q_dev, int B, int k, const uint64_t* req, int mode_in, float* out_s,
                     int64_t* out_r, cudaStream_t st, QueryOpts* opts) {
  QueryOpts local_opts;
  QueryOpts& o = opts ? *opts : local_opts;
  o.fused = o.done_armed = false;
  const bool pipelined = (mode_in & VS_Q_PIPELINED) != 0;
  const int mode = mode_in & 0xff;
  if (mode != VS_Q_AUTO && mode != VS_Q_SCAN && mode != VS_Q_TENSOR) return fail(VS_ERR_ARG, "bad query mode %d", mode_in);
  if (B <= 0) return VS_OK;
  if (k <= 0 || k > kMaxK) return fail(VS_ERR_ARG, "k=%d out of range [1,%d]", k, kMaxK);
  // the scratch buffers (partial lists, tickets, select state) are shared by all queries of this
  // index: serialise against the stream that used them last
  CU(order_after_last(ix, st));
  bool use_mask = false;
  uint64_t reqw[kMaskWords] = {0, 0, 0, 0};
  if (req)
    for (int w = 0; w < kMaskWords; ++w) {
      reqw[w] = req[w];
      if (req[w]) use_mask = true;
    }
  if (ix->n == 0 || (use_mask && !ix->mask)) {
    // empty shard, or a filter requested while no row carries any bit: all slots empty
    CU(launch_fill_empty(out_s, out_r, (int64_t)B * k, st));
    touch(ix);
    return VS_OK;
  }
  const int path = pick_path(ix, mode, B, k);
  if (path == VS_Q_TENSOR) {
    if (ix->dtype != VS_BF16) return fail(VS_ERR_UNSUPPORTED, "tensor path needs bf16 storage");
    if (k > kMaxTensorK) return fail(VS_ERR_UNSUPPORTED, "tensor path supports k <= %d", kMaxTensorK);
    if (!tensor_dim_ok(ix->dim)) return fail(VS_ERR_UNSUPPORTED, "tensor path needs dim %% 8 == 0 and dim <= 4096");
    const TensorArgs ta = tensor_args(ix, reqw, use_mask);
    CU(ix->d_tensor.reserve(tensor_workspace_bytes(B, ix->dim, k, ix->sm_count)));
    touch(ix);
    CU(launch_tensor_topk(ta, q_dev, B, k, ix->d_tensor.p, out_s, out_r, ix->sm_count, st));
    ix->last_path = VS_Q_TENSOR;
    return VS_OK;
  }
  // ---- scan path ----
  const int64_t ld_bytes = ix->ld * ix->esize;
  if (scan_rows_per_tile(ix->dtype, ld_bytes) < 0)
    return fail(VS_ERR_UNSUPPORTED, "row pitch %lld bytes exceeds the scan kernel's 4096-byte limit", (long long)ld_bytes);
  const bool large_k = k > kMaxFusedK;
  const int step = large_k ? 4 : kScanBatch;
  if (!large_k) {
    CU(ix->d_part_s.reserve((size_t)step * ix->sm_count * kScanMaxCtasPerSm * k * sizeof(float)));
    CU(ix->d_part_r.reserve((size_t)step * ix->sm_count * kScanMaxCtasPerSm * k * sizeof(uint32_t)));
  } else {
    CU(ix->d_scores.reserve((size_t)step * ix->n * sizeof(float)));
    CU(ix->d_select.reserve(select_workspace_bytes(step)));
  }
  if (ix->tickets_n < (size_t)kScanBatch) {
    CU(ix->d_tickets.reserve(kScanBatch * sizeof(unsigned int)));
    CU(cudaMemsetAsync(ix->d_tickets.p, 0, ix->d_tickets.bytes, st));
    ix->tickets_n = kScanBatch;
    touch(ix);
  }
  const bool single_launch = !large_k && B <= step;
  for (int b0 = 0; b0 < B; b0 += step) {
    const int nb = B - b0 < step ? B - b0 : step;
    ScanArgs a;
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
    // The deferred griddepcontrol.wait (the next query streams while this one merges) is only legal when
    // the grid right before this one on the stream is a scan of this library AND nothing this kernel reads
    // is younger than it: rows/norms/bits untouched (chain_ok), and the caller vouches for q
    // (VS_Q_PIPELINED) or q belongs to this very call (b0 > 0: the first launch waited or was vouched for).
    a.early_wait = (b0 > 0 || (pipelined && ix->chain_ok && ix->chain_stream == st)) ? 0 : 1;
    bool advance = false;
    if (o.want_fused && single_launch) {
      advance = o.push_slot0 < 0;
      a.xg = o.push_slot0 >= 0 ? make_exchange(ix, o.push_slot0, false, true) : make_exchange(ix, 0, true, false);
      o.fused = true;
    }
    if (o.done_flag && single_launch) {
      a.done_flag = o.done_flag;
      a.done_seq = o.done_seq;
      o.done_armed = true;
    }
    if (o.q_host && single_launch && B == 1 && ix->dim <= kMaxInlineQ) a.q_host = o.q_host;
    else if (!q_dev) return fail(VS_ERR_ARG, "no device query buffer");
    const cudaError_t le = launch_scan(a, ix->sm_count, st);
    if (le != cudaSuccess) {
      o.fused = o.done_armed = false;
      touch(ix);
      CU(le);
    }
    if (advance) ix->xc.epoch = a.xg.epoch;
    ix->chain_ok = !large_k;
    ix->chain_stream = st;
    if (large_k)
      CU(launch_select((const float*)ix->d_scores.p, ix->n, nb, k, ix->row_base, ix->row_stride, ix->d_select.p, a.out_s, a.out_r, st));
  }
  ix->last_path = VS_Q_SCAN;
  return VS_OK;
}

int add_dev_locked(vs_index* ix, const float* rows_dev, int64_t n, int64_t* first_row, cudaStream_t st) {
  if (first_row) *first_row = ix->n;
  if (n == 0) return VS_OK;
  if (ix->n + n > 0xFFFFFFF0LL) return fail(VS_ERR_ARG, "shard would exceed 2^32 rows");
  int rc = grow_locked(ix, ix->n + n);
  if (rc) return rc;
  CU(order_after_last(ix, st));   // queries in flight on another stream read the tail group bounds
  char* dst = (char*)ix->rows + (size_t)ix->n * ix->ld * ix->esize;
  CU(launch_ingest(rows_dev, n, ix->dim, ix->dtype, dst, ix->ld, ix->inv + ix->n, st));
  if (ix->mask) CU(cudaMemsetAsync(ix->mask + (size_t)ix->n * kMaskWords, 0, (size_t)n * kMaskWords * 8, st));
  const int64_t old_n = ix->n;
  ix->n += n;
  touch(ix);
  CU(refresh_gmin(ix, old_n, ix->n, st));
  if (st != ix->stream) {
    // later work on the index's own stream (host-buffer queries) must see these rows
    CU(cudaEventRecord(ix->ev, st));
    CU(cudaStreamWaitEvent(ix->stream, ix->ev, 0));
  }
  return VS_OK;
}

int exchange_error_locked(vs_index* ix, bool clear) {
  if (!ix->xc.h_err) return VS_OK;
  volatile unsigned int* e = ix->xc.h_err;
  if (*e == 0) return VS_OK;
  if (clear) *e = 0;
  return fail(VS_ERR_EXCHANGE, "peer exchange timed out on shard %d of %d: a peer never pushed its candidates within ~3 s; "
                               "the affected results were returned EMPTY%s",
              ix->xc.rank, ix->xc.G, clear ? "" : " (vs_exchange_clear_error re-arms the exchange)");
}

static int exchange_ready(vs_index* ix, int k) {
  if (!ix->xc.local || !ix->xc.attached) return fail(VS_ERR_ARG, "exchange not created/attached");
  if (k <= 0 || k > ix->xc.kmax) return fail(VS_ERR_UNSUPPORTED, "k=%d exceeds the exchange's k_max=%d", k, ix->xc.kmax);
  if (ix->row_base < 0 || ix->row_base + ix->n * ix->row_stride > 0xFFFFFFF0LL)
    return fail(VS_ERR_UNSUPPORTED, "sharded queries need global rows < 2^32");
  return exchange_error_locked(ix, false);   // a timed-out exchange poisons the handle until it is acknowledged
}

// the exchange kernel on [B,k] candidates, in chunks of B_max slots
static int exchange_chunks(vs_index* ix, const float* cs, const int64_t* cr, int B, int k, float* out_s, int64_t* out_r,
                           cudaStream_t st) {
  touch(ix);
  for (int b0 = 0; b0 < B; b0 += ix->xc.Bmax) {
    const int nb = B - b0 < ix->xc.Bmax ? B - b0 : ix->xc.Bmax;
    const XchgParams x = make_exchange(ix, 0, true, false);
    CU(launch_exchange_merge(cs + (size_t)b0 * k, cr + (size_t)b0 * k, x, nb, k, out_s + (size_t)b0 * k,
                             out_r + (size_t)b0 * k, ix->sm_count, st));
    ix->xc.epoch = x.epoch;
  }
  return VS_OK;
}

int sharded_query_dev_locked(vs_index* ix, const float* q_dev, int B, int k, const uint64_t* req, int mode, float* out_s,
                             int64_t* out_r, cudaStream_t st, QueryOpts* opts) {
  int rc = exchange_ready(ix, k);
  if (rc) return rc;
  QueryOpts local_opts;
  QueryOpts& o = opts ? *opts : local_opts;
  if (ix->xc.G == 1) {
    o.want_fused = false;
    return query_dev_locked(ix, q_dev, B, k, req, mode, out_s, out_r, st, &o);
  }
  const int path = pick_path(ix, mode & 0xff, B, k);
  // fused form: the shard's result never leaves the scan kernel (one launch, exchange inside)
  if (path == VS_Q_SCAN && B <= kScanBatch && B <= ix->xc.Bmax && k <= kMaxFusedK) {
    o.want_fused = true;
    o.push_slot0 = -1;
    rc = query_dev_locked(ix, q_dev, B, k, req, (mode & ~0xff) | VS_Q_SCAN, out_s, out_r, st, &o);
    if (rc) return rc;
    if (o.fused) return VS_OK;
    // (empty shard / a filter nobody carries bits for: an all-empty local result sits in out_*: exchange it)
    CU(cudaMemcpyAsync(ix->d_xs.p, out_s, (size_t)B * k * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(ix->d_xr.p, out_r, (size_t)B * k * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    return exchange_chunks(ix, (const float*)ix->d_xs.p, (const int64_t*)ix->d_xr.p, B, k, out_s, out_r, st);
  }
  // general form, in chunks of B_max queries: local query (K1 batches or K2) into scratch, then the
  // exchange kernel; the scratch was reserved by vs_exchange_create
  o.want_fused = false;
  o.done_flag = nullptr;
  for (int b0 = 0; b0 < B; b0 += ix->xc.Bmax) {
    const int nb = B - b0 < ix->xc.Bmax ? B - b0 : ix->xc.Bmax;
    rc = query_dev_locked(ix, q_dev + (size_t)b0 * ix->dim, nb, k, req, (mode & ~0xff) | path, (float*)ix->d_xs.p,
                          (int64_t*)ix->d_xr.p, st, &o);
    if (rc) return rc;
    rc = exchange_chunks(ix, (const float*)ix->d_xs.p, (const int64_t*)ix->d_xr.p, nb, k, out_s + (size_t)b0 * k,
                         out_r + (size_t)b0 * k, st);
    if (rc) return rc;
  }
  return VS_OK;
}

int exchange_create_locked(vs_index* ix, int G, int rank, int B_max, int k_max) {
  const size_t bytes = exchange_bytes(B_max, k_max, G);
  // plain cudaMalloc (not a pool allocation): required for cudaIpcGetMemHandle
  CU(cudaMalloc(&ix->xc.local, bytes));
  CU(cudaMemset(ix->xc.local, 0, bytes));
  if (!ix->xc.h_err) {
    // the kernels raise this word on a timed-out wait; mapped + portable so every host entry point
    // (and every device of a single-process group) can reach it without a CUDA call
    CU(cudaHostAlloc((void**)&ix->xc.h_err, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(ix->xc.h_err, 0, 64);
  }
  // no allocation on the sharded query path: a cudaFree there would synchronise the device while a
  // peer may be spinning on this rank's push.  Reserve every scratch buffer for (B_max, k_max) now.
  CU(ix->d_xs.reserve((size_t)B_max * k_max * sizeof(float)));
  CU(ix->d_xr.reserve((size_t)B_max * k_max * sizeof(int64_t)));
  CU(ix->d_part_s.reserve((size_t)kScanBatch * ix->sm_count * kScanMaxCtasPerSm * k_max * sizeof(float)));
  CU(ix->d_part_r.reserve((size_t)kScanBatch * ix->sm_count * kScanMaxCtasPerSm * k_max * sizeof(uint32_t)));
  if (ix->tickets_n < (size_t)kScanBatch) {
    CU(ix->d_tickets.reserve(kScanBatch * sizeof(unsigned int)));
    CU(cudaMemset(ix->d_tickets.p, 0, ix->d_tickets.bytes));
    ix->tickets_n = kScanBatch;
  }
  if (ix->dtype == VS_BF16 && tensor_dim_ok(ix->dim)) {
    const int kt = k_max < kMaxTensorK ? k_max : kMaxTensorK;
    CU(ix->d_tensor.reserve(tensor_workspace_bytes(B_max, ix->dim, kt, ix->sm_count)));
  }
  CU(preload_exchange_kernels());
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

}  // namespace vs

using namespace vs;

extern "C" {

const char* vs_last_error(void) { return vs::g_err; }
int vs_abi_version(void) { return 2; }
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
    int rc = grow_locked(ix, capacity_rows);
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
  if (ix->last_stream_valid && ix->last_stream != ix->stream) cudaStreamSynchronize(ix->last_stream);
  if (ix->rows) cudaFree(ix->rows);
  if (ix->inv) cudaFree(ix->inv);
  if (ix->gmin) cudaFree(ix->gmin);
  if (ix->mask) cudaFree(ix->mask);
  DevBuf* bufs[] = {&ix->d_q,      &ix->d_out_s,  &ix->d_out_r,  &ix->d_part_s, &ix->d_part_r, &ix->d_tickets,
                    &ix->d_scores, &ix->d_select, &ix->d_tensor, &ix->d_stage,  &ix->d_misc,   &ix->d_xs,
                    &ix->d_xr};
  for (DevBuf* b : bufs) b->release();
  ix->h_in.release();
  ix->h_out.release();
  for (int p = 0; p < kMaxPeers; ++p)
    if (ix->xc.ipc_opened[p]) cudaIpcCloseMemHandle(ix->xc.peers[p]);
  if (ix->xc.local) cudaFree(ix->xc.local);
  if (ix->xc.h_err) cudaFreeHost(ix->xc.h_err);
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
int vs_device(const vs_index_t* ix) { return ix ? ix->device : -1; }

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

int vs_reserve(vs_index_t* ix, int64_t capacity_rows) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  return grow_locked(ix, capacity_rows);
}

int vs_add_dev(vs_index_t* ix, const float* rows_dev, int64_t n, int64_t* first_row, void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (n < 0 || (n > 0 && !rows_dev)) return fail(VS_ERR_ARG, "bad rows/n");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  return add_dev_locked(ix, rows_dev, n, first_row, pick_stream(ix, stream));
}

int vs_add_host(vs_index_t* ix, const float* rows, int64_t n, int64_t* first_row) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (n < 0 || (n > 0 && !rows)) return fail(VS_ERR_ARG, "bad rows/n");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  if (first_row) *first_row = ix->n;
  if (n == 0) return VS_OK;
  // stage through a device buffer in chunks of <= 64 MiB
  const int64_t chunk_rows = (64LL << 20) / ((int64_t)ix->dim * 4) > 0 ? (64LL << 20) / ((int64_t)ix->dim * 4) : 1;
  CU(ix->d_stage.reserve((size_t)(n < chunk_rows ? n : chunk_rows) * ix->dim * 4));
  for (int64_t r0 = 0; r0 < n; r0 += chunk_rows) {
    const int64_t nr = n - r0 < chunk_rows ? n - r0 : chunk_rows;
    CU(cudaMemcpyAsync(ix->d_stage.p, rows + (size_t)r0 * ix->dim, (size_t)nr * ix->dim * 4, cudaMemcpyHostToDevice,
                       ix->stream));
    int rc = add_dev_locked(ix, (const float*)ix->d_stage.p, nr, nullptr, ix->stream);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ix->stream));   // the staging buffer is reused by the next chunk
  }
  return VS_OK;
}

int vs_add_raw_host(vs_index_t* ix, const void* stored_rows, int64_t n, int64_t* first_row) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (n < 0 || (n > 0 && !stored_rows)) return fail(VS_ERR_ARG, "bad rows/n");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  if (first_row) *first_row = ix->n;
  if (n == 0) return VS_OK;
  if (ix->n + n > 0xFFFFFFF0LL) return fail(VS_ERR_ARG, "shard would exceed 2^32 rows");
  int rc = grow_locked(ix, ix->n + n);
  if (rc) return rc;
  CU(order_after_last(ix, ix->stream));
  // Rows already in the storage dtype (the persistence slab): chunked H2D through two pinned staging
  // buffers so the host-side memcpy of chunk c+1 overlaps the DMA of chunk c; the inverse norms are
  // recomputed on the device from the stored rows (identical to the ones ingest produced).
  const size_t row_bytes = (size_t)ix->dim * ix->esize, pitch = (size_t)ix->ld * ix->esize;
  const int64_t chunk_rows = std::max<int64_t>(1, (32LL << 20) / (int64_t)row_bytes);
  CU(ix->h_in.reserve(2 * (size_t)chunk_rows * row_bytes));
  cudaEvent_t evs[2] = {nullptr, nullptr};
  CU(cudaEventCreateWithFlags(&evs[0], cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&evs[1], cudaEventDisableTiming));
  int rc2 = VS_OK;
  int64_t c = 0;
  for (int64_t r0 = 0; r0 < n && rc2 == VS_OK; r0 += chunk_rows, ++c) {
    const int64_t nr = std::min(chunk_rows, n - r0);
    char* hb = (char*)ix->h_in.p + (size_t)(c & 1) * chunk_rows * row_bytes;
    if (c >= 2 && cudaEventSynchronize(evs[c & 1]) != cudaSuccess) rc2 = fail(VS_ERR_CUDA, "event sync failed");
    memcpy(hb, (const char*)stored_rows + (size_t)r0 * row_bytes, (size_t)nr * row_bytes);
    char* dst = (char*)ix->rows + (size_t)(ix->n + r0) * pitch;
    cudaError_t e = cudaMemcpy2DAsync(dst, pitch, hb, row_bytes, row_bytes, (size_t)nr, cudaMemcpyHostToDevice, ix->stream);
    if (e == cudaSuccess && pitch > row_bytes)
      e = cudaMemset2DAsync(dst + row_bytes, pitch, 0, pitch - row_bytes, (size_t)nr, ix->stream);
    if (e == cudaSuccess) e = cudaEventRecord(evs[c & 1], ix->stream);
    if (e != cudaSuccess) rc2 = fail(VS_ERR_CUDA, "raw row upload: %s", cudaGetErrorString(e));
  }
  if (rc2 == VS_OK) {
    cudaError_t e = launch_renorm(ix->rows, ix->n, n, ix->dim, ix->dtype, ix->ld, ix->inv, ix->stream);
    if (e == cudaSuccess && ix->mask)
      e = cudaMemsetAsync(ix->mask + (size_t)ix->n * kMaskWords, 0, (size_t)n * kMaskWords * 8, ix->stream);
    if (e != cudaSuccess) rc2 = fail(VS_ERR_CUDA, "renorm: %s", cudaGetErrorString(e));
  }
  if (rc2 == VS_OK) {
    const int64_t old_n = ix->n;
    ix->n += n;
    touch(ix);
    if (refresh_gmin(ix, old_n, ix->n, ix->stream) != cudaSuccess) rc2 = fail(VS_ERR_CUDA, "group bounds refresh failed");
  }
  cudaStreamSynchronize(ix->stream);
  cudaEventDestroy(evs[0]);
  cudaEventDestroy(evs[1]);
  cudaGetLastError();
  return rc2;
}

int vs_get_raw_host(const vs_index_t* cix, int64_t first_row, int64_t n, void* out) {
  vs_index* ix = const_cast<vs_index*>(cix);
  if (!ix || (!out && n > 0)) return fail(VS_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (first_row < 0 || n < 0 || first_row + n > ix->n) return fail(VS_ERR_ARG, "row range out of bounds");
  if (n == 0) return VS_OK;
  DeviceGuard g(ix->device);
  CU(order_after_last(ix, ix->stream));
  const size_t row_bytes = (size_t)ix->dim * ix->esize, pitch = (size_t)ix->ld * ix->esize;
  const char* src = (const char*)ix->rows + (size_t)first_row * pitch;
  CU(cudaMemcpy2DAsync(out, row_bytes, src, pitch, row_bytes, (size_t)n, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_remove(vs_index_t* ix, int64_t row, int64_t* moved_from) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (row < 0 || row >= ix->n) return fail(VS_ERR_ARG, "row %lld out of range [0,%lld)", (long long)row, (long long)ix->n);
  DeviceGuard g(ix->device);
  CU(order_after_last(ix, ix->stream));   // a query still in flight on a caller stream must not see a half-moved row
  const int64_t last = ix->n - 1;
  if (moved_from) *moved_from = row == last ? -1 : last;
  if (row != last) {
    const size_t rb = (size_t)ix->ld * ix->esize;
    CU(cudaMemcpyAsync((char*)ix->rows + row * rb, (char*)ix->rows + last * rb, rb, cudaMemcpyDeviceToDevice, ix->stream));
    CU(cudaMemcpyAsync(ix->inv + row, ix->inv + last, sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
    if (ix->mask)
      CU(cudaMemcpyAsync(ix->mask + row * kMaskWords, ix->mask + last * kMaskWords, kMaskWords * 8,
                         cudaMemcpyDeviceToDevice, ix->stream));
  }
  if (ix->mask) CU(cudaMemsetAsync(ix->mask + last * kMaskWords, 0, kMaskWords * 8, ix->stream));
  ix->n = last;
  touch(ix);
  if (row != last) CU(refresh_gmin(ix, row, row + 1, ix->stream));
  CU(refresh_gmin(ix, last > 0 ? last - 1 : 0, last + 1, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_remove_rows(vs_index_t* ix, const int64_t* rows, int64_t m, int64_t* moved_src, int64_t* moved_dst, int64_t* n_moved) {
  if (!ix || (m > 0 && !rows) || !n_moved) return fail(VS_ERR_ARG, "NULL argument");
  *n_moved = 0;
  if (m <= 0) return VS_OK;
  if (!moved_src || !moved_dst) return fail(VS_ERR_ARG, "moved_src / moved_dst must hold m entries");
  std::lock_guard<std::mutex> lk(ix->mu);
  std::vector<int64_t> del(rows, rows + m);
  std::sort(del.begin(), del.end());
  if (std::adjacent_find(del.begin(), del.end()) != del.end()) return fail(VS_ERR_ARG, "duplicate row in the delete list");
  if (del.front() < 0 || del.back() >= ix->n) return fail(VS_ERR_ARG, "row out of range [0,%lld)", (long long)ix->n);
  DeviceGuard g(ix->device);
  CU(order_after_last(ix, ix->stream));
  // compaction plan: the holes below the new count are filled, in ascending order, by the surviving rows at
  // or above it -- sources and destinations are disjoint, so one kernel moves them all
  const int64_t new_n = ix->n - m;
  std::vector<int64_t> pairs;
  {
    size_t di = std::lower_bound(del.begin(), del.end(), new_n) - del.begin();   // deleted rows >= new_n start here
    int64_t src = new_n;
    for (size_t h = 0; h < del.size() && del[h] < new_n; ++h) {
      while (di < del.size() && del[di] == src) {
        ++di;
        ++src;
      }
      pairs.push_back(src);
      pairs.push_back(del[h]);
      ++src;
    }
  }
  const int64_t np = (int64_t)pairs.size() / 2;
  if (np > 0) {
    CU(ix->d_misc.reserve((size_t)np * 16));
    CU(cudaMemcpyAsync(ix->d_misc.p, pairs.data(), (size_t)np * 16, cudaMemcpyHostToDevice, ix->stream));
    CU(launch_move_rows(ix->rows, ix->inv, ix->mask, ix->rows, ix->inv, ix->mask, (const int64_t*)ix->d_misc.p, np,
                        ix->ld * ix->esize, ix->stream));
  }
  if (ix->mask) CU(cudaMemsetAsync(ix->mask + new_n * kMaskWords, 0, (size_t)m * kMaskWords * 8, ix->stream));
  ix->n = new_n;
  touch(ix);
  CU(refresh_gmin(ix, np > 0 ? pairs[1] : (new_n > 0 ? new_n - 1 : 0), ix->n + 1, ix->stream));   // first hole .. (new) tail
  CU(cudaStreamSynchronize(ix->stream));
  for (int64_t i = 0; i < np; ++i) {
    moved_src[i] = pairs[2 * i];
    moved_dst[i] = pairs[2 * i + 1];
  }
  *n_moved = np;
  return VS_OK;
}

int vs_set_row_host(vs_index_t* ix, int64_t row, const float* vec) {
  if (!ix || !vec) return fail(VS_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (row < 0 || row >= ix->n) return fail(VS_ERR_ARG, "row %lld out of range [0,%lld)", (long long)row, (long long)ix->n);
  DeviceGuard g(ix->device);
  CU(order_after_last(ix, ix->stream));
  CU(ix->d_stage.reserve((size_t)ix->dim * 4));
  CU(cudaMemcpyAsync(ix->d_stage.p, vec, (size_t)ix->dim * 4, cudaMemcpyHostToDevice, ix->stream));
  char* dst = (char*)ix->rows + (size_t)row * ix->ld * ix->esize;
  CU(launch_ingest((const float*)ix->d_stage.p, 1, ix->dim, ix->dtype, dst, ix->ld, ix->inv + row, ix->stream));
  touch(ix);
  CU(refresh_gmin(ix, row, row + 1, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_copy_row(vs_index_t* dst, int64_t dst_row, vs_index_t* src, int64_t src_row) {
  if (!dst || !src) return fail(VS_ERR_ARG, "NULL argument");
  if (dst->dim != src->dim || dst->dtype != src->dtype) return fail(VS_ERR_ARG, "indexes differ in dim / dtype");
  std::unique_lock<std::mutex> l1, l2;
  if (dst != src) {
    std::lock(dst->mu, src->mu);
    l1 = std::unique_lock<std::mutex>(dst->mu, std::adopt_lock);
    l2 = std::unique_lock<std::mutex>(src->mu, std::adopt_lock);
  } else {
    l1 = std::unique_lock<std::mutex>(dst->mu);
  }
  if (dst_row < 0 || dst_row >= dst->n || src_row < 0 || src_row >= src->n) return fail(VS_ERR_ARG, "row out of range");
  const size_t rb = (size_t)dst->ld * dst->esize;
  if (dst != src) {
    DeviceGuard gs(src->device);
    CU(cudaStreamSynchronize(src->stream));
    if (src->last_stream_valid && src->last_stream != src->stream) CU(cudaStreamSynchronize(src->last_stream));
  }
  DeviceGuard g(dst->device);
  CU(order_after_last(dst, dst->stream));
  CU(cudaMemcpyPeerAsync((char*)dst->rows + dst_row * rb, dst->device, (const char*)src->rows + src_row * rb, src->device, rb,
                         dst->stream));
  CU(cudaMemcpyPeerAsync(dst->inv + dst_row, dst->device, src->inv + src_row, src->device, sizeof(float), dst->stream));
  if (src->mask) {
    int rc = ensure_mask(dst);
    if (rc) return rc;
    CU(cudaMemcpyPeerAsync(dst->mask + dst_row * kMaskWords, dst->device, src->mask + src_row * kMaskWords, src->device,
                           kMaskWords * 8, dst->stream));
  } else if (dst->mask) {
    CU(cudaMemsetAsync(dst->mask + dst_row * kMaskWords, 0, kMaskWords * 8, dst->stream));
  }
  touch(dst);
  CU(refresh_gmin(dst, dst_row, dst_row + 1, dst->stream));
  CU(cudaStreamSynchronize(dst->stream));
  return VS_OK;
}

int vs_move_rows(vs_index_t* dst, vs_index_t* src, const int64_t* src_rows, const int64_t* dst_rows, int64_t n) {
  if (!dst || !src || (n > 0 && (!src_rows || !dst_rows))) return fail(VS_ERR_ARG, "NULL argument");
  if (dst->dim != src->dim || dst->dtype != src->dtype) return fail(VS_ERR_ARG, "indexes differ in dim / dtype");
  if (n <= 0) return VS_OK;
  std::unique_lock<std::mutex> l1, l2;
  if (dst != src) {
    std::lock(dst->mu, src->mu);
    l1 = std::unique_lock<std::mutex>(dst->mu, std::adopt_lock);
    l2 = std::unique_lock<std::mutex>(src->mu, std::adopt_lock);
  } else {
    l1 = std::unique_lock<std::mutex>(dst->mu);
  }
  std::vector<int64_t> pairs((size_t)2 * n);
  int64_t lo = dst->n;
  for (int64_t i = 0; i < n; ++i) {
    if (src_rows[i] < 0 || src_rows[i] >= src->n || dst_rows[i] < 0 || dst_rows[i] >= dst->n) return fail(VS_ERR_ARG, "row out of range");
    pairs[2 * i] = src_rows[i];
    pairs[2 * i + 1] = dst_rows[i];
    lo = std::min(lo, dst_rows[i]);
  }
  if (dst != src) {
    DeviceGuard gs(src->device);
    CU(cudaStreamSynchronize(src->stream));
    if (src->last_stream_valid && src->last_stream != src->stream) CU(cudaStreamSynchronize(src->last_stream));
  }
  DeviceGuard g(dst->device);
  if (src->device != dst->device) {
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, dst->device, src->device));
    if (!can) return fail(VS_ERR_UNSUPPORTED, "device %d cannot access peer device %d", dst->device, src->device);
    cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
    cudaGetLastError();
  }
  if (src->mask) {
    int rc = ensure_mask(dst);
    if (rc) return rc;
  }
  CU(order_after_last(dst, dst->stream));
  CU(dst->d_misc.reserve((size_t)n * 16));
  CU(cudaMemcpyAsync(dst->d_misc.p, pairs.data(), (size_t)n * 16, cudaMemcpyHostToDevice, dst->stream));
  CU(launch_move_rows(src->rows, src->inv, src->mask, dst->rows, dst->inv, dst->mask, (const int64_t*)dst->d_misc.p, n,
                      dst->ld * dst->esize, dst->stream));
  touch(dst);
  CU(refresh_gmin(dst, lo, dst->n, dst->stream));
  CU(cudaStreamSynchronize(dst->stream));
  return VS_OK;
}

int vs_truncate(vs_index_t* ix, int64_t new_count) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (new_count < 0 || new_count > ix->n) return fail(VS_ERR_ARG, "new_count %lld out of range [0,%lld]", (long long)new_count, (long long)ix->n);
  if (new_count == ix->n) return VS_OK;
  DeviceGuard g(ix->device);
  CU(order_after_last(ix, ix->stream));
  if (ix->mask) CU(cudaMemsetAsync(ix->mask + new_count * kMaskWords, 0, (size_t)(ix->n - new_count) * kMaskWords * 8, ix->stream));
  ix->n = new_count;
  touch(ix);
  CU(refresh_gmin(ix, new_count > 0 ? new_count - 1 : 0, new_count + 1, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_clear(vs_index_t* ix) { return vs_truncate(ix, 0); }

int vs_replicate_from(vs_index_t* dst, vs_index_t* src, int64_t dst_first, int64_t dst_stride) {
  if (!dst || !src || dst == src) return fail(VS_ERR_ARG, "need two different indexes");
  if (dst->dim != src->dim || dst->dtype != src->dtype) return fail(VS_ERR_ARG, "indexes differ in dim / dtype");
  if (dst_first < 0 || dst_stride < 1) return fail(VS_ERR_ARG, "bad dst_first / dst_stride");
  std::lock(dst->mu, src->mu);
  std::unique_lock<std::mutex> l1(dst->mu, std::adopt_lock), l2(src->mu, std::adopt_lock);
  const int64_t n = src->n;
  if (n == 0) return VS_OK;
  const int64_t need = dst_first + (n - 1) * dst_stride + 1;
  {
    DeviceGuard gs(src->device);
    CU(cudaStreamSynchronize(src->stream));
    if (src->last_stream_valid && src->last_stream != src->stream) CU(cudaStreamSynchronize(src->last_stream));
  }
  DeviceGuard g(dst->device);
  if (src->device != dst->device) {
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, dst->device, src->device));
    if (!can) return fail(VS_ERR_UNSUPPORTED, "device %d cannot access peer device %d", dst->device, src->device);
    cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
    cudaGetLastError();
  }
  int rc = grow_locked(dst, need);
  if (rc) return rc;
  CU(order_after_last(dst, dst->stream));
  CU(launch_strided_copy(src->rows, src->inv, dst->rows, dst->inv, n, dst_first, dst_stride, dst->ld * dst->esize, dst->stream));
  if (need > dst->n) dst->n = need;
  touch(dst);
  CU(refresh_gmin(dst, 0, dst->n, dst->stream));
  CU(cudaStreamSynchronize(dst->stream));
  return VS_OK;
}

int vs_set_mask_bits(vs_index_t* ix, int64_t row, const uint64_t bits[VS_MASK_WORDS]) {
  if (!ix || !bits) return fail(VS_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (row < 0 || row >= ix->n) return fail(VS_ERR_ARG, "row %lld out of range", (long long)row);
  DeviceGuard g(ix->device);
  int rc = ensure_mask(ix);
  if (rc) return rc;
  CU(order_after_last(ix, ix->stream));
  CU(cudaMemcpyAsync(ix->mask + row * kMaskWords, bits, kMaskWords * 8, cudaMemcpyHostToDevice, ix->stream));
  touch(ix);
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
  CU(order_after_last(ix, ix->stream));
  CU(cudaMemcpyAsync(ix->mask + first_row * kMaskWords, bits, (size_t)n * kMaskWords * 8, cudaMemcpyHostToDevice,
                     ix->stream));
  touch(ix);
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_get_mask_bits(const vs_index_t* cix, int64_t row, uint64_t bits[VS_MASK_WORDS]) {
  vs_index* ix = const_cast<vs_index*>(cix);
  if (!ix || !bits) return fail(VS_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (row < 0 || row >= ix->n) return fail(VS_ERR_ARG, "row %lld out of range", (long long)row);
  memset(bits, 0, kMaskWords * 8);
  if (!ix->mask) return VS_OK;
  DeviceGuard g(ix->device);
  CU(cudaMemcpyAsync(bits, ix->mask + row * kMaskWords, kMaskWords * 8, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_get_mask_bits_range(const vs_index_t* cix, int64_t first_row, int64_t n, uint64_t* bits) {
  vs_index* ix = const_cast<vs_index*>(cix);
  if (!ix || (!bits && n > 0)) return fail(VS_ERR_ARG, "NULL argument");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (first_row < 0 || n < 0 || first_row + n > ix->n) return fail(VS_ERR_ARG, "row range out of bounds");
  if (n == 0) return VS_OK;
  memset(bits, 0, (size_t)n * kMaskWords * 8);
  if (!ix->mask) return VS_OK;
  DeviceGuard g(ix->device);
  CU(cudaMemcpyAsync(bits, ix->mask + first_row * kMaskWords, (size_t)n * kMaskWords * 8, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

int vs_apply_sweep_bits_dev(vs_index_t* ix, const uint32_t* words_dev, int bit, void* stream) {
  if (!ix || !words_dev) return fail(VS_ERR_ARG, "NULL argument");
  if (bit < 0 || bit >= 64 * kMaskWords) return fail(VS_ERR_ARG, "filter bit %d out of range [0,%d)", bit, 64 * kMaskWords);
  std::lock_guard<std::mutex> lk(ix->mu);
  if (ix->n == 0) return VS_OK;
  DeviceGuard g(ix->device);
  int rc = ensure_mask(ix);
  if (rc) return rc;
  cudaStream_t st = pick_stream(ix, stream);
  CU(order_after_last(ix, st));
  CU(launch_apply_sweep_bits(words_dev, ix->n, bit, ix->mask, st));
  touch(ix);
  if (st != ix->stream) {
    CU(cudaEventRecord(ix->ev, st));
    CU(cudaStreamWaitEvent(ix->stream, ix->ev, 0));
  }
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
  CU(launch_export(src, n, ix->dim, ix->dtype, ix->ld, (float*)ix->d_stage.p, ix->stream));
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
  CU(launch_export(src, n, ix->dim, ix->dtype, ix->ld, out_dev, pick_stream(ix, stream)));
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

// D2H of a [B,k] result that sits as [scores | rows] in d_out_s, one copy + sync
static int fetch_result(vs_index* ix, int B, int k, float* out_scores, int64_t* out_rows) {
  const size_t sbytes = (size_t)B * k * 4, rbytes = (size_t)B * k * 8, roff = (sbytes + 15) & ~(size_t)15;
  CU(cudaMemcpyAsync(ix->h_out.p, ix->d_out_s.p, roff + rbytes, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  memcpy(out_scores, ix->h_out.p, sbytes);
  memcpy(out_rows, (char*)ix->h_out.p + roff, rbytes);
  return VS_OK;
}
static int reserve_result(vs_index* ix, int B, int k, float** d_s, int64_t** d_r) {
  const size_t sbytes = (size_t)B * k * 4, rbytes = (size_t)B * k * 8, roff = (sbytes + 15) & ~(size_t)15;
  CU(ix->d_out_s.reserve(roff + rbytes + 64));   // scores | rows in ONE buffer: one D2H
  CU(ix->h_out.reserve(roff + rbytes + 64));
  *d_s = (float*)ix->d_out_s.p;
  *d_r = (int64_t*)((char*)ix->d_out_s.p + roff);
  return VS_OK;
}

int vs_query_topk_host(vs_index_t* ix, const float* q, int B, int k, const uint64_t* require_bits, int mode,
                       float* out_scores, int64_t* out_rows) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B < 0 || (B > 0 && (!q || !out_scores || !out_rows))) return fail(VS_ERR_ARG, "NULL buffer");
  if (B == 0) return VS_OK;
  if (k <= 0 || k > kMaxK) return fail(VS_ERR_ARG, "k=%d out of range [1,%d]", k, kMaxK);
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  const size_t qbytes = (size_t)B * ix->dim * 4;
  float* d_s;
  int64_t* d_r;
  int rc = reserve_result(ix, B, k, &d_s, &d_r);
  if (rc) return rc;
  CU(ix->d_q.reserve(qbytes));
  CU(ix->h_in.reserve(qbytes));
  memcpy(ix->h_in.p, q, qbytes);  // pinned staging so the H2D copy is a true async DMA
  CU(cudaMemcpyAsync(ix->d_q.p, ix->h_in.p, qbytes, cudaMemcpyHostToDevice, ix->stream));
  rc = query_dev_locked(ix, (const float*)ix->d_q.p, B, k, require_bits, mode & 0xff, d_s, d_r, ix->stream);
  if (rc) return rc;
  return fetch_result(ix, B, k, out_scores, out_rows);
}

int vs_blend_dev(vs_index_t* ix, const float* img_dev, const float* txt_dev, const double* w_dev, int B, float* out_dev,
                 void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B < 0 || (B > 0 && (!img_dev || !txt_dev || !w_dev || !out_dev))) return fail(VS_ERR_ARG, "NULL buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  cudaStream_t st = pick_stream(ix, stream);
  if (ix->chain_stream == st) touch(ix);   // the next scan on this stream follows a non-scan kernel that writes its q
  CU(launch_blend(img_dev, txt_dev, w_dev, B, ix->dim, out_dev, st));
  return VS_OK;
}

int vs_query_multimodal_host(vs_index_t* ix, const float* img, const float* txt, const double* w, int B, int k,
                             const uint64_t* require_bits, int mode, float* out_scores, int64_t* out_rows) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B < 0 || (B > 0 && (!img || !txt || !w || !out_scores || !out_rows))) return fail(VS_ERR_ARG, "NULL buffer");
  if (B == 0) return VS_OK;
  if (k <= 0 || k > kMaxK) return fail(VS_ERR_ARG, "k=%d out of range [1,%d]", k, kMaxK);
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  float* d_s;
  int64_t* d_r;
  int rc = reserve_result(ix, B, k, &d_s, &d_r);
  if (rc) return rc;
  // stage (img | txt | weights) on the device and blend into the tail of the same buffer
  const size_t qbytes = (size_t)B * ix->dim * 4, wbytes = (size_t)B * 8;
  const size_t boff = (2 * qbytes + wbytes + 15) & ~(size_t)15;
  CU(ix->d_stage.reserve(boff + qbytes));
  CU(ix->h_in.reserve(2 * qbytes + wbytes));
  char* hin = (char*)ix->h_in.p;
  memcpy(hin, img, qbytes);
  memcpy(hin + qbytes, txt, qbytes);
  memcpy(hin + 2 * qbytes, w, wbytes);
  char* dst = (char*)ix->d_stage.p;
  CU(cudaMemcpyAsync(dst, hin, 2 * qbytes + wbytes, cudaMemcpyHostToDevice, ix->stream));
  float* blended = (float*)(dst + boff);
  touch(ix);
  CU(launch_blend((const float*)dst, (const float*)(dst + qbytes), (const double*)(dst + 2 * qbytes), B, ix->dim, blended,
                  ix->stream));
  rc = query_dev_locked(ix, blended, B, k, require_bits, mode & 0xff, d_s, d_r, ix->stream);
  if (rc) return rc;
  return fetch_result(ix, B, k, out_scores, out_rows);
}

int vs_merge_topk_dev(vs_index_t* ix, const float* cand_scores_dev, const int64_t* cand_rows_dev, int G, int B, int k,
                      float* out_scores_dev, int64_t* out_rows_dev, void* stream) {
  if (!cand_scores_dev || !cand_rows_dev || !out_scores_dev || !out_rows_dev) return fail(VS_ERR_ARG, "NULL buffer");
  if (G <= 0 || B <= 0 || k <= 0 || (int64_t)G * k > 16384) return fail(VS_ERR_ARG, "G*k must be in [1,16384]");
  cudaStream_t st = stream ? (cudaStream_t)stream : (ix ? ix->stream : (cudaStream_t) nullptr);
  if (ix) {
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    if (ix->chain_stream == st) touch(ix);
    CU(launch_merge(cand_scores_dev, cand_rows_dev, G, B, k, out_scores_dev, out_rows_dev, st));
  } else {
    CU(launch_merge(cand_scores_dev, cand_rows_dev, G, B, k, out_scores_dev, out_rows_dev, st));
  }
  return VS_OK;
}

size_t vs_exchange_bytes(int B_max, int k_max, int G) {
  if (B_max <= 0 || k_max <= 0 || G <= 0 || G > kMaxPeers) return 0;
  return exchange_bytes(B_max, k_max, G);
}

int vs_exchange_create(vs_index_t* ix, int G, int rank, int B_max, int k_max) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (G <= 0 || G > kMaxPeers || rank < 0 || rank >= G)
    return fail(VS_ERR_ARG, "need 1 <= G <= %d and 0 <= rank < G", kMaxPeers);
  if (B_max <= 0 || B_max > 65536 || k_max <= 0 || k_max > kMaxFusedK)
    return fail(VS_ERR_ARG, "need 1 <= B_max <= 65536 and 1 <= k_max <= %d", kMaxFusedK);
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

int vs_query_topk_sharded_dev(vs_index_t* ix, const float* q_dev, int B, int k, const uint64_t* require_bits, int mode,
                              float* out_scores_dev, int64_t* out_rows_dev, void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B <= 0 || !q_dev || !out_scores_dev || !out_rows_dev) return fail(VS_ERR_ARG, "bad B or NULL buffer");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  return sharded_query_dev_locked(ix, q_dev, B, k, require_bits, mode, out_scores_dev, out_rows_dev, pick_stream(ix, stream));
}

int vs_query_topk_sharded_host(vs_index_t* ix, const float* q, int B, int k, const uint64_t* require_bits, int mode,
                               float* out_scores, int64_t* out_rows) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (B <= 0 || !q || !out_scores || !out_rows) return fail(VS_ERR_ARG, "bad B or NULL buffer");
  if (k <= 0 || k > kMaxK) return fail(VS_ERR_ARG, "k=%d out of range [1,%d]", k, kMaxK);
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  const size_t qbytes = (size_t)B * ix->dim * 4;
  float* d_s;
  int64_t* d_r;
  int rc = reserve_result(ix, B, k, &d_s, &d_r);
  if (rc) return rc;
  CU(ix->d_q.reserve(qbytes));
  CU(ix->h_in.reserve(qbytes));
  memcpy(ix->h_in.p, q, qbytes);
  CU(cudaMemcpyAsync(ix->d_q.p, ix->h_in.p, qbytes, cudaMemcpyHostToDevice, ix->stream));
  rc = sharded_query_dev_locked(ix, (const float*)ix->d_q.p, B, k, require_bits, mode & 0xff, d_s, d_r, ix->stream);
  if (rc) return rc;
  rc = fetch_result(ix, B, k, out_scores, out_rows);
  if (rc) return rc;
  return exchange_error_locked(ix, true);   // a peer never arrived: the rows came back empty, say so
}

int vs_exchange_begin(vs_index_t* ix) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (!ix->xc.local || !ix->xc.attached) return fail(VS_ERR_ARG, "exchange not created/attached");
  int rc = exchange_error_locked(ix, false);
  if (rc) return rc;
  ix->xc.epoch = next_epoch(ix->xc.epoch);
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
  const int path = pick_path(ix, mode & 0xff, B, k);
  float* xs = (float*)ix->d_xs.p;
  int64_t* xr = (int64_t*)ix->d_xr.p;
  QueryOpts o;
  if (path == VS_Q_SCAN && B <= kScanBatch && k <= kMaxFusedK) {
    o.want_fused = true;
    o.push_slot0 = slot0;
  }
  rc = query_dev_locked(ix, q_dev, B, k, require_bits, (mode & ~0xff) | path, xs, xr, st, &o);
  if (rc) return rc;
  if (o.fused) return VS_OK;          // the scan kernel's last CTA pushed the candidates
  touch(ix);
  CU(launch_exchange_merge(xs, xr, make_exchange(ix, slot0, false, true), B, k, nullptr, nullptr, ix->sm_count, st, 1));
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
  CU(order_after_last(ix, st));
  touch(ix);
  CU(launch_exchange_merge(nullptr, nullptr, make_exchange(ix, 0, false, false), B, k, out_scores_dev, out_rows_dev,
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
  CU(order_after_last(ix, st));
  return exchange_chunks(ix, cand_scores_dev, cand_rows_dev, B, k, out_scores_dev, out_rows_dev, st);
}

int vs_exchange_error(vs_index_t* ix) {
  if (!ix || !ix->xc.h_err) return 0;
  return (int)*(volatile unsigned int*)ix->xc.h_err;
}

int vs_exchange_clear_error(vs_index_t* ix) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (ix->xc.h_err) *(volatile unsigned int*)ix->xc.h_err = 0;
  return VS_OK;
}

int64_t vs_filter_words(const vs_index_t* ix) {
  if (!ix) return 0;
  return (ix->n + 255) / 256 * 8;   // whole 256-row tiles: the sweep writes full tiles
}

static int filter_sweep_dev_locked(vs_index* ix, const float* prompts_dev, int F, float tau, uint32_t* out_bits_dev, cudaStream_t st) {
  if (ix->dtype != VS_BF16) return fail(VS_ERR_UNSUPPORTED, "filter sweep needs bf16 storage (tcgen05 path)");
  if (!tensor_dim_ok(ix->dim)) return fail(VS_ERR_UNSUPPORTED, "tensor path needs dim %% 8 == 0 and dim <= 4096");
  if (ix->n == 0) return VS_OK;
  const TensorArgs ta = tensor_args(ix, nullptr, false);
  CU(ix->d_tensor.reserve(tensor_workspace_bytes(F, ix->dim, 1, ix->sm_count)));
  CU(order_after_last(ix, st));
  touch(ix);
  CU(launch_tensor_filter(ta, prompts_dev, F, tau, ix->d_tensor.p, out_bits_dev, vs_filter_words(ix), ix->sm_count, st));
  return VS_OK;
}

int vs_filter_sweep_dev(vs_index_t* ix, const float* prompts_dev, int F, float tau, uint32_t* out_bits_dev, void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (F <= 0 || !prompts_dev || !out_bits_dev) return fail(VS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  return filter_sweep_dev_locked(ix, prompts_dev, F, tau, out_bits_dev, pick_stream(ix, stream));
}

int vs_filter_sweep_host(vs_index_t* ix, const float* prompts, int F, float tau, uint32_t* out_bits) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (F <= 0 || !prompts || !out_bits) return fail(VS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  const size_t pbytes = (size_t)F * ix->dim * 4;
  const size_t obytes = (size_t)F * vs_filter_words(ix) * 4;
  CU(ix->d_q.reserve(pbytes));
  CU(ix->d_misc.reserve(obytes + 16));
  CU(cudaMemcpyAsync(ix->d_q.p, prompts, pbytes, cudaMemcpyHostToDevice, ix->stream));
  int rc = filter_sweep_dev_locked(ix, (const float*)ix->d_q.p, F, tau, (uint32_t*)ix->d_misc.p, ix->stream);
  if (rc) return rc;
  if (obytes) CU(cudaMemcpyAsync(out_bits, ix->d_misc.p, obytes, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return VS_OK;
}

static int dedup_dev_locked(vs_index* ix, int64_t row_lo, int64_t row_hi, float tau, int64_t cap, int64_t* out_i_dev,
                            int64_t* out_j_dev, float* out_score_dev, unsigned long long* out_count_dev, cudaStream_t st) {
  if (ix->dtype != VS_BF16) return fail(VS_ERR_UNSUPPORTED, "dedup needs bf16 storage (tcgen05 path)");
  if (!tensor_dim_ok(ix->dim)) return fail(VS_ERR_UNSUPPORTED, "tensor path needs dim %% 8 == 0 and dim <= 4096");
  if (row_lo < 0) row_lo = 0;
  if (row_hi > ix->n) row_hi = ix->n;
  CU(order_after_last(ix, st));
  touch(ix);
  CU(cudaMemsetAsync(out_count_dev, 0, sizeof(unsigned long long), st));
  if (row_lo >= row_hi) return VS_OK;
  const TensorArgs ta = tensor_args(ix, nullptr, false);
  CU(ix->d_tensor.reserve(tensor_workspace_bytes(128, ix->dim, 1, ix->sm_count)));
  CU(launch_tensor_dedup(ta, row_lo, row_hi, tau, cap, out_i_dev, out_j_dev, out_score_dev, out_count_dev, ix->d_tensor.p,
                         ix->sm_count, st));
  return VS_OK;
}

int vs_dedup_dev(vs_index_t* ix, int64_t row_lo, int64_t row_hi, float tau, int64_t cap, int64_t* out_i_dev,
                 int64_t* out_j_dev, float* out_score_dev, unsigned long long* out_count_dev, void* stream) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (!out_i_dev || !out_j_dev || !out_score_dev || !out_count_dev || cap < 0) return fail(VS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  return dedup_dev_locked(ix, row_lo, row_hi, tau, cap, out_i_dev, out_j_dev, out_score_dev, out_count_dev, pick_stream(ix, stream));
}

int vs_dedup_host(vs_index_t* ix, int64_t row_lo, int64_t row_hi, float tau, int64_t cap, int64_t* out_i, int64_t* out_j,
                  float* out_score, int64_t* out_count) {
  if (!ix) return fail(VS_ERR_ARG, "index is NULL");
  if (!out_i || !out_j || !out_score || !out_count || cap < 0) return fail(VS_ERR_ARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  CU(ix->d_misc.reserve((size_t)cap * 20 + 64));
  void* buf = ix->d_misc.p;
  unsigned long long* d_count = (unsigned long long*)buf;
  int64_t* d_i = (int64_t*)((char*)buf + 16);
  int64_t* d_j = d_i + cap;
  float* d_s = (float*)(d_j + cap);
  int rc = dedup_dev_locked(ix, row_lo, row_hi, tau, cap, d_i, d_j, d_s, d_count, ix->stream);
  if (rc) return rc;
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
