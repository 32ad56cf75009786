// group.cu -- vs_group_t: ONE host process driving the row shards on all GPUs of the box.
//
// The reference is a single uvicorn process whose routes call collection.query one request at a time
// (backend/run.py:10-14, backend/app/main.py:748-805).  A group is that deployment shape: G shards
// (global row g lives on shard g % G), one resident worker thread per GPU, and a request/response
// path that contains no stream synchronise and no result copy:
//   front thread : memcpy q into a pinned, portable, host-MAPPED request area; bump `seq`
//   worker g     : (spinning on `seq`) one small H2D of q into its GPU, ONE fused scan launch -- the
//                  kernel's last CTA pushes the shard's candidates to every peer over NVLink and merges
//                  (csrc/exchange.cuh); shard 0's kernel writes the global [k] result straight into
//                  host-mapped memory and raises a host-mapped flag (st.release.sys)
//   front thread : polls the flag, copies [B, k] out.
// Batches that do not fit the fused form (tcgen05 path, B > 64) run local query + exchange kernel with
// the result still written to the mapped area; k above the exchange's k_max (the UI's "All" = 1000)
// gathers the shards' candidates onto GPU 0 with peer copies and merges there (K5).
// Everything else a collection needs (add / remove / filter bits / sweep / dedup) goes through the
// per-shard vs_index_t handles returned by vs_group_shard().
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <new>
#include <thread>

#include "index_internal.h"

using namespace vs;
#define CU(call) VS_CU(call)

namespace {

struct Request {
  int B = 0, k = 0, mode = 0, blend = 0, gather = 0;
  bool use_bits = false;
  uint64_t bits[kMaskWords] = {0, 0, 0, 0};
};

constexpr double kFlagTimeoutUs = 30e6;   // see run_request

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}

}  // namespace

struct vs_group {
  int G = 0, dim = 0, dtype = 0, b_max = 0, k_max = 0;
  vs_index* ix[kMaxPeers] = {};
  // request / response area: pinned + portable + mapped (one allocation)
  char* h_area = nullptr;
  char* h_in = nullptr;            // [B][dim] f32, or [img | txt | w] for a blend request
  float* h_out_s = nullptr;        // [B][k]
  int64_t* h_out_r = nullptr;
  unsigned int* h_done = nullptr;  // [b_max]: == seq once query b's result is in h_out_*
  size_t in_bytes = 0, out_s_bytes = 0, out_r_bytes = 0;
  // large-k gather buffers on GPU 0: [G][B][k]
  DevBuf g_cand;
  cudaEvent_t ev[kMaxPeers] = {};
  // dispatch
  Request req;
  std::atomic<unsigned> seq{0};
  std::atomic<unsigned> acked[kMaxPeers];
  int rc[kMaxPeers] = {};
  char err[kMaxPeers][512] = {};
  std::thread workers[kMaxPeers];
  std::atomic<bool> stop{false};
  std::atomic<int> sleepers{0};
  std::mutex cv_mu;
  std::condition_variable cv;
  std::mutex front_mu;   // one request in flight
  // host-side timeline of the LAST request (microseconds since its entry): introspection for bench.py / tuning
  double t_submit_us = 0, t_enqueued_us = 0, t_done_us = 0, t_return_us = 0;
};

namespace {

// worker g: enqueue request `s` on its GPU.  Runs with ix->mu held.
int serve_locked(vs_group* gr, int g, unsigned s) {
  vs_index* ix = gr->ix[g];
  const Request& r = gr->req;
  const int B = r.B, k = r.k;
  const size_t qbytes = (size_t)B * gr->dim * 4;
  const float* q = nullptr;
  bool inline_q = false;
  if (r.blend) {
    const size_t wbytes = (size_t)B * 8, boff = (2 * qbytes + wbytes + 15) & ~(size_t)15;
    char* dst = (char*)ix->d_stage.p;   // [img | txt | w | blended], reserved at create for b_max
    CU(cudaMemcpyAsync(dst, gr->h_in, 2 * qbytes + wbytes, cudaMemcpyHostToDevice, ix->stream));
    float* blended = (float*)(dst + boff);
    touch(ix);
    CU(launch_blend((const float*)dst, (const float*)(dst + qbytes), (const double*)(dst + 2 * qbytes), B, gr->dim, blended,
                    ix->stream));
    q = blended;
  } else if (B == 1 && gr->dim <= kMaxInlineQ && !r.gather && r.mode != VS_Q_TENSOR && k <= kMaxFusedK) {
    inline_q = true;   // one request: the query rides in the scan kernel's launch packet, no H2D copy at all
  } else {
    CU(cudaMemcpyAsync(ix->d_q.p, gr->h_in, qbytes, cudaMemcpyHostToDevice, ix->stream));
    q = (const float*)ix->d_q.p;
  }
  const uint64_t* bits = r.use_bits ? r.bits : nullptr;
  if (r.gather) {
    // k above the exchange's k_max: local top-k into this shard's scratch; shard 0 gathers + merges
    const size_t sbytes = (size_t)B * k * 4, roff = (sbytes + 15) & ~(size_t)15;
    CU(ix->d_out_s.reserve(roff + (size_t)B * k * 8 + 64));
    int rc = query_dev_locked(ix, q, B, k, bits, r.mode, (float*)ix->d_out_s.p, (int64_t*)((char*)ix->d_out_s.p + roff),
                              ix->stream);
    if (rc) return rc;
    CU(cudaEventRecord(gr->ev[g], ix->stream));
    return VS_OK;
  }
  // shard 0 reports: its kernels write the global result into the host-mapped response area
  float* out_s = g == 0 ? gr->h_out_s : (float*)ix->d_out_s.p;
  int64_t* out_r = g == 0 ? gr->h_out_r : (int64_t*)((char*)ix->d_out_s.p + (((size_t)B * k * 4 + 15) & ~(size_t)15));
  QueryOpts o;
  if (g == 0) {
    o.done_flag = gr->h_done;
    o.done_seq = s;
  }
  if (inline_q) {
    o.q_host = (const float*)gr->h_in;
    q = (const float*)ix->d_q.p;   // only dereferenced if the shard is empty / the path cannot inline (it then holds stale data:
                                   // copy it after all so that every path stays correct)
    if (ix->n == 0) CU(cudaMemcpyAsync(ix->d_q.p, gr->h_in, qbytes, cudaMemcpyHostToDevice, ix->stream));
  }
  int rc = sharded_query_dev_locked(ix, q, B, k, bits, (r.mode == VS_Q_AUTO && inline_q) ? VS_Q_SCAN : r.mode, out_s, out_r,
                                    ix->stream, &o);
  if (rc) return rc;
  if (g == 0 && !o.done_armed) {
    // not the single-launch fused form: the last kernel of the chain wrote the mapped area; wait for it
    CU(cudaStreamSynchronize(ix->stream));
    std::atomic_thread_fence(std::memory_order_seq_cst);
    for (int b = 0; b < B; ++b) ((volatile unsigned int*)gr->h_done)[b] = s;
  }
  return VS_OK;
}

// shard 0, gather form: wait for every shard's local result, pull it over NVLink, merge (K5), report
int gather_locked(vs_group* gr, unsigned s) {
  vs_index* ix0 = gr->ix[0];
  const Request& r = gr->req;
  const int B = r.B, k = r.k, G = gr->G;
  const size_t sbytes = (size_t)B * k * 4, rbytes = (size_t)B * k * 8, roff = (sbytes + 15) & ~(size_t)15;
  const size_t soff_all = 0, roff_all = ((size_t)G * sbytes + 255) & ~(size_t)255;
  CU(gr->g_cand.reserve(roff_all + (size_t)G * rbytes));
  for (int g = 0; g < G; ++g) {
    while (gr->acked[g].load(std::memory_order_acquire) != s && g != 0) cpu_relax();   // shard g has enqueued (event recorded)
    if (g != 0 && gr->rc[g] != VS_OK) return fail(gr->rc[g], "%s", gr->err[g]);
    vs_index* ix = gr->ix[g];
    CU(cudaStreamWaitEvent(ix0->stream, gr->ev[g], 0));
    CU(cudaMemcpyPeerAsync((char*)gr->g_cand.p + soff_all + g * sbytes, ix0->device, ix->d_out_s.p, ix->device, sbytes, ix0->stream));
    CU(cudaMemcpyPeerAsync((char*)gr->g_cand.p + roff_all + g * rbytes, ix0->device, (char*)ix->d_out_s.p + roff, ix->device,
                           rbytes, ix0->stream));
  }
  touch(ix0);
  CU(launch_merge((const float*)((char*)gr->g_cand.p + soff_all), (const int64_t*)((char*)gr->g_cand.p + roff_all), G, B, k,
                  gr->h_out_s, gr->h_out_r, ix0->stream));
  CU(cudaStreamSynchronize(ix0->stream));
  std::atomic_thread_fence(std::memory_order_seq_cst);
  for (int b = 0; b < B; ++b) ((volatile unsigned int*)gr->h_done)[b] = s;
  return VS_OK;
}

void worker_main(vs_group* gr, int g) {
  vs_index* ix = gr->ix[g];
  cudaSetDevice(ix->device);
  unsigned seen = 0;
  for (;;) {
    // wait for the next request: spin ~1 ms (requests arrive back to back under load), then sleep
    unsigned s = gr->seq.load(std::memory_order_acquire);
    for (int spins = 0; s == seen && !gr->stop.load(std::memory_order_relaxed); ++spins) {
      if (spins < 200000) {
        cpu_relax();
      } else {
        std::unique_lock<std::mutex> lk(gr->cv_mu);
        gr->sleepers.fetch_add(1);
        gr->cv.wait_for(lk, std::chrono::milliseconds(50),
                        [&] { return gr->seq.load(std::memory_order_acquire) != seen || gr->stop.load(); });
        gr->sleepers.fetch_sub(1);
        spins = 0;
      }
      s = gr->seq.load(std::memory_order_acquire);
    }
    if (gr->stop.load()) return;
    seen = s;
    int rc;
    {
      std::lock_guard<std::mutex> lk(ix->mu);
      rc = serve_locked(gr, g, s);
      if (rc == VS_OK && g == 0 && gr->req.gather) rc = gather_locked(gr, s);
    }
    gr->rc[g] = rc;
    if (rc != VS_OK) {
      strncpy(gr->err[g], vs_last_error(), sizeof(gr->err[g]) - 1);
      gr->err[g][sizeof(gr->err[g]) - 1] = 0;
    }
    gr->acked[g].store(s, std::memory_order_release);
  }
}

// after a failed request: drain every GPU, re-align the exchange epochs, re-arm the error words
void resync(vs_group* gr) {
  uint32_t e = 0;
  for (int g = 0; g < gr->G; ++g) {
    vs_index* ix = gr->ix[g];
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard dg(ix->device);
    cudaStreamSynchronize(ix->stream);
    cudaGetLastError();
    if (ix->xc.epoch > e) e = ix->xc.epoch;
  }
  for (int g = 0; g < gr->G; ++g) {
    vs_index* ix = gr->ix[g];
    std::lock_guard<std::mutex> lk(ix->mu);
    ix->xc.epoch = e;
    if (ix->xc.h_err) *(volatile unsigned int*)ix->xc.h_err = 0;
    touch(ix);
  }
}

int run_request(vs_group* gr, const void* in, size_t in_bytes, int B, int k, const uint64_t* bits, int mode, int blend,
                float* out_scores, int64_t* out_rows) {
  // caller holds front_mu
  const auto t0 = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
  Request& r = gr->req;
  r.B = B;
  r.k = k;
  r.mode = mode & 0xff;
  r.blend = blend;
  r.gather = k > gr->k_max ? 1 : 0;
  r.use_bits = false;
  for (int w = 0; w < kMaskWords; ++w) {
    r.bits[w] = bits ? bits[w] : 0;
    if (r.bits[w]) r.use_bits = true;
  }
  memcpy(gr->h_in, in, in_bytes);
  unsigned s = gr->seq.load(std::memory_order_relaxed) + 1;
  if (s == 0) s = 1;
  gr->seq.store(s, std::memory_order_release);
  gr->t_submit_us = since();
  if (gr->sleepers.load() > 0) {
    std::lock_guard<std::mutex> lk(gr->cv_mu);
    gr->cv.notify_all();
  }
  // every worker has enqueued (or failed)
  int rc = VS_OK;
  for (int g = 0; g < gr->G; ++g) {
    while (gr->acked[g].load(std::memory_order_acquire) != s) cpu_relax();
    if (gr->rc[g] != VS_OK && rc == VS_OK) rc = fail(gr->rc[g], "shard %d: %s", g, gr->err[g]);
  }
  gr->t_enqueued_us = since();
  if (rc != VS_OK) {
    resync(gr);
    return rc;
  }
  // shard 0's result: flags in host-mapped memory, raised by the kernel itself on the fused path
  // (bounded: a kernel that died of an asynchronous CUDA error never raises its flag, and a server thread must not
  // spin forever on it.  The longest legitimate fused request -- 64 scans of a 150M-row shard -- takes ~1.5 s.)
  volatile unsigned int* done = gr->h_done;
  for (int b = 0; b < B; ++b) {
    for (unsigned spins = 1; done[b] != s; ++spins) {
      cpu_relax();
      if ((spins & 0xFFFFFu) == 0 && since() > kFlagTimeoutUs) {
        resync(gr);
        return fail(VS_ERR_CUDA, "no completion flag for query %d of the request within %.0f s (device fault?)", b,
                    kFlagTimeoutUs / 1e6);
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  gr->t_done_us = since();
  memcpy(out_scores, gr->h_out_s, (size_t)B * k * 4);
  memcpy(out_rows, gr->h_out_r, (size_t)B * k * 8);
  for (int g = 0; g < gr->G; ++g)
    if (vs_exchange_error(gr->ix[g])) {
      resync(gr);
      return fail(VS_ERR_EXCHANGE, "peer exchange timed out on shard %d of %d (a GPU never pushed its candidates within ~3 s)", g,
                  gr->G);
    }
  gr->t_return_us = since();
  return VS_OK;
}

}  // namespace

extern "C" {

int vs_group_create(int n_dev, const int* devices, int dim, int dtype, int64_t capacity_rows_total, int b_max, int k_max,
                    vs_group_t** out) {
  if (!out) return fail(VS_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (n_dev < 1 || n_dev > kMaxPeers) return fail(VS_ERR_ARG, "need 1 <= n_dev <= %d", kMaxPeers);
  if (b_max < 1 || b_max > 65536 || k_max < 1 || k_max > kMaxFusedK)
    return fail(VS_ERR_ARG, "need 1 <= b_max <= 65536 and 1 <= k_max <= %d", kMaxFusedK);
  vs_group* gr = new (std::nothrow) vs_group();
  if (!gr) return fail(VS_ERR_OOM, "host allocation failed");
  gr->G = n_dev;
  gr->dim = dim;
  gr->dtype = dtype;
  gr->b_max = b_max;
  gr->k_max = k_max;
  for (int g = 0; g < kMaxPeers; ++g) gr->acked[g].store(0);
  auto destroy_partial = [&](int rc) {
    for (int g = 0; g < n_dev; ++g)
      if (gr->ix[g]) vs_destroy(gr->ix[g]);
    if (gr->h_area) cudaFreeHost(gr->h_area);
    delete gr;
    return rc;
  };
  const int64_t cap_each = capacity_rows_total > 0 ? (capacity_rows_total + n_dev - 1) / n_dev : 0;
  for (int g = 0; g < n_dev; ++g) {
    int rc = vs_create(devices ? devices[g] : g, dim, dtype, cap_each, &gr->ix[g]);
    if (rc == VS_OK) rc = vs_set_row_map(gr->ix[g], g, n_dev);
    if (rc == VS_OK) rc = vs_exchange_create(gr->ix[g], n_dev, g, b_max, k_max);
    if (rc != VS_OK) return destroy_partial(rc);
  }
  void* ptrs[kMaxPeers] = {};
  for (int g = 0; g < n_dev; ++g) ptrs[g] = vs_exchange_local_ptr(gr->ix[g]);
  for (int g = 0; g < n_dev; ++g) {
    if (n_dev > 1) {
      int rc = vs_exchange_attach(gr->ix[g], nullptr, ptrs);
      if (rc != VS_OK) return destroy_partial(rc);
    }
  }
  // request/response area.  Response sized for the gather form as well (k up to kMaxK).
  const int k_resp = kMaxK;
  gr->in_bytes = ((size_t)2 * b_max * dim * 4 + (size_t)b_max * 8 + 255) & ~(size_t)255;
  gr->out_s_bytes = ((size_t)b_max * k_resp * 4 + 255) & ~(size_t)255;
  gr->out_r_bytes = ((size_t)b_max * k_resp * 8 + 255) & ~(size_t)255;
  const size_t done_bytes = ((size_t)b_max * 4 + 255) & ~(size_t)255;
  cudaError_t e = cudaHostAlloc((void**)&gr->h_area, gr->in_bytes + gr->out_s_bytes + gr->out_r_bytes + done_bytes,
                                cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return destroy_partial(fail(VS_ERR_OOM, "cudaHostAlloc(request area): %s", cudaGetErrorString(e)));
  }
  gr->h_in = gr->h_area;
  gr->h_out_s = (float*)(gr->h_area + gr->in_bytes);
  gr->h_out_r = (int64_t*)(gr->h_area + gr->in_bytes + gr->out_s_bytes);
  gr->h_done = (unsigned int*)(gr->h_area + gr->in_bytes + gr->out_s_bytes + gr->out_r_bytes);
  memset(gr->h_done, 0, done_bytes);
  // nothing may allocate on the request path (a cudaFree synchronises a GPU whose peers may be spinning)
  for (int g = 0; g < n_dev; ++g) {
    vs_index* ix = gr->ix[g];
    DeviceGuard dg(ix->device);
    const size_t qb = (size_t)b_max * dim * 4;
    cudaError_t e2 = ix->d_q.reserve(qb);
    if (e2 == cudaSuccess) e2 = ix->d_stage.reserve(3 * qb + (size_t)b_max * 8 + 64);
    if (e2 == cudaSuccess) e2 = ix->d_out_s.reserve((size_t)b_max * k_max * 12 + 128);
    if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&gr->ev[g], cudaEventDisableTiming);
    if (e2 != cudaSuccess) {
      cudaGetLastError();
      return destroy_partial(fail(VS_ERR_OOM, "group scratch on device %d: %s", ix->device, cudaGetErrorString(e2)));
    }
  }
  for (int g = 0; g < n_dev; ++g) gr->workers[g] = std::thread(worker_main, gr, g);
  *out = gr;
  return VS_OK;
}

int vs_group_destroy(vs_group_t* gr) {
  if (!gr) return VS_OK;
  {
    std::lock_guard<std::mutex> lk(gr->cv_mu);
    gr->stop.store(true);
    gr->cv.notify_all();
  }
  for (int g = 0; g < gr->G; ++g)
    if (gr->workers[g].joinable()) gr->workers[g].join();
  for (int g = 0; g < gr->G; ++g) {
    if (gr->ev[g]) {
      DeviceGuard dg(gr->ix[g]->device);
      cudaEventDestroy(gr->ev[g]);
    }
  }
  {
    DeviceGuard dg(gr->ix[0]->device);
    cudaStreamSynchronize(gr->ix[0]->stream);
    gr->g_cand.release();
  }
  for (int g = 0; g < gr->G; ++g) vs_destroy(gr->ix[g]);
  if (gr->h_area) cudaFreeHost(gr->h_area);
  cudaGetLastError();
  delete gr;
  return VS_OK;
}

int vs_group_size(const vs_group_t* gr) { return gr ? gr->G : 0; }
vs_index_t* vs_group_shard(vs_group_t* gr, int shard) { return (gr && shard >= 0 && shard < gr->G) ? gr->ix[shard] : nullptr; }
int vs_group_last_timing(const vs_group_t* gr, double out_us[4]) {
  if (!gr || !out_us) return fail(VS_ERR_ARG, "NULL argument");
  out_us[0] = gr->t_submit_us;
  out_us[1] = gr->t_enqueued_us;
  out_us[2] = gr->t_done_us;
  out_us[3] = gr->t_return_us;
  return VS_OK;
}
int64_t vs_group_count(const vs_group_t* gr) {
  int64_t n = 0;
  if (gr)
    for (int g = 0; g < gr->G; ++g) n += gr->ix[g]->n;
  return n;
}

int vs_group_query_host(vs_group_t* gr, const float* q, int B, int k, const uint64_t* require_bits, int mode,
                        float* out_scores, int64_t* out_rows) {
  if (!gr) return fail(VS_ERR_ARG, "group is NULL");
  if (B < 0 || (B > 0 && (!q || !out_scores || !out_rows))) return fail(VS_ERR_ARG, "NULL buffer");
  if (k <= 0 || k > kMaxK) return fail(VS_ERR_ARG, "k=%d out of range [1,%d]", k, kMaxK);
  if ((int64_t)gr->G * k > 16384 && k > gr->k_max) return fail(VS_ERR_UNSUPPORTED, "G*k exceeds the merge kernel's 16384 candidates");
  std::lock_guard<std::mutex> lk(gr->front_mu);
  for (int b0 = 0; b0 < B; b0 += gr->b_max) {
    const int nb = B - b0 < gr->b_max ? B - b0 : gr->b_max;
    int rc = run_request(gr, q + (size_t)b0 * gr->dim, (size_t)nb * gr->dim * 4, nb, k, require_bits, mode, 0,
                         out_scores + (size_t)b0 * k, out_rows + (size_t)b0 * k);
    if (rc) return rc;
  }
  return VS_OK;
}

int vs_group_query_multimodal_host(vs_group_t* gr, const float* img, const float* txt, const double* w, int B, int k,
                                   const uint64_t* require_bits, int mode, float* out_scores, int64_t* out_rows) {
  if (!gr) return fail(VS_ERR_ARG, "group is NULL");
  if (B < 0 || (B > 0 && (!img || !txt || !w || !out_scores || !out_rows))) return fail(VS_ERR_ARG, "NULL buffer");
  if (k <= 0 || k > kMaxK) return fail(VS_ERR_ARG, "k=%d out of range [1,%d]", k, kMaxK);
  if ((int64_t)gr->G * k > 16384 && k > gr->k_max) return fail(VS_ERR_UNSUPPORTED, "G*k exceeds the merge kernel's 16384 candidates");
  std::lock_guard<std::mutex> lk(gr->front_mu);
  // pack [img | txt | w] of each chunk contiguously, as the blend staging expects
  std::unique_ptr<char[]> pack(new (std::nothrow) char[gr->in_bytes]);
  if (!pack) return fail(VS_ERR_OOM, "host allocation failed");
  for (int b0 = 0; b0 < B; b0 += gr->b_max) {
    const int nb = B - b0 < gr->b_max ? B - b0 : gr->b_max;
    const size_t qb = (size_t)nb * gr->dim * 4;
    memcpy(pack.get(), img + (size_t)b0 * gr->dim, qb);
    memcpy(pack.get() + qb, txt + (size_t)b0 * gr->dim, qb);
    memcpy(pack.get() + 2 * qb, w + b0, (size_t)nb * 8);
    int rc = run_request(gr, pack.get(), 2 * qb + (size_t)nb * 8, nb, k, require_bits, mode, 1, out_scores + (size_t)b0 * k,
                         out_rows + (size_t)b0 * k);
    if (rc) return rc;
  }
  return VS_OK;
}

}  // extern "C"
