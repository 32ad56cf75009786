/* vecsearch_b200.h -- C ABI of the B200-native exact cosine vector-search engine.
 *
 * This is the drop-in boundary for the reference's vector-store seam.  The reference
 * (parsakhaz/multimodal-image-similarity-search) has no FFI layer of its own: the seam is the
 * duck-typed chromadb Collection returned by init_chromadb() (backend/app/utils.py:104-138)
 * and used as the module global `collection` (backend/app/main.py:77,530).  Each entry point
 * below names the Collection call (reference file:line) whose arithmetic it replaces; the
 * Python class that re-creates the Collection surface on top of these symbols lives in
 * multimodal-image-similarity-search_b200/collection.py and INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *  - plain C: pointers + sizes, int status return (0 = ok, <0 = error, text via
 *    vs_last_error()); no exceptions, no torch types, no hidden allocation on the query path
 *    beyond per-index scratch that is grown once and reused.
 *  - one vs_index_t == one row shard resident in the HBM of ONE device.  Multi-GPU = either ONE
 *    process holding a vs_group_t (one shard per GPU, worker thread per GPU, the reference's
 *    single-server shape) or one process per GPU, each owning a shard; either way the candidates
 *    are exchanged over NVLink peer memory from inside the query kernel (vs_exchange_*), or by
 *    the host layer (NCCL all-gather) and merged with vs_merge_topk_dev.
 *  - thread safety: every entry point locks the handle for its whole body, so concurrent callers
 *    on one handle serialise (staging buffers are per handle); different handles are independent.
 *    "_dev" entry points return when the work is enqueued; mutations (add / remove / set_row /
 *    filter bits) order themselves after queries still in flight on caller streams.
 *  - "_host" entry points take HOST buffers and include the H2D/D2H copies and the final
 *    stream synchronise; "_dev" entry points take DEVICE buffers and are asynchronous on the
 *    given cudaStream_t (passed as void*; NULL = the index's own stream).
 *  - scores are float32 cosine similarities (distance = 1 - score); rows are int64 shard-local
 *    row numbers plus the index's row_base; ranking is (score desc, row asc); empty result
 *    slots (k > count) carry score = -inf and row = -1.
 *  - there is NO CPU fallback: every compute entry point fails with VS_ERR_CUDA when no
 *    sm_100 device is usable.
 */
#ifndef VECSEARCH_B200_H
#define VECSEARCH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vs_index vs_index_t;

enum { VS_F32 = 0, VS_BF16 = 1 };

enum {
  VS_OK = 0,
  VS_ERR_ARG = -1,    /* bad argument */
  VS_ERR_CUDA = -2,   /* CUDA runtime / driver error, or no usable device */
  VS_ERR_OOM = -3,    /* device allocation failed */
  VS_ERR_UNSUPPORTED = -4,
  VS_ERR_OVERFLOW = -5, /* caller-provided output buffer too small (dedup pairs) */
  VS_ERR_EXCHANGE = -6  /* a peer shard never delivered its candidates: the affected results are EMPTY, not stale */
};

/* number of filter bits carried per row (4 x u64) -- the reference UI never has more than a
 * handful of filters (filters.json); 256 matches BASELINE config 4. */
#define VS_MASK_WORDS 4

/* Thread-local text of the last error raised on the calling thread. */
const char* vs_last_error(void);

/* ABI version of this header (bumped on any signature change). */
int vs_abi_version(void);

/* ---- lifecycle: replaces chromadb.PersistentClient(...).create_collection(name,
 *      metadata={"hnsw:space":"cosine"})  (backend/app/utils.py:113-130).  Only the cosine
 *      space exists.  `capacity_rows` is a reservation hint; the slab grows by doubling. */
int vs_create(int device, int dim, int dtype, int64_t capacity_rows, vs_index_t** out);
int vs_destroy(vs_index_t* ix);

/* Collection.count()  (init_db.py:58). */
int64_t vs_count(const vs_index_t* ix);
int vs_dim(const vs_index_t* ix);
int vs_dtype(const vs_index_t* ix);
int vs_device(const vs_index_t* ix);
/* Grow the slab to hold at least capacity_rows rows now (bulk loads: no re-allocation later). */
int vs_reserve(vs_index_t* ix, int64_t capacity_rows);

/* Offset added to every row number this shard reports (global row = row_base + local row). */
int vs_set_row_base(vs_index_t* ix, int64_t row_base);
/* General form: reported row = row_base + local row * row_stride.  Row-STRIPED shards (global row g
 * on shard g % G) use (shard, G), so every kernel reports and tie-breaks on true global rows. */
int vs_set_row_map(vs_index_t* ix, int64_t row_base, int64_t row_stride);

/* ---- ingest: Collection.add(ids, embeddings, ...)  (backend/app/main.py:735-740).
 *      Appends n float32 rows [n, dim] (row-major, contiguous); the kernel casts to the storage
 *      dtype and stores 1/(||row||+1e-30) of the STORED row beside it.  *first_row receives the
 *      local row number of the first appended row.  id<->row bookkeeping is the host layer's. */
int vs_add_host(vs_index_t* ix, const float* rows, int64_t n, int64_t* first_row);
int vs_add_dev(vs_index_t* ix, const float* rows_dev, int64_t n, int64_t* first_row, void* stream);

/* Persistence slab (the store the reference reopens at backend/app/utils.py:109-123 and
 * walks at backend/app/main.py:522-579): rows in the STORAGE dtype (bf16 / f32), dim elements each,
 * no pitch padding.  vs_get_raw_host copies stored rows out bit for bit; vs_add_raw_host appends such
 * rows through double-buffered pinned staging and recomputes the inverse norms on the device, so a
 * reloaded collection scores bit-identically to the saved one. */
int vs_add_raw_host(vs_index_t* ix, const void* stored_rows, int64_t n, int64_t* first_row);
int vs_get_raw_host(const vs_index_t* ix, int64_t first_row, int64_t n, void* out);

/* ---- Collection.delete(ids)  (backend/app/main.py:1069).  Removes local row `row` by moving
 *      the LAST row into its place (dense slab, no tombstones).  *moved_from = the row that was
 *      moved (== old count-1), or -1 if `row` was the last row. */
int vs_remove(vs_index_t* ix, int64_t row, int64_t* moved_from);
/* Bulk form (reset_system deletes every id in one call, backend/app/main.py:1065-1069): removes the m
 * distinct rows `rows[]` with ONE compaction kernel and one synchronise.  The new count is count - m;
 * every hole below it is filled by a surviving row from at or above it.  The moves are reported as
 * moved_src[i] -> moved_dst[i], i < *n_moved <= m (both arrays must hold m entries) so the host layer
 * can re-point its id<->row map. */
int vs_remove_rows(vs_index_t* ix, const int64_t* rows, int64_t m, int64_t* moved_src, int64_t* moved_dst,
                   int64_t* n_moved);
/* Drop the rows at and above new_count (keeps the allocation). */
int vs_truncate(vs_index_t* ix, int64_t new_count);
/* dst row <- src row (vector, inverse norm, filter bits), possibly across two GPUs of this process
 * (a group's delete moves the LAST global row into the hole, which may live on another shard). */
int vs_copy_row(vs_index_t* dst, int64_t dst_row, vs_index_t* src, int64_t src_row);
/* Batched form: dst row dst_rows[i] <- src row src_rows[i], i < n, in ONE kernel on dst's GPU (the source
 * shard is read over NVLink peer memory).  Sources and destinations must be disjoint rows when dst == src.
 * A group's bulk delete issues at most G*G of these and then vs_truncate()s every shard. */
int vs_move_rows(vs_index_t* dst, vs_index_t* src, const int64_t* src_rows, const int64_t* dst_rows, int64_t n);
/* dst row (dst_first + l * dst_stride) <- src row l for every row of src (vectors + inverse norms), read
 * over NVLink peer memory by a kernel on dst's GPU; dst grows as needed.  Replicates the row-striped
 * shards of a group into one full index per GPU for the all-pairs pass. */
int vs_replicate_from(vs_index_t* dst, vs_index_t* src, int64_t dst_first, int64_t dst_stride);

/* Overwrite the vector of an existing row in place (the row-striped sharded collection moves the
 * LAST global row into a deleted row's slot, which may live on another shard). */
int vs_set_row_host(vs_index_t* ix, int64_t row, const float* vec);

/* Drop all rows, keep the allocation  (reset path, backend/app/main.py:1058-1098). */
int vs_clear(vs_index_t* ix);

/* ---- per-row filter bits: bit f set <=> stored answer for filter f is "yes"
 *      (the predicate of backend/app/main.py:215, written by :1010-1033). */
int vs_set_mask_bits(vs_index_t* ix, int64_t row, const uint64_t bits[VS_MASK_WORDS]);
int vs_get_mask_bits(const vs_index_t* ix, int64_t row, uint64_t bits[VS_MASK_WORDS]);
/* Bulk form for ingest: bits[n][VS_MASK_WORDS] for rows [first_row, first_row + n) in ONE copy. */
int vs_set_mask_bits_range(vs_index_t* ix, int64_t first_row, int64_t n, const uint64_t* bits);
int vs_get_mask_bits_range(const vs_index_t* ix, int64_t first_row, int64_t n, uint64_t* bits);
/* Filter sweep -> stored bits without leaving the GPU: filter bit `bit` of every row r := bit r of
 * words_dev (one filter's row of vs_filter_sweep_dev output).  The CLIP-side analogue of the answers
 * the Moondream loop writes at backend/app/main.py:1010-1033. */
int vs_apply_sweep_bits_dev(vs_index_t* ix, const uint32_t* words_dev, int bit, void* stream);

/* Copy stored rows back as float32 (Collection.get(include=["embeddings"]) and persistence). */
int vs_get_rows_host(const vs_index_t* ix, int64_t first_row, int64_t n, float* out);
/* Same into a DEVICE buffer [n, dim] float32 (replicating shards for the all-pairs pass). */
int vs_get_rows_dev(const vs_index_t* ix, int64_t first_row, int64_t n, float* out_dev, void* stream);

/* ---- query: Collection.query(query_embeddings=[...], n_results=k, ...)
 *      (backend/app/main.py:761-765; legacy app.py:310-314).
 *      q: [B, dim] float32, NOT assumed unit-norm (normalised on device).
 *      require_bits: NULL, or VS_MASK_WORDS words applied to every query: only rows whose
 *      filter bits contain all required bits compete ("pre" filter mode, fused in the scan).
 *      mode: VS_Q_AUTO picks the HBM-bound scan (B small, or f32 storage) or the tcgen05
 *      batched kernel (bf16 storage, B >= 16); the other values force one path.
 *      out_scores/out_rows: [B, k].  k <= 1024 (main.py:757 caps "All" at 1000).
 *      VS_Q_PIPELINED may be OR-ed into `mode` of a "_dev" call: the caller vouches that the query
 *      buffer was complete BEFORE the previous vecsearch launch on that stream (e.g. a batch of
 *      queries prepared up front and issued one call each).  The scan kernel then starts streaming
 *      while the previous query still merges (programmatic dependent launch with the
 *      griddepcontrol.wait deferred past the scan).  Without it -- the default -- the kernel waits
 *      for everything before it in the stream before reading q or any row. */
enum { VS_Q_AUTO = 0, VS_Q_SCAN = 1, VS_Q_TENSOR = 2, VS_Q_PIPELINED = 0x100 };
int vs_query_topk_host(vs_index_t* ix, const float* q, int B, int k, const uint64_t* require_bits,
                       int mode, float* out_scores, int64_t* out_rows);
int vs_query_topk_dev(vs_index_t* ix, const float* q_dev, int B, int k, const uint64_t* require_bits,
                      int mode, float* out_scores_dev, int64_t* out_rows_dev, void* stream);

/* ---- multimodal blend: search_multimodal  (backend/app/main.py:850-860):
 *      c = w*img/||img|| + (1-w)*txt/||txt||, c /= ||c||, for B (img, txt, w) triples -> [B, dim]
 *      float32 on the device; feed the result to vs_query_topk_dev.  Weights are float64 (the
 *      reference's python float): w and 1-w are each rounded to float32 once, as numpy does. */
int vs_blend_dev(vs_index_t* ix, const float* img_dev, const float* txt_dev, const double* w_dev,
                 int B, float* out_dev, void* stream);
/* search_multimodal end to end with HOST buffers: H2D of both embeddings and the weights, blend
 * kernel, query, D2H (the body of backend/app/main.py:829-867 after the two CLIP calls). */
int vs_query_multimodal_host(vs_index_t* ix, const float* img, const float* txt, const double* w, int B,
                             int k, const uint64_t* require_bits, int mode, float* out_scores,
                             int64_t* out_rows);

/* ---- k-way merge of per-shard candidates after the all-gather:
 *      cand_scores/cand_rows: [G, B, k] device buffers (rows are global, <0 = empty) ->
 *      out [B, k].  `device` selects the GPU when ix is NULL. */
int vs_merge_topk_dev(vs_index_t* ix, const float* cand_scores_dev, const int64_t* cand_rows_dev,
                      int G, int B, int k, float* out_scores_dev, int64_t* out_rows_dev, void* stream);

/* ---- sharded query with the exchange fused in (replaces local query +
 *      ncclAllGather + vs_merge_topk_dev).  Every rank (one process per GPU, or several handles in
 *      one process) owns an exchange buffer that all G ranks have mapped over NVLink peer memory:
 *        1. vs_exchange_create(ix, G, rank, B_max, k_max)    allocate + zero the local buffer
 *        2. vs_exchange_ipc_handle(ix, h)                    64-byte cudaIpcMemHandle_t to publish
 *           (or vs_exchange_local_ptr for handles living in the same process)
 *        3. vs_exchange_attach(ix, handles, peer_ptrs)       map the peers' buffers: for peer g the
 *           pointer peer_ptrs[g] is used when given, else handles[64*g..] is opened with
 *           cudaIpcOpenMemHandle (entry `rank` is ignored).  Call after ALL ranks finished step 1.
 *      vs_query_topk_sharded_dev then behaves like vs_query_topk_dev but returns the GLOBAL top-k
 *      on every rank: for B <= 64 on the scan path the LAST CTA of the fused scan kernel stores
 *      the shard's k candidates straight into every peer's buffer (P2P stores), raises per-query
 *      flags (st.release.sys), spins (bounded) on the G local flags and merges -- one kernel per
 *      query, no NCCL call, no host synchronisation.  Larger batches / the tcgen05 path run the
 *      local query and one exchange kernel with the same wire protocol.  All ranks must issue the
 *      same sequence of sharded queries (same B, k); k <= k_max <= 128; global rows < 2^32.
 *      vs_exchange_merge_dev exposes the exchange kernel alone for caller-made [B,k] candidates.
 *      A wait that times out (~3 s: a peer never pushed) NEVER merges stale lists: the affected
 *      results are written EMPTY (-inf, -1) and a sticky error word in host-mapped memory is raised.
 *      vs_exchange_error() reads it (no CUDA call); the "_host" entry points return VS_ERR_EXCHANGE and
 *      re-arm; "_dev" entry points refuse further exchanges until vs_exchange_clear_error(). */
size_t vs_exchange_bytes(int B_max, int k_max, int G);
int vs_exchange_create(vs_index_t* ix, int G, int rank, int B_max, int k_max);
int vs_exchange_ipc_handle(vs_index_t* ix, unsigned char handle_out[64]);
void* vs_exchange_local_ptr(vs_index_t* ix);
int vs_exchange_attach(vs_index_t* ix, const unsigned char* ipc_handles, void* const* peer_ptrs);
int vs_query_topk_sharded_dev(vs_index_t* ix, const float* q_dev, int B, int k, const uint64_t* require_bits,
                              int mode, float* out_scores_dev, int64_t* out_rows_dev, void* stream);
int vs_exchange_merge_dev(vs_index_t* ix, const float* cand_scores_dev, const int64_t* cand_rows_dev, int B,
                          int k, float* out_scores_dev, int64_t* out_rows_dev, void* stream);
int vs_exchange_error(vs_index_t* ix);
int vs_exchange_clear_error(vs_index_t* ix);
/* HOST-buffer flavour of vs_query_topk_sharded_dev (one call per request on every rank: pinned H2D of
 * the replicated queries, query kernel with the fused exchange, D2H of the global result, sync). */
int vs_query_topk_sharded_host(vs_index_t* ix, const float* q, int B, int k, const uint64_t* require_bits,
                               int mode, float* out_scores, int64_t* out_rows);
/* Deferred form for a stream of independent queries (throughput mode): vs_exchange_begin opens a
 * new exchange; each vs_query_topk_push_dev runs the local query and pushes its candidates into
 * slots [slot0, slot0+B) of every peer (fused into the scan kernel as above, no waiting); ONE
 * vs_exchange_collect_dev(B_total, k) then waits for and merges slots [0, B_total) -> [B_total, k].
 * The ranks meet once per batch instead of once per query.  Same k for every push of an exchange. */
int vs_exchange_begin(vs_index_t* ix);
int vs_query_topk_push_dev(vs_index_t* ix, const float* q_dev, int B, int k, const uint64_t* require_bits,
                           int mode, int slot0, void* stream);
int vs_exchange_collect_dev(vs_index_t* ix, int B, int k, float* out_scores_dev, int64_t* out_rows_dev,
                            void* stream);

/* ---- vs_group_t: the reference's deployment shape -- ONE server process (backend/run.py:10-14) -- on all
 *      GPUs of the box.  n_dev row shards (global row g lives on shard g % n_dev at local row g / n_dev;
 *      every kernel reports global rows), their exchange buffers wired by pointer, and one resident worker
 *      thread per GPU.  vs_group_query_host is the request/response path of search_similar
 *      (backend/app/main.py:748-805): the query is written into a pinned host-mapped area, every worker
 *      issues one small H2D + ONE fused scan launch on its GPU, the kernels exchange candidates over NVLink
 *      and shard 0's kernel writes the global [B, k] result and a completion flag straight into host-mapped
 *      memory, which the calling thread polls -- no stream synchronise, no result copy, no torchrun.
 *      B > 64 / the tcgen05 path / k > k_max (gathered onto GPU 0 over NVLink and merged there) use the
 *      same entry point.  Ingest and maintenance go through the shards: vs_group_shard(g, s) returns the
 *      vs_index_t of shard s (row map (s, n_dev) already set).  devices == NULL means 0..n_dev-1.
 *      Errors: a peer that never pushes makes the kernels give up after ~3 s (VS_ERR_EXCHANGE, results empty); a
 *      completion flag that never arrives (device fault) ends the poll after 30 s with VS_ERR_CUDA.  Both re-align the
 *      shards' exchange epochs so that the next request starts clean. */
typedef struct vs_group vs_group_t;
int vs_group_create(int n_dev, const int* devices, int dim, int dtype, int64_t capacity_rows_total, int b_max,
                    int k_max, vs_group_t** out);
int vs_group_destroy(vs_group_t* g);
int vs_group_size(const vs_group_t* g);
vs_index_t* vs_group_shard(vs_group_t* g, int shard);
int64_t vs_group_count(const vs_group_t* g);
/* Host-side timeline of the LAST vs_group_query* request, microseconds since its entry: [0] query published to the
 * workers, [1] every worker has enqueued its launch, [2] completion flags seen, [3] result copied out (introspection). */
int vs_group_last_timing(const vs_group_t* g, double out_us[4]);
int vs_group_query_host(vs_group_t* g, const float* q, int B, int k, const uint64_t* require_bits, int mode,
                        float* out_scores, int64_t* out_rows);
/* search_multimodal (backend/app/main.py:829-867) on a group: every GPU blends its own copy of the
 * (img, txt, w) triples (vs_blend_dev's kernel), then the same query path. */
int vs_group_query_multimodal_host(vs_group_t* g, const float* img, const float* txt, const double* w, int B,
                                   int k, const uint64_t* require_bits, int mode, float* out_scores,
                                   int64_t* out_rows);

/* ---- filter sweep (BASELINE config 4; CLIP-side analogue of process_filter_on_all_images,
 *      backend/app/main.py:939-1056): prompts [F, dim] float32 (host or device) -> bit mask
 *      out_bits[F][words_per_filter] (uint32, bit n%32 of word n/32 set <=> cos >= tau),
 *      words_per_filter = vs_filter_words() = 8 * ceil(count/256).  tcgen05, bf16 storage. */
int64_t vs_filter_words(const vs_index_t* ix);
int vs_filter_sweep_dev(vs_index_t* ix, const float* prompts_dev, int F, float tau,
                        uint32_t* out_bits_dev, void* stream);
int vs_filter_sweep_host(vs_index_t* ix, const float* prompts, int F, float tau, uint32_t* out_bits);

/* ---- all-pairs duplicate detection (BASELINE config 5; embedding analogue of the pHash id
 *      lookup, backend/app/main.py:627-640): pairs (i<j) with cos >= tau among the rows
 *      [row_lo, row_hi) x [0, count) of this index (row ranges let G ranks split the triangle).
 *      Pairs are appended unordered to out_i/out_j/out_score (device, capacity `cap`);
 *      *out_count_dev receives the number found (may exceed cap -> VS_ERR_OVERFLOW from the
 *      host variant; the device variant just stops writing). */
int vs_dedup_dev(vs_index_t* ix, int64_t row_lo, int64_t row_hi, float tau, int64_t cap,
                 int64_t* out_i_dev, int64_t* out_j_dev, float* out_score_dev,
                 unsigned long long* out_count_dev, void* stream);
int vs_dedup_host(vs_index_t* ix, int64_t row_lo, int64_t row_hi, float tau, int64_t cap,
                  int64_t* out_i, int64_t* out_j, float* out_score, int64_t* out_count);

/* ---- introspection for bench.py / tests: number of kernels this library has launched on the
 *      calling process since load (the "gpu_launches" claim), and the last query's path. */
uint64_t vs_launch_count(void);
int vs_last_query_path(const vs_index_t* ix);   /* VS_Q_SCAN or VS_Q_TENSOR */
int vs_device_sm_count(const vs_index_t* ix);

#ifdef __cplusplus
}
#endif
#endif /* VECSEARCH_B200_H */
