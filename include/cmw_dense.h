/* cmw_dense.h -- C ABI of libcmwdense.so: the B200-native dense-retrieval hot path of cmw-rag.
 *
 * The reference (arterm-sedov/cmw-rag) is pure Python and has NO native boundary for this
 * path: scoring + top-k run inside an external ChromaDB server reached over HTTP.  Each entry
 * point below names the reference interface it stands in for (paths relative to the reference
 * checkout).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C types only; every *_dev pointer is a CUDA device pointer on the store's device,
 *     every *_host pointer is ordinary host memory; `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream).  Device-pointer entry points are stream-ordered and do not
 *     synchronise; outputs are valid when the stream reaches that point.
 *   - return value: 0 = ok, negative = error; the message is in cmw_last_error() (thread-local).
 *   - row ids are int64 = row number in append order + the store's id offset (row shards).
 *   - unused output slots (k larger than the number of live rows): id = -1, score = -inf.
 *   - threading: the *_host entry points of one store are serialised internally (they share its pinned
 *     staging buffers and stream); device-pointer entry points may run concurrently on different
 *     streams when each call has its own workspace; append / tombstone are exclusive with searches
 *     of the same store.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef CMW_DENSE_H_
#define CMW_DENSE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMW_ABI_VERSION 6

/* metric -- rag_engine/storage/vector_store.py:48-51 fixes the collection to cosine
 * ({"hnsw:space": "cosine"}); inner product is the north star's second metric. */
#define CMW_METRIC_COSINE 0
#define CMW_METRIC_IP 1

/* search mode */
/* F32_EXACT: fp64-accumulated scores from the fp32 rows; a query whose out_flags entry is 0 carries a PROOF that
 * its ids are those of the exact fp64 oracle (ties -> lower id): every row that was not rescored has a filter
 * score which, plus a rigorous bound on the filter's error, stays below the k-th exact score.  The bound behind
 * the bf16 tensor-core filter is r_q + (1 + r_q) R_c + D 2^-23 (rounding residual of the query tile, largest
 * rounding residual of any stored row, fp32 accumulation), both residuals measured from the values actually
 * written; behind the fp32 scan filter it is (D/32 + 12) 2^-24.  A query the proof fails for is flagged
 * CMW_FLAG_UNCERTIFIED (its ids are still the best of the candidates that were rescored); cmw_search_host
 * re-runs flagged queries through wider / more precise filters by itself.  cmw_set_option("strict_certificate", 0)
 * trades the proof for a statistical bound (independent roundings, 8 sigma) and ~4 % more throughput. */
#define CMW_MODE_F32_EXACT 0
#define CMW_MODE_BF16 1      /* 16-bit tile operands (bf16 or fp16, as the store holds), fp32 accumulation; approximate (reported as recall@k) */
/* optional algorithm override, OR-ed into `mode` (default: chosen from the batch size) */
#define CMW_ALGO_AUTO (0 << 8)
#define CMW_ALGO_SCAN (1 << 8) /* K1: TMA-staged streaming dot product, 1-4 queries per pass (HBM-bound);
                                  F32_EXACT scans the fp32 tiles, BF16 the bf16 tiles */
#define CMW_ALGO_GEMM (2 << 8) /* K2: tcgen05/TMEM GEMM with fused top-k epilogue */
#define CMW_ALGO_GEMM_TF32 (3 << 8) /* K2 over the fp32 tiles themselves, read by kind::tf32 MMAs (exact mode only):
                                  what a store WITHOUT 16-bit tiles runs by default -- one pass over the fp32 rows for
                                  any batch size, where K1 needs one pass per 4 queries */
/* slab schedule override, OR-ed into `mode`: fixed slabs small enough that the candidate pool can
 * never overflow, whatever the row order (slower; cmw_search_host falls back to it by itself) */
#define CMW_SLABS_SAFE (1 << 16)
/* keep the largest candidate set (K' = 1024) per query: the most head-room for the certificate */
#define CMW_KPRIME_MAX (1 << 17)

/* store flags */
#define CMW_STORE_F32 1u  /* keep row-major fp32 tiles (needed by CMW_MODE_F32_EXACT) */
#define CMW_STORE_BF16 2u /* keep row-major bf16 tiles of the L2-normalised rows */
#define CMW_STORE_F16 4u  /* the 16-bit tiles hold fp16 instead of bf16 (alternative to CMW_STORE_BF16): the same
                             bytes and tensor-core rate, 11 significand bits instead of 8 -- the rounding residual
                             the exactness certificate must absorb is 8x smaller (eps ~ 6.5e-4 instead of 3.5e-3 at
                             D = 1536), so ~2.5x fewer rows are rescored.  L2-normalised rows never leave fp16's
                             range; elements that would round below 2^-14 are flushed to zero and counted. */

/* per-query flags written by cmw_search (out_flags) */
#define CMW_FLAG_UNCERTIFIED 1 /* the exactness certificate could not be established, or a candidate pool overflowed */
#define CMW_FLAG_PEER_TIMEOUT 2 /* cmw_exchange_merge / cmw_peer_gather: a peer never published its data (results invalid) */

/* limits (checked, not silent): k <= CMW_MAX_K; cmw_multivector S*k <= CMW_MAX_MULTIVECTOR_ENTRIES;
 * cmw_merge_topk / cmw_shard_merge / cmw_exchange_merge G*k <= CMW_MAX_MERGE_ENTRIES.  k above ~330 cannot be
 * certified behind the bf16 filter (its candidate set K' = max(k + 108, 3k + 20) is capped at 1024): those
 * queries come back flagged from cmw_search and are repaired through the fp32 scan by cmw_search_host. */
#define CMW_MAX_K 1024
#define CMW_MAX_MULTIVECTOR_ENTRIES 2048
#define CMW_MAX_MERGE_ENTRIES 8192

typedef struct cmw_store cmw_store;

typedef struct cmw_store_info {
    int32_t device;
    int32_t dim;
    uint32_t flags;
    int32_t sm_count;
    int64_t capacity_rows;
    int64_t rows;      /* appended so far (including tombstoned) */
    int64_t live_rows; /* rows - tombstoned */
    int64_t id_offset;
    int64_t hbm_bytes; /* device memory held by the store */
} cmw_store_info;

const char* cmw_last_error(void);
int cmw_abi_version(void);

/* ---- corpus store: replaces the Chroma collection of
 *      rag_engine/storage/vector_store.py:44-52 (get_or_create_collection) ---- */
int cmw_store_create(int device, int dim, int64_t capacity_rows, uint32_t flags, int64_t id_offset,
                     cmw_store** out);
int cmw_store_destroy(cmw_store* s);
int cmw_store_get_info(const cmw_store* s, cmw_store_info* out);

/* K0 ingest -- replaces collection.add(embeddings=...) of vector_store.py:68-82.
 * rows: fp32 [n, dim] row-major.  kb_gid: int32 [n] dense group number of the row's normalised
 * kbId (rag_engine/utils/metadata_utils.py:20-32), negative = falsy kbId; may be NULL (= -1).
 * Builds fp32 tiles, bf16 tiles of the normalised rows, norm / inv_norm. */
int cmw_store_append_f32(cmw_store* s, const float* rows_dev, const int32_t* kb_gid_dev, int64_t n,
                         void* stream);
int cmw_store_append_host_f32(cmw_store* s, const float* rows_host, const int32_t* kb_gid_host,
                              int64_t n);
/* Device-to-device re-ingest (store maintenance without a host round trip: growing a full collection, dropping
 * tombstoned rows physically): appends n rows of `src` to `dst` -- the contiguous range [src_row0, src_row0 + n), or
 * the LOCAL row numbers src_rows_dev i64[n] (valid rows of src; the caller picks the live ones) -- with their
 * kb_gid; tiles and norms are rebuilt by K0 in dst's own formats.  Same device, same dim; src needs fp32 tiles. */
int cmw_store_copy_rows(cmw_store* dst, const cmw_store* src, const int64_t* src_rows_dev, int64_t src_row0,
                        int64_t n, void* stream);
/* replaces collection.delete(where=...) of vector_store.py:102-105 once the host has resolved the
 * filter to row numbers (LOCAL rows, i.e. without id_offset).  Tombstoned rows are never returned. */
int cmw_store_tombstone(cmw_store* s, const int64_t* rows_dev, int64_t n, void* stream);
int cmw_store_tombstone_host(cmw_store* s, const int64_t* rows_host, int64_t n);
/* device pointer to kb_gid[rows] (for cmw_multivector) */
const int32_t* cmw_store_kb_gid_dev(const cmw_store* s);
/* Read rows back to the host (persistence: the analogue of the Chroma server's --path directory,
 * systemd/cmw-rag-chroma.service:11 of the reference).  rows_host f32 [n, dim] (needs CMW_STORE_F32),
 * kb_gid_host i32 [n], live_host u8 [n] (0 = tombstoned); any may be NULL.  Synchronises. */
int cmw_store_read_rows_f32(cmw_store* s, int64_t row0, int64_t n, float* rows_host, int32_t* kb_gid_host,
                            uint8_t* live_host);

/* ---- search: replaces collection.query(query_embeddings, n_results) of
 *      vector_store.py:54-66 for a whole batch of query vectors ---- */
size_t cmw_search_workspace_bytes(const cmw_store* s, int batch, int k, int mode);
/* queries: fp32 [batch, dim].  out_scores f32 [batch,k], out_ids i64 [batch,k];
 * out_scores64 f64 [batch,k] (may be NULL; the scores the ranking used, needed for an exact
 * cross-shard merge); out_flags i32 [batch] (may be NULL).  ws: >= cmw_search_workspace_bytes. */
int cmw_search(cmw_store* s, const float* queries_dev, int batch, int k, int metric, int mode,
               float* out_scores_dev, int64_t* out_ids_dev, double* out_scores64_dev,
               int32_t* out_flags_dev, void* ws_dev, size_t ws_bytes, void* stream);
/* End-to-end form with HOST buffers (what a ctypes / HTTP-replacing caller uses): pinned staging,
 * H2D of the queries, search, D2H of the results, synchronises before returning. */
int cmw_search_host(cmw_store* s, const float* queries_host, int batch, int k, int metric, int mode,
                    float* out_scores_host, int64_t* out_ids_host, int32_t* out_flags_host);

/* Pipelined form of cmw_search_host for callers that keep several requests in flight -- the reference does
 * (S concurrent awaits per request: rag_engine/retrieval/retriever.py:179-182; concurrent requests:
 * rag_engine/config/settings.py:166).  submit enqueues H2D (copy-in stream) -> search (the store's compute
 * stream, in order over all tickets) -> D2H (copy-out stream) and returns a ticket without waiting, so the
 * copies of one ticket overlap the kernels of its neighbours; wait blocks until that ticket's results are in
 * the caller's buffers (and runs the repair chain for flagged queries).  At most CMW_HOST_SLOTS tickets per
 * store are in flight: a further submit fails with -4 until one is waited for.  The query and output buffers
 * must stay valid and untouched from submit to wait.  Page-locked buffers are used for DMA directly, pageable
 * ones are staged.  Thread-safe; tickets may be waited for in any order, each exactly once. */
#define CMW_HOST_SLOTS 4
int cmw_search_host_submit(cmw_store* s, const float* queries_host, int batch, int k, int metric, int mode,
                           float* out_scores_host, int64_t* out_ids_host, int32_t* out_flags_host,
                           int* ticket_out);
int cmw_search_host_wait(cmw_store* s, int ticket);

/* ---- multi-vector reduction: replaces the Python loops of
 *      rag_engine/retrieval/retriever.py:185-194 (ordered union, first-seen dedup),
 *      :208-210 (pre-rerank cap), :229-242 (truncate + group by normalised kbId, max score),
 *      :307 (stable sort by score descending) for Q long queries at once.
 * ids i64 [Q,S,k] / scores f32 [Q,S,k]: per-segment top-k (LOCAL or global ids; negative = pad).
 * kb_gid: int32 table indexed by (id - id_offset).  P = (0 < prl < S*k) ? prl : S*k.
 * limit: group only the first `limit` candidates (0 = all)  [retriever.py:229-231].
 * Outputs (any may be NULL): cand_ids i64[Q,P], cand_scores f32[Q,P] (first-seen occurrence),
 * cand_best f32[Q,P] (max over occurrences), cand_n i32[Q], cand_grp i32[Q,P] (group index in
 * first-appearance order or -1), grp_gid i32[Q,P], grp_max f32[Q,P], grp_cnt i32[Q,P],
 * grp_first i32[Q,P], grp_order i32[Q,P] (group indices, stable score-desc), grp_n i32[Q]. */
int cmw_multivector(const int32_t* kb_gid_dev, int64_t kb_rows, int64_t id_offset,
                    const int64_t* ids_dev, const float* scores_dev, int Q, int S, int k, int prl,
                    int limit, int64_t* cand_ids, float* cand_scores, float* cand_best,
                    int32_t* cand_n, int32_t* cand_grp, int32_t* grp_gid, float* grp_max,
                    int32_t* grp_cnt, int32_t* grp_first, int32_t* grp_order, int32_t* grp_n,
                    void* stream);

/* ---- cross-shard merge (after the NCCL all-gather of per-GPU candidates):
 * scores f64 [G,B,k_in], ids i64 [G,B,k_in] -> top k_out by (score desc, id asc). */
int cmw_merge_topk(const double* scores_dev, const int64_t* ids_dev, int G, int B, int k_in,
                   int k_out, float* out_scores_dev, int64_t* out_ids_dev,
                   double* out_scores64_dev, void* stream);

/* ---- row-sharded search: cmw_search cut in two around ONE small exchange, so that the fp64
 * rescoring -- the part of a search that does not shrink with the shard -- is shared between the shards instead of
 * repeated on each.  Every rank holds a row shard (store with id_offset = its first row) and the same queries:
 *   1. cmw_search_filter   prep + filter + compaction; out_filter_topk f32 [batch,k] = the best k FILTER scores
 *   2. all-gather 1        [G,batch,k] f32                                  (NCCL, or any transport)
 *   3. cmw_shard_kth       k-th best filter score over all shards, per query -> global_kth f32 [batch]
 *   4. cmw_search_finish   rescoring of the local candidates that can still reach the GLOBAL top-k (filter score
 *                          >= global_kth - 2 eps), selection; writes one packed block of cmw_shard_block_bytes:
 *                          f64 scores [batch,k] | i64 ids [batch,k] | f64 {t, eps} [batch,2] | i32 flags [batch]
 *   5. all-gather 2        G blocks
 *   6. cmw_shard_merge     G*k -> k_out by (score desc, id asc) + the cross-shard certificate
 *                          (k-th merged exact score > max over shards of t + eps)
 * global_kth_dev may be NULL (each shard then uses its own k-th: steps 2-3 skipped; CMW_MODE_BF16 always does).
 * filter and finish take the SAME queries / batch / k / metric / mode / workspace; nothing else may use that
 * workspace in between.  The reference has no counterpart (one Chroma server: vector_store.py:34-42). */
int cmw_search_filter(cmw_store* s, const float* queries_dev, int batch, int k, int metric, int mode,
                      float* out_filter_topk_dev, void* ws_dev, size_t ws_bytes, void* stream);
int cmw_shard_kth(const float* filter_topk_gathered_dev, int G, int B, int k, float* out_kth_dev, void* stream);
size_t cmw_shard_block_bytes(int batch, int k);
int cmw_search_finish(cmw_store* s, const float* queries_dev, int batch, int k, int metric, int mode,
                      const float* global_kth_dev, void* block_dev, void* ws_dev, size_t ws_bytes, void* stream);
int cmw_shard_merge(const void* blocks_dev, int G, int B, int k, int k_out, float* out_scores_dev,
                    int64_t* out_ids_dev, double* out_scores64_dev, int32_t* out_flags_dev, void* stream);
/* the same with the status word of cmw_peer_gather: if *peer_status_dev != 0 when the kernel runs, every query
 * comes back empty (ids -1) with that value as its flags (CMW_FLAG_PEER_TIMEOUT) */
int cmw_shard_merge_ex(const void* blocks_dev, int G, int B, int k, int k_out, float* out_scores_dev,
                       int64_t* out_ids_dev, double* out_scores64_dev, int32_t* out_flags_dev,
                       const int32_t* peer_status_dev, void* stream);

/* ---- fused exchange + merge over NVLink peer memory (alternative to all-gather + cmw_merge_topk).
 * Every rank allocates one peer buffer (cmw_peer_alloc: cudaMalloc + cudaIpc handle, zero-initialised,
 * cmw_peer_buffer_bytes bytes), exchanges the 64-byte handles out of band (torch.distributed, MPI, a
 * file...), opens the others (cmw_peer_open) and passes the G device pointers (its own at index `rank`)
 * to cmw_exchange_merge.  That call launches, stream-ordered and without host synchronisation: a send
 * kernel that stores this rank's [B,k] (f64 score, i64 id) candidates into every peer's buffer over
 * NVLink and publishes an epoch flag, and a merge kernel that waits for the G flags and reduces
 * G*k -> k_out per query by (score desc, id asc).  `epoch` must be non-zero and increase by 1 per call
 * on every rank (buffers are double-buffered by its parity); B <= max_batch, k <= max_k as allocated. */
size_t cmw_peer_buffer_bytes(int G, int max_batch, int max_k);
int cmw_peer_alloc(int device, size_t bytes, void** dev_ptr, void* ipc_handle_out /* 64 bytes */);
int cmw_peer_open(int device, const void* ipc_handle /* 64 bytes */, void** dev_ptr);
int cmw_peer_close(void* dev_ptr);
int cmw_peer_free(void* dev_ptr);
int cmw_exchange_merge(void* const* peer_bufs_host, int G, int rank, int max_batch, int max_k, int B, int k,
                       int k_out, uint32_t epoch, const double* scores64_local_dev, const int64_t* ids_local_dev,
                       float* out_scores_dev, int64_t* out_ids_dev, double* out_scores64_dev, void* stream);
/* Same, with per-query status.  flags_local_dev i32 [B] (may be NULL): this shard's own search flags, which travel
 * with the candidates; out_flags_dev i32 [B] (may be NULL) = their OR over all shards, or CMW_FLAG_PEER_TIMEOUT for
 * every query when a peer's flag did not arrive within `timeout_ms` (0 = the default, 2000 ms; the wait is bounded
 * so that a dead or out-of-step peer cannot hang this GPU) or when a peer published a different (B, k). */
int cmw_exchange_merge_ex(void* const* peer_bufs_host, int G, int rank, int max_batch, int max_k, int B, int k,
                          int k_out, uint32_t epoch, const double* scores64_local_dev, const int64_t* ids_local_dev,
                          const int32_t* flags_local_dev, float* out_scores_dev, int64_t* out_ids_dev,
                          double* out_scores64_dev, int32_t* out_flags_dev, int timeout_ms, void* stream);

/* ---- the two exchanges of the two-phase row-sharded search over NVLink peer memory instead of NCCL.
 * One peer buffer per rank of cmw_peer_gather_bytes(G, max_bytes_per_rank) bytes (cmw_peer_alloc / cmw_peer_open as
 * above).  cmw_peer_gather launches, stream-ordered and without host synchronisation: a send kernel that stores
 * `nbytes` of src_dev into this rank's slot of EVERY rank's buffer (st.global over NVLink 5 / NVSwitch), fences
 * system-wide and raises this rank's epoch flag on every peer; then a one-block kernel that holds the stream until
 * all G slots of this epoch have landed -- for at most timeout_ms (0 = 2 s), after which (or when a peer published a
 * different nbytes) CMW_FLAG_PEER_TIMEOUT is OR-ed into *status_dev (sticky: the exchange is out of step, rebuild
 * it).  *gathered_dev_out = the G slots, rank-major and contiguous (G * nbytes bytes) IN this rank's peer buffer:
 * what all-gather into one tensor would have produced, read in place by cmw_shard_kth / cmw_shard_merge_ex.  It
 * stays valid until the call after next (two regions alternate by epoch parity).  Every rank must call with the
 * same nbytes and the same epoch sequence 1, 2, 3, ...  (epoch != 0). */
size_t cmw_peer_gather_bytes(int G, size_t max_bytes_per_rank);
int cmw_peer_gather(void* const* peer_bufs_host, int G, int rank, size_t max_bytes_per_rank, const void* src_dev,
                    size_t nbytes, uint32_t epoch, int timeout_ms, int32_t* status_dev, void** gathered_dev_out,
                    void* stream);

/* ---- instrumentation ---- */
/* number of kernels this library has launched since load (all stores, all streams) */
int64_t cmw_kernel_launches(void);
 /* Per-phase device timing for roofline reports: while enabled, cmw_search brackets its phases with
 * CUDA events on the caller's stream.  cmw_profile_read synchronises those events and returns the
 * accumulated milliseconds since the last enable/read: ms[0] = filter kernels (K1 scan / K2 GEMM),
 * ms[1] = pool compaction, ms[2] = finalisation (K3 rescoring + select, or emit), ms[3] = query
 * preparation, ms[4] = cmw_shard_kth, ms[5] = cmw_shard_merge; counts[i] = kernels launched in that phase.
 * n = array length (<= 6). */
int cmw_profile_enable(int on);
int cmw_profile_read(double* ms, int64_t* counts, int n);
/* Process-wide tunables (defaults in parentheses):
 *   "scan_max_batch" (0)      batches up to this size use K1 (scan), larger ones K2 (GEMM)
 *   "gemm_enabled" (1), "gemm_2cta" (1), "gemm_2cta_min_batch" (128), "gemm_clc" (1)   K2 kernel selection
 *   "kprime" (0 = automatic)  candidates kept per query between slabs and handed to K3
 *   "strict_certificate" (1)  1 = rigorous residual bound behind the bf16 filter (K' = max(k + 108, 3k + 20));
 *                             0 = statistical bound, bf16_sigmas x sigma with u = 2^-8 (K' = max(k + 64, 2k))
 *   "bf16_eps" (0 = from the bound above), "bf16_sigmas" (8), "f32_eps" (0 = (D/32 + 12) 2^-24)   certificate bounds
 *   "repair" (2)              cmw_search_host repair chain for flagged queries: 0 off, 1 stage 1, 2 both stages
 *   "wide_dense" (1)          batches up to 32: first slab of up to 131072 rows through a scratch matrix, then the rest of
 *                             the corpus in one launch when the expected admissions fit the pool
 *   "pdl" (1)                 batches up to 32 over a collection that fits one wide first slab: the chain of small
 *                             dependent kernels that IS such a search is launched with programmatic stream
 *                             serialization (each kernel's launch and prologue overlap its predecessor's tail;
 *                             griddepcontrol.wait before the first dependent access)
 *   "scan_permute" (1)        K2 scans the row tiles in a stride permutation: every slab is a representative sample
 *                             of the corpus, the admission thresholds hold whatever order the corpus is stored in
 *   "host_overlap" (0)        pipelined host API: 1 = a ticket's finalisation runs next to the following ticket's filter (measured: no gain)
 *   "slab_growth" (0 = automatic)   cap on the geometric growth of the slabs
 *   "f16_bits" (11)           significand bits kept in fp16 tiles, 8..11, read when a store is created (fewer bits:
 *                             less multiplier power under the power cap, larger -- measured -- rounding residual)
 *   "pool_cap" (read-only)    candidate-pool slots per query */
int cmw_set_option(const char* name, double value);
double cmw_get_option(const char* name);

#ifdef __cplusplus
}
#endif
#endif /* CMW_DENSE_H_ */
