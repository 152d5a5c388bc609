// common.cuh -- shared declarations of libcmwdense.so (internal; the public surface is include/cmw_dense.h).
//
// Everything here is sm_100a-only device code plus the small amount of host glue the C ABI needs.
// No torch types, no CPU fallback: a missing CUDA device makes every compute entry point fail.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cmw_dense.h"

namespace cmw {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_kernel_launches;
extern std::atomic<int> g_pdl;
extern thread_local int t_pdl_search;

#define CMW_CUDA_OK(expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::cmw::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                  \
                             cudaGetErrorString(_e));                                            \
            return -2;                                                                           \
        }                                                                                        \
    } while (0)

#define CMW_REQUIRE(cond, ...)                                                                   \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            ::cmw::set_error(__VA_ARGS__);                                                       \
            return -1;                                                                           \
        }                                                                                        \
    } while (0)

#define CMW_LAUNCHED() (::cmw::g_kernel_launches.fetch_add(1, std::memory_order_relaxed))

#ifdef __CUDACC__
// kernel<<<grid, block, smem, stream>>>(args...), optionally as a programmatic dependent launch (ptx.cuh: pdl_wait)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && ::cmw::t_pdl_search != 0 && ::cmw::g_pdl.load(std::memory_order_relaxed) != 0) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---------------------------------------------------------------------------------------------
// candidate pool: the state shared by the filter kernels (K1 scan, K2 GEMM), the compaction
// kernel and the finalisation kernels.  One pool per query, resident in the caller's workspace.
//   scores[b][CAP] approximate (filter) scores, ids[b][CAP] LOCAL row numbers, cnt[b] entries
//   appended so far (may exceed CAP -> overflow), thr[b] admission threshold (score >= thr).
// ---------------------------------------------------------------------------------------------
constexpr int kPoolCap = 4096;      // entries per query
constexpr int kDenseSlabRows = 4096;  // first slab: every row is written (no threshold yet); = kPoolCap
constexpr int kMaxKPrime = 1024;
// Small batches take a WIDE first slab: its scores and ids go to a scratch matrix in the workspace, laid out as
// up to 32 pool-sized segments per query, instead of the pools (fewer slabs -- filter launches and compactions,
// 7-12 us each however little they do -- per search).  How many segments a search uses depends on K' (their
// survivors must fit one pool): 17 at k = 100, all 32 at k <= 50 -- a 100k-row collection searched for its top-20
// (the reference's production shape) is ONE slab: 6 kernels per search.
constexpr int kWideDenseRows = 131072;
constexpr int kWideSegments = kWideDenseRows / kPoolCap;
constexpr int kWideDenseMaxBatch = 32;
constexpr int kScanMaxQueries = 4;  // queries K1 scores per pass over the corpus

struct Pool {
    float* scores;   // [B, kPoolCap]
    int32_t* ids;    // [B, kPoolCap]
    int32_t* cnt;    // [B]
    float* thr;      // [B]
    int32_t* ovf;    // [B]  1 = some candidate was dropped because the pool was full
};

// per-row multiplier selector (see store.cu): which array turns a raw dot into the filter score
enum RowMul : int { ROWMUL_INV_NORM = 0, ROWMUL_NORM = 1, ROWMUL_LIVE = 2 };

// One in-flight ticket of the pipelined host API (cmw_search_host_submit / _wait): its own pinned staging
// block, device I/O block and search workspace, plus the events that chain copy-in -> compute -> copy-out.
constexpr int kHostSlots = 4;
struct HostSlot {
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    void* dev_io = nullptr;
    size_t dev_io_bytes = 0;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    cudaEvent_t ev_in = nullptr, ev_fork = nullptr, ev_compute = nullptr, ev_done = nullptr;
    bool busy = false;
    // the request, kept for cmw_search_host_wait
    const float* q_host = nullptr;
    int batch = 0, k = 0, metric = 0, mode = 0;
    float* out_scores = nullptr;
    int64_t* out_ids = nullptr;
    int32_t* out_flags = nullptr;
    bool out_pinned = false;
};

struct Store {
    int device = 0;
    int dim = 0;
    uint32_t flags = 0;
    int sm_count = 0;
    int64_t capacity = 0;
    int64_t rows = 0;
    int64_t dead = 0;
    int64_t id_offset = 0;
    size_t hbm_bytes = 0;
    float* f32 = nullptr;            // [capacity, dim] raw rows (CMW_STORE_F32)
    // [capacity, dim] L2-normalised rows as 16-bit tiles: bf16 (CMW_STORE_BF16) or fp16 (CMW_STORE_F16, `half_tiles`).
    // Same bytes and tensor-core rate; fp16 keeps 11 significand bits instead of 8, so the rounding residual the
    // exactness certificate has to absorb is 8x smaller -- and unit vectors never leave fp16's range
    // (results below its smallest normal, 2^-14, are flushed to zero at conversion and counted in the residual)
    __nv_bfloat16* bf16 = nullptr;
    bool half_tiles = false;
    int half_bits = 11;  // significand bits kept in fp16 tiles (Options::f16_bits at create time)
    float* inv_norm = nullptr;       // [cap4] 1/|c| (0 for a zero row, NaN when tombstoned)
    float* norm = nullptr;           // [cap4] |c|   (NaN when tombstoned)
    float* live = nullptr;           // [cap4] 1.0   (NaN when tombstoned)
    double* norm64 = nullptr;        // [capacity] |c| in fp64
    int32_t* kb_gid = nullptr;       // [capacity]
    // device scalars (float bits, rounded up): [0] max |c|, [1] max |c/|c||_4, [2] max |c/|c| - tile16(c/|c|)|_2
    // (the rounding residual of the 16-bit tiles: rigorous certificate), [3] max |c - tf32_trunc(c)|_2 / |c| (the
    // same for fp32 rows read by tf32 MMAs); [16..17] tombstone counter
    uint32_t* maxnorm_bits = nullptr;
    // tombstones per block of 256 rows, as the device counted them (dead_blk) and as a host prefix sum
    // (dead_prefix[i] = dead rows among the first 256*i): the slab schedule needs the number of LIVE rows a
    // slab has seen, not the number of rows
    uint32_t* dead_blk = nullptr;          // device [capacity / 256 + 1]
    std::vector<int64_t> dead_prefix;      // host   [capacity / 256 + 2]; empty = no tombstones
    // host-API resources (cmw_search_host / *_host variants)
    cudaStream_t stream = nullptr;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    void* dev_io = nullptr;
    size_t dev_io_bytes = 0;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // pipelined host API: copy streams either side of `stream` (the compute stream) and the ticket slots
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaStream_t tail = nullptr;  // finalisation of a ticket, next to the following ticket's filter
    HostSlot slots[kHostSlots];
    // TMA descriptors of the 16-bit tiles and of the fp32 tiles (K2; the latter feeds kind::tf32 MMAs on stores
    // that keep no 16-bit tiles), encoded at create time
    alignas(64) CUtensorMap tmap_bf16;
    bool tmap_ok = false;
    alignas(64) CUtensorMap tmap_f32;
    bool tmap_f32_ok = false;
    std::mutex host_mu;  // serialises the *_host entry points, which share the staging resources above
};

// options (cmw_set_option)
struct Options {
    // Certificate behind the bf16 tensor-core filter: a bound eps on |filter score - exact score| (cosine units).
    //   strict_certificate = 1 (default): RIGOROUS.  With q^ = q/|q|, c^ = c/|c| and q~, c~ their bf16 tiles,
    //       |q^.c^ - q~.c~| <= |q^ - q~| |c^| + |q~| |c^ - c~|  <=  r_q + (1 + r_q) R_c
    //     where r_q = |q^ - q~|_2 is computed per query by the prep kernel and R_c = max over the stored rows of
    //     |c^ - c~|_2 is tracked at ingest (both in fp64 from the values actually written), plus D * 2^-23 for
    //     the fp32 accumulation of the tensor pipe whatever its order and rounding mode.  Round-to-nearest bf16
    //     has unit round-off 2^-8, so r <= 2^-8 always; for dense embeddings r ~ 0.42 * 2^-8 and eps ~ 3.5e-3.
    //   strict_certificate = 0: STATISTICAL.  The rounding errors of the two operands are taken as independent,
    //     relative size <= u = 2^-8 each: sigma <= u sqrt(2/3) |q^|_4 |c^|_4, eps = bf16_sigmas * sigma.  ~4x
    //     tighter (fewer rows rescored) but a probabilistic statement that adversarially aligned roundings break.
    //   bf16_eps > 0: fixed override of either.
    double bf16_eps = 0;
    double bf16_sigmas = 8;
    double strict_certificate = 1;
    // fp32 FMA filter (K1): 0 = automatic, (D/32 + 12) * 2^-24 -- the length of one lane's FMA chain plus the
    // shuffle tree, the query scaling and the row multiplier, each one rounding of 2^-24 on |q||c| <= 1
    double f32_eps = 0;
    double kprime = 0;        // 0 = automatic
    // Batches up to this use K1 (scan), larger ones K2 (GEMM).  0 = K2 for every batch size: even at
    // batch 1 the tensor-core filter reads the bf16 tiles (half the bytes of the fp32 scan) and the
    // fp64 rescoring keeps the answer exact, so it is the faster exact path; K1 remains the fp32-filter
    // path (CMW_ALGO_SCAN), the fallback for failed certificates and the path of bf16-less stores.
    double scan_max_batch = 0;
    double gemm_enabled = 1;
    double gemm_2cta = 1;            // CTA-pair (cta_group::2) K2 kernel for large batches
    double gemm_2cta_min_batch = 128;  // padded batches of 128 / 192 gain 2-5 % from the halved query-operand traffic
    double gemm_clc = 1;             // CTA-pair kernel: dynamic item scheduling through cluster launch control
    // host API repair chain for flagged queries: 0 = off (flags are only reported), 1 = stage 1 only (same
    // filter, K' = 1024, overflow-proof slabs), 2 = also stage 2 (fp32 scan filter).  Exact ties larger than
    // K' - k straddling the k-th place stay flagged whatever the stage -- the ids are still the lowest of
    // the tie by construction -- so tie-heavy corpora may prefer 1.
    double repair = 2;
    // pipelined host API: 1 = a ticket's finalisation (fp64 rescoring + selection) runs on its own stream next
    // to the following ticket's filter; 0 = every ticket strictly after the previous one
    double host_overlap = 0;
    // batches up to 32: a 65536-row first slab through a scratch matrix, and the rest of the corpus in ONE launch
    // when the expected admissions fit the pool (two filter launches per search instead of four at 1M rows)
    double wide_dense = 1;
    // K2 scans the row tiles in a stride permutation, so that every slab is a representative sample of the corpus
    double scan_permute = 1;
    double slab_growth = 0;   // 0 = automatic ((cap - K') / (3 K'), at most 8); else the fixed growth factor
    // Significand bits kept in fp16 tiles (8..11, read when a store is created; its queries follow).  The step is
    // power-capped and the multiplier array's power grows with the operand width, so fewer bits buy tensor-core
    // clock at the price of a larger (measured, still rigorous) rounding residual: a tuning knob between bf16's
    // 8 bits and fp16's 11.
    double f16_bits = 11;
};
extern Options g_opt;

// ---------------------------------------------------------------------------------------------
// launchers implemented in the other translation units (all stream-ordered, return 0 / <0)
// ---------------------------------------------------------------------------------------------
struct ScanArgs {
    const void* rows;      // base pointer of the store's tiles (fp32 or bf16)
    int elt_bytes;         // 4 or 2
    int half_tiles;        // 2-byte elements: 1 = fp16, 0 = bf16
    int dim;
    const float* row_mul;  // per-row multiplier (indexed by LOCAL row)
    int64_t row_begin, row_end;
    const float* q;        // [nq, dim] prepared fp32 queries
    int nq;                // 1 .. kScanMaxQueries
    Pool pool;             // already offset to the first of the nq queries
    int dense;             // 1 = write every row at slot (row - row_begin)
    float* wide_scores;    // dense only: scratch [nq, wide_stride] for the wide first slab (NULL = the pools)
    int32_t* wide_ids;
    int wide_stride;
    int sm_count;
};
int launch_scan(const ScanArgs& a, cudaStream_t stream);

struct GemmArgs {
    const Store* store;
    const __nv_bfloat16* q_bf16;  // [bpad, dim] 16-bit query tile in the store's tile format
    int batch;                    // real queries
    int bpad;                     // padded to the N tile
    const float* row_mul;
    int64_t row_begin, row_end;
    Pool pool;
    int dense;
    float* wide_scores;           // dense only: scratch [batch, wide_stride] (NULL = the pools)
    int32_t* wide_ids;
    int wide_stride;
    const Options* opt;           // the options snapshot of this search (never g_opt: it may change mid-search)
    int tf32;                     // 1 = operands are the fp32 tiles / q_tf32, kind::tf32 MMAs (1-CTA kernel)
    const float* q_tf32;          // [bpad, dim] queries rounded to tf32
};
int launch_gemm(const GemmArgs& a, cudaStream_t stream);
bool gemm_supported(const Store* s);
bool gemm_tf32_supported(const Store* s);
bool gemm_scan_permuted(const Store* s, const Options& o, int bpad, bool tf32);

int launch_prep_queries(const float* q, int batch, int bpad, int dim, int metric, double* qn64, double* q4,
                        double* qres, float* q_f32, __nv_bfloat16* q_bf16, int half_tiles, float* q_tf32, Pool pool,
                        int dense_count, Pool seg, int wide_rows, cudaStream_t stream);

// how the certificate / rescoring cut bound is obtained (see Options::bf16_eps)
enum CertKind : int { CERT_FIXED = 0, CERT_STATISTICAL = 1, CERT_RIGOROUS = 2 };
struct CertParams {
    int kind;
    double eps_fixed;       // CERT_FIXED: the bound itself (cosine units)
    double sigmas;          // CERT_STATISTICAL: multiples of the sigma bound from q4 and the store's max row 4-norm
    double acc_slack;       // CERT_RIGOROUS / STATISTICAL: fp32 accumulation slack of the filter, D * 2^-23
    double tile_u;          // unit round-off of the 16-bit tile format: 2^-8 (bf16) or 2^-11 (fp16)
    const double* q4;       // [B] |q/|q||_4
    const double* qres;     // [B] |q^ - bf16(q^)|_2 relative to |q^| (the query tile's rounding residual)
    const double* qn64;     // [B] |q|
    const uint32_t* norms;  // store scalars: [0] max |c|, [1] max 4-norm, [2] / [3] max row residual (float bits)
    int res_slot;           // which residual applies: 2 = 16-bit tiles, 3 = fp32 rows read as tf32
    int metric;
    // row-sharded searches: per-query rescoring cut handed in from outside (the k-th best filter score over ALL
    // shards minus 2 eps); NULL = this store's own k-th filter score
    const float* global_kth;
};
int launch_pool_compact(Pool pool, int batch, int kprime, int final, cudaStream_t stream);
// wide first slab: the best kprime (and ties) of the 16 scratch segments of every query -> its pool
int launch_wide_select(Pool seg, Pool pool, int batch, int kprime, int final, cudaStream_t stream);
// exact fp64 rescoring of the pool's first min(cnt, kprime) entries + final (score desc, id asc)
// selection with the exactness certificate
int launch_rescore_select(const Store* s, Pool pool, int batch, int k, int kprime, int metric,
                          const float* q_raw, const CertParams& cert, double* exact_ws,
                          float* out_scores, int64_t* out_ids, double* out_scores64,
                          int32_t* out_flags, double* out_aux, cudaStream_t stream);
// bf16 mode: emit the pool's best k as they are
int launch_pool_emit(const Store* s, Pool pool, int batch, int k, const double* qn64, int metric, float* out_scores,
                     int64_t* out_ids, double* out_scores64, int32_t* out_flags, double* out_aux,
                     cudaStream_t stream);

// row-sharded search (pool.cu)
// packed per-shard result block (what all-gather 2 moves): f64 scores [B,k] | i64 ids [B,k] | f64 aux [B,2] |
// i32 flags [B]
struct ShardBlock {
    size_t scores, ids, aux, flags, total;
};
__host__ __device__ inline ShardBlock shard_block(int B, int k) {
    ShardBlock s;
    s.scores = 0;
    s.ids = (size_t)B * k * 8;
    s.aux = s.ids + (size_t)B * k * 8;
    s.flags = s.aux + (size_t)B * 16;
    s.total = (s.flags + (size_t)B * 4 + 255) / 256 * 256;
    return s;
}

int launch_pool_topk_scores(Pool pool, int batch, int k, float* out, cudaStream_t stream);
int launch_shard_kth(const float* gathered, int G, int B, int k, float* out_kth, cudaStream_t stream);
int launch_shard_merge(const void* blocks, int G, int B, int k, int k_out, float* out_scores, int64_t* out_ids,
                       double* out_scores64, int32_t* out_flags, const int32_t* peer_status, cudaStream_t stream);

// host-API resources (store.cu)
int ensure_pinned(Store* s, size_t bytes);
int ensure_dev_io(Store* s, size_t bytes);
int ensure_ws(Store* s, size_t bytes);
int get_stream(Store* s, cudaStream_t* out);
// pageable host -> pinned block -> device, pipelined over a few host threads (all pieces on `stream`)
int stage_h2d(int device, uint8_t* pin, uint8_t* dev, const uint8_t* src, size_t bytes, cudaStream_t stream);

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: remember what was set where
struct SmemAttrCache {
    size_t set[64] = {};
    bool needs(size_t bytes) const {
        int dev = 0;
        cudaGetDevice(&dev);
        return bytes > set[dev & 63];
    }
    void done(size_t bytes) {
        int dev = 0;
        cudaGetDevice(&dev);
        set[dev & 63] = bytes;
    }
};

inline int next_pow2_host(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t f32_orderable(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending uint <=> ascending float
}
__device__ __forceinline__ float f32_from_orderable(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ uint64_t f64_orderable(double d) {
    uint64_t u = (uint64_t)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_orderable(uint64_t u) {
    return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
}
// Pool scores are in FILTER UNITS: the score of the L2-normalised query -- the cosine, or for the inner product
// |c| cos = (exact score) / |q|.  (Normalising the query for the filter in both metrics keeps the 16-bit tiles
// inside fp16's range whatever the scale of the caller's vectors.)  qscale turns filter units into exact units.
__device__ __forceinline__ double cert_qscale(const CertParams& c, int b) {
    return c.metric == CMW_METRIC_IP ? c.qn64[b] : 1.0;
}
// bound, in filter units, on |filter score - exact score| for query b and ANY stored row (Options::strict_certificate)
__device__ __forceinline__ double cert_eps(const CertParams& c, int b) {
    double e = c.eps_fixed;
    if (c.kind != CERT_FIXED) {
        const double rq = c.qres[b];
        const double rc = (double)__uint_as_float(c.norms[c.res_slot]);
        const double rigorous = rq + (1.0 + rq) * rc;  // Cauchy-Schwarz on the two rounding residuals
        e = rigorous;
        if (c.kind == CERT_STATISTICAL) {
            // independent roundings of relative size <= u
            const double bound = c.sigmas * c.tile_u * 0.816496580927726 * c.q4[b] *
                                 (double)__uint_as_float(c.norms[1]);
            if (bound < e) e = bound;
        }
        e += c.acc_slack * (1.0 + rq) * (1.0 + rc) + 4e-7;
    }
    if (c.metric == CMW_METRIC_IP) e *= (double)__uint_as_float(c.norms[0]);  // rows of norm up to max |c|
    return e;
}
// fp32 -> one 16-bit tile element; `stored` = the value the tile now holds (what the residuals are measured from)
// half_tiles: 0 = bf16, else fp16 keeping that many significand bits (11 = all)
__device__ __forceinline__ uint16_t to_tile16(float x, int half_tiles, float& stored) {
    if (half_tiles) {
        __half h = __float2half_rn(x);
        if (half_tiles < 11) {  // round the significand to fewer bits (carry into the exponent is what IEEE wants)
            const unsigned drop = 11u - (unsigned)half_tiles;
            unsigned short u = __half_as_ushort(h);
            u = (unsigned short)((u + (1u << (drop - 1))) & ~((1u << drop) - 1u));
            h = __ushort_as_half(u);
        }
        float v = __half2float(h);
        if (fabsf(v) < 6.103515625e-05f) {  // below 2^-14: no subnormals in the tiles, whatever the MMA does with them
            h = __ushort_as_half((unsigned short)0);
            v = 0.f;
        }
        stored = v;
        return __half_as_ushort(h);
    }
    const __nv_bfloat16 bv = __float2bfloat16_rn(x);
    stored = __bfloat162float(bv);
    return __bfloat16_as_ushort(bv);
}
// key whose ASCENDING order is (score descending, id ascending)
__device__ __forceinline__ uint64_t desc_key(float score, uint32_t id) {
    if (score == 0.f) score = 0.f;  // -0.0 and +0.0 are the same score
    return ((uint64_t)(~f32_orderable(score)) << 32) | (uint64_t)id;
}
__device__ __forceinline__ float desc_key_score(uint64_t key) {
    return f32_from_orderable(~(uint32_t)(key >> 32));
}

// In-place ascending bitonic sort of n (power of two) 64-bit keys in shared memory by the whole
// CTA.  Ends with a __syncthreads().
__device__ __forceinline__ void bitonic_sort_u64(uint64_t* keys, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                uint64_t a = keys[lo], b = keys[hi];
                if ((a > b) == up) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
        }
    }
    __syncthreads();
}

// (hi, lo) 128-bit keys, ascending lexicographic
__device__ __forceinline__ void bitonic_sort_u128(uint64_t* hi_keys, uint64_t* lo_keys, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                uint64_t ah = hi_keys[lo], bh = hi_keys[hi];
                uint64_t al = lo_keys[lo], bl = lo_keys[hi];
                bool gt = (ah > bh) || (ah == bh && al > bl);
                if (gt == up) {
                    hi_keys[lo] = bh;
                    hi_keys[hi] = ah;
                    lo_keys[lo] = bl;
                    lo_keys[hi] = al;
                }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
#endif  // __CUDACC__

}  // namespace cmw
