// api.cu -- cmw_search / cmw_search_host: orchestration of one batched top-k search.
//
// Replaces ChromaStore.similarity_search_async -> collection.query(query_embeddings=[q], n_results=k)
// (rag_engine/storage/vector_store.py:54-66 of the reference) for a whole batch of query vectors.
//
// Pipeline (all stream-ordered, no host synchronisation in cmw_search):
//   prep queries (fp64 norms, scaled fp32 + bf16 copies)
//   for each slab of rows (first slab dense, later slabs growing geometrically):
//       filter kernel (K1 scan or K2 tcgen05 GEMM) admits rows with score >= thr into the pools
//       compaction keeps the best K' per query and raises thr
//   (batches <= 32: a first slab of up to 131072 rows through a scratch matrix + two-level selection, then the rest of
//    the corpus in one launch when its expected admissions fit the pool)
//   cmw_search_host / _submit / _wait: the same behind H2D / D2H copies, blocking or pipelined
//   F32_EXACT: fp64 rescoring of the K' survivors from the fp32 tiles + final selection + certificate
//   BF16:      emit the pool's best k
#include <string.h>

#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace cmw {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct WsLayout {
    size_t qn64, q4, qres, q_f32, q_bf16, q_tf32, pool_scores, pool_ids, pool_cnt, pool_thr, pool_ovf, exact, wide, total;
    int bpad;
};

// Padded batch = n_groups query groups of `nt` columns each (gemm_group_width, gemm_common.cuh).  Up to 256 queries:
// one group, padded to 16.  Beyond: as many groups as 256-wide ones would take, but only as wide as needed, in
// steps of 64 -- 384 queries are two groups of 192, not two of 256 with a quarter of the tensor work on padding.
static int pad_batch(int batch) {
    if (batch <= 16) return 16;
    if (batch <= 256) return (int)align_up((size_t)batch, 16);
    const int groups = (batch + 255) / 256;
    const int nt = (int)align_up((size_t)((batch + groups - 1) / groups), 64);
    return nt * groups;
}

static WsLayout ws_layout(int dim, int batch, int kprime, bool tf32 = true) {
    WsLayout w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    w.bpad = pad_batch(batch);
    w.qn64 = take((size_t)batch * sizeof(double));
    w.q4 = take((size_t)batch * sizeof(double));
    w.qres = take((size_t)batch * sizeof(double));
    w.q_f32 = take((size_t)batch * dim * sizeof(float));
    w.q_bf16 = take((size_t)w.bpad * dim * sizeof(__nv_bfloat16));
    w.q_tf32 = take(tf32 ? (size_t)w.bpad * dim * sizeof(float) : 0);  // only stores that filter through tf32 MMAs
    w.pool_scores = take((size_t)batch * kPoolCap * sizeof(float));
    w.pool_ids = take((size_t)batch * kPoolCap * sizeof(int32_t));
    w.pool_cnt = take((size_t)w.bpad * sizeof(int32_t));
    w.pool_thr = take((size_t)w.bpad * sizeof(float));
    w.pool_ovf = take((size_t)w.bpad * sizeof(int32_t));
    w.exact = take((size_t)batch * kprime * sizeof(double));
    // scratch of the wide first slab, small batches only: scores + ids in kWideSegments pool-sized segments per query, and
    // the segments' cnt / thr / ovf
    w.wide = take(batch <= kWideDenseMaxBatch
                      ? (size_t)batch * kWideDenseRows * 8 + (size_t)batch * kWideSegments * 12 + 1024
                      : 0);
    w.total = off;
    return w;
}

// does this search filter with fp32 rows through kind::tf32 MMAs?  Exact mode only (the tf32 filter has no
// approximate-mode meaning), when asked for (CMW_ALGO_GEMM_TF32) or when the store keeps no 16-bit tiles
static bool want_tf32(const Store* s, int mode) {
    if ((mode & 0xff) != CMW_MODE_F32_EXACT || !gemm_tf32_supported(s)) return false;
    const int algo = mode & 0xff00;
    return algo == CMW_ALGO_GEMM_TF32 || (algo != CMW_ALGO_SCAN && !gemm_supported(s));
}

static bool use_gemm(const Store* s, const Options& opt, int batch, int mode) {
    const int algo = mode & 0xff00;
    if (algo == CMW_ALGO_SCAN) return false;
    if (!gemm_supported(s) && !want_tf32(s, mode)) return false;
    if (algo == CMW_ALGO_GEMM || algo == CMW_ALGO_GEMM_TF32) return true;
    return opt.gemm_enabled != 0 && batch > (int)opt.scan_max_batch;
}

// K': candidates kept per query between slabs and handed to K3.  The certificate needs the K'-th best filter
// score to lie more than eps below the k-th exact score, so K' follows the bound in use:
//   fp32 scan filter (eps ~ 3.6e-6): k + 28;
//   fp16 tiles, rigorous bound (eps ~ 6.5e-4 at D = 1536), and either format with the statistical bound
//   (eps ~ 8.6e-4 for bf16): max(k + 64, 2k);
//   bf16 tiles, rigorous bound (eps ~ 3.5e-3): max(k + 108, 3k + 20) -- at 1M iid rows the k-th and the K'-th
//   score are then ~7.8e-3 apart for k = 100 (5 standard deviations of the order statistics over eps).
// k above ~330 (bf16 tiles) / ~500 (fp16) cannot be certified behind the tensor-core filter (K' is capped at
// 1024): such queries are flagged and the host API repairs them through the fp32 scan.
static int pick_kprime(const Options& opt, int k, int mode, bool gemm, bool half_tiles) {
    int kp;
    if (mode & CMW_KPRIME_MAX) {
        kp = kMaxKPrime;
    } else if (opt.kprime > 0) {
        kp = (int)opt.kprime;
    } else if ((mode & 0xff) == CMW_MODE_BF16) {
        kp = k;
    } else if (gemm && opt.strict_certificate != 0 && opt.bf16_eps <= 0 && !half_tiles) {
        kp = (3 * k + 20 > k + 108) ? 3 * k + 20 : k + 108;
    } else if (gemm) {
        kp = (k + 64 > 2 * k) ? k + 64 : 2 * k;
    } else {
        kp = k + 28;
    }
    if (kp < k) kp = k;
    kp = (int)align_up((size_t)kp, 32);
    if (kp > kMaxKPrime) kp = kMaxKPrime;
    return kp;
}

// Layout of one host call's I/O block (the same offsets in the pinned staging buffer and in its device twin):
// queries | scores | ids | flags.
struct HostIo {
    size_t q_bytes, sc_bytes, id_bytes, fl_bytes, total;
};
static HostIo host_io(const Store* s, int batch, int k) {
    HostIo io;
    io.q_bytes = align_up((size_t)batch * s->dim * sizeof(float), 256);
    io.sc_bytes = align_up((size_t)batch * k * sizeof(float), 256);
    io.id_bytes = align_up((size_t)batch * k * sizeof(int64_t), 256);
    io.fl_bytes = align_up((size_t)batch * sizeof(int32_t), 256);
    io.total = io.q_bytes + io.sc_bytes + io.id_bytes + io.fl_bytes;
    return io;
}

// page-locked caller buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory) are used for DMA
// directly; pageable ones go through a pinned staging buffer
static bool is_pinned(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

static int ensure_buffer(void** ptr, size_t* have, size_t bytes, bool host) {
    if (*have >= bytes) return 0;
    if (*ptr) {
        if (host) cudaFreeHost(*ptr);
        else cudaFree(*ptr);
    }
    *ptr = nullptr;
    *have = 0;
    if (host) CMW_CUDA_OK(cudaMallocHost(ptr, bytes));
    else CMW_CUDA_OK(cudaMalloc(ptr, bytes));
    *have = bytes;
    return 0;
}

static int ensure_blocking_buffers(cmw_store* h, Store* s, const HostIo& io, int batch, int k, int mode) {
    int rc;
    if ((rc = ensure_pinned(s, io.total))) return rc;
    if ((rc = ensure_dev_io(s, io.total))) return rc;
    return ensure_ws(s, cmw_search_workspace_bytes(h, batch, k, mode));
}

// One blocking search of nb <= batch queries through the store's own staging buffers (offsets of `io`, which
// was laid out for the whole batch): H2D, cmw_search, D2H, synchronise.  direct_* non-NULL = page-locked
// caller buffers that receive scores / ids by DMA; flags always land in the pinned buffer.
static int run_blocking(cmw_store* h, Store* s, cudaStream_t stream, const HostIo& io, const float* q_src, int nb,
                        int k, int metric, int run_mode, float* direct_scores, int64_t* direct_ids) {
    uint8_t* pin = reinterpret_cast<uint8_t*>(s->pinned);
    uint8_t* dev = reinterpret_cast<uint8_t*>(s->dev_io);
    const size_t qb = (size_t)nb * s->dim * sizeof(float);
    if (is_pinned(q_src)) {
        CMW_CUDA_OK(cudaMemcpyAsync(dev, q_src, qb, cudaMemcpyHostToDevice, stream));
    } else if (stage_h2d(s->device, pin, dev, reinterpret_cast<const uint8_t*>(q_src), qb, stream)) {
        // pageable -> pinned -> device, pipelined: a few host threads copy 1 MB pieces into the pinned
        // staging buffer and enqueue each piece's H2D as soon as it is staged, so the DMA of one
        // piece overlaps the memcpy of the next (all pieces precede the search on the same stream)
        return -2;
    }
    const size_t sc_off = io.q_bytes, id_off = io.q_bytes + io.sc_bytes, fl_off = id_off + io.id_bytes;
    int r = cmw_search(h, reinterpret_cast<const float*>(dev), nb, k, metric, run_mode,
                       reinterpret_cast<float*>(dev + sc_off), reinterpret_cast<int64_t*>(dev + id_off), nullptr,
                       reinterpret_cast<int32_t*>(dev + fl_off), s->ws, s->ws_bytes, stream);
    if (r) return r;
    if (direct_scores && direct_ids) {
        CMW_CUDA_OK(cudaMemcpyAsync(direct_scores, dev + sc_off, (size_t)nb * k * sizeof(float),
                                    cudaMemcpyDeviceToHost, stream));
        CMW_CUDA_OK(cudaMemcpyAsync(direct_ids, dev + id_off, (size_t)nb * k * sizeof(int64_t),
                                    cudaMemcpyDeviceToHost, stream));
        CMW_CUDA_OK(cudaMemcpyAsync(pin + fl_off, dev + fl_off, io.fl_bytes, cudaMemcpyDeviceToHost, stream));
    } else {
        // one D2H for scores + ids + flags (contiguous in dev_io)
        CMW_CUDA_OK(cudaMemcpyAsync(pin + sc_off, dev + sc_off, io.sc_bytes + io.id_bytes + io.fl_bytes,
                                    cudaMemcpyDeviceToHost, stream));
    }
    CMW_CUDA_OK(cudaStreamSynchronize(stream));
    return 0;
}

// Repair chain for flagged queries (a failed certificate: many scores within eps of the k-th; or a
// pool overflow: adversarial row order).  `first_flags` = the flags of the first pass (copied out before
// the staging buffers are reused).  Each stage re-runs only the queries still flagged, batched:
//   1. the same filter with the largest K' (1024) on the overflow-proof slab schedule -- a handful
//      of milliseconds whatever the number of queries when the filter is K2;
//   2. (exact mode) the fp32 scan filter, certificate bound three orders of magnitude tighter, also
//      with the largest K' and the overflow-proof schedule.
// bf16 mode can only be flagged by an overflow: stage 1 settles it.
static int repair_flagged(cmw_store* h, Store* s, cudaStream_t stream, const HostIo& io, const float* queries_host,
                          int batch, int k, int metric, int mode, float* out_scores_host, int64_t* out_ids_host,
                          int32_t* out_flags_host, const int32_t* first_flags) {
    std::vector<int> redo;
    for (int b = 0; b < batch; ++b) {
        if (out_flags_host) out_flags_host[b] = first_flags[b];
        if (first_flags[b] & CMW_FLAG_UNCERTIFIED) redo.push_back(b);
    }
    if (redo.empty()) return 0;
    const uint8_t* pin = reinterpret_cast<const uint8_t*>(s->pinned);
    const float* sc = reinterpret_cast<const float*>(pin + io.q_bytes);
    const int64_t* id = reinterpret_cast<const int64_t*>(pin + io.q_bytes + io.sc_bytes);
    const int32_t* fl = reinterpret_cast<const int32_t*>(pin + io.q_bytes + io.sc_bytes + io.id_bytes);
    const bool exact = (mode & 0xff) == CMW_MODE_F32_EXACT;
    const bool first_was_gemm = use_gemm(s, g_opt, batch, mode);
    if ((mode & 0xff00) == CMW_ALGO_GEMM_TF32) mode = (mode & ~0xff00) | CMW_ALGO_AUTO;  // repairs choose their own filter
    int stages[2];
    int nstage = 0;
    if (g_opt.repair >= 1 && !((mode & CMW_SLABS_SAFE) && (mode & CMW_KPRIME_MAX)))
        stages[nstage++] = mode | CMW_SLABS_SAFE | CMW_KPRIME_MAX;
    if (g_opt.repair >= 2 && exact && first_was_gemm)
        stages[nstage++] = CMW_MODE_F32_EXACT | CMW_ALGO_SCAN | CMW_SLABS_SAFE | CMW_KPRIME_MAX;
    for (int st = 0; st < nstage && !redo.empty(); ++st) {
        std::vector<float> q2((size_t)redo.size() * s->dim);
        for (size_t i = 0; i < redo.size(); ++i)
            memcpy(q2.data() + i * s->dim, queries_host + (size_t)redo[i] * s->dim, (size_t)s->dim * sizeof(float));
        int rc = run_blocking(h, s, stream, io, q2.data(), (int)redo.size(), k, metric, stages[st], nullptr, nullptr);
        if (rc) return rc;
        std::vector<int> still;
        for (size_t i = 0; i < redo.size(); ++i) {
            const int b = redo[i];
            memcpy(out_scores_host + (size_t)b * k, sc + i * k, (size_t)k * sizeof(float));
            memcpy(out_ids_host + (size_t)b * k, id + i * k, (size_t)k * sizeof(int64_t));
            if (out_flags_host) out_flags_host[b] = fl[i];
            if (fl[i] & CMW_FLAG_UNCERTIFIED) still.push_back(b);
        }
        redo.swap(still);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// optional per-phase timing (cmw_profile_enable / cmw_profile_read)
// ---------------------------------------------------------------------------------------------
struct PhaseRecord {
    int phase;
    cudaEvent_t e0, e1;
    long long launches;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<PhaseRecord> g_prof_records;
static std::vector<cudaEvent_t> g_prof_free;
constexpr int kPhases = 6;  // filter, compact, finalize, prep, shard kth, shard merge
static double g_prof_ms[kPhases] = {};
static long long g_prof_counts[kPhases] = {};

static cudaEvent_t prof_event() {
    if (!g_prof_free.empty()) {
        cudaEvent_t e = g_prof_free.back();
        g_prof_free.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct PhaseTimer {
    bool on;
    PhaseRecord rec;
    cudaStream_t stream;
    long long start;
    PhaseTimer(int phase, cudaStream_t st) : on(g_prof_on), stream(st) {
        if (!on) return;
        std::lock_guard<std::mutex> lk(g_prof_mu);
        rec.phase = phase;
        rec.e0 = prof_event();
        rec.e1 = prof_event();
        start = g_kernel_launches.load();
        cudaEventRecord(rec.e0, stream);
    }
    void stop() {
        if (!on) return;
        on = false;
        cudaEventRecord(rec.e1, stream);
        rec.launches = g_kernel_launches.load() - start;
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof_records.push_back(rec);
    }
    ~PhaseTimer() { stop(); }
};

static void prof_drain() {
    for (auto& r : g_prof_records) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            g_prof_ms[r.phase] += ms;
            g_prof_counts[r.phase] += r.launches;
        }
        g_prof_free.push_back(r.e0);
        g_prof_free.push_back(r.e1);
    }
    g_prof_records.clear();
}

}  // namespace cmw

using namespace cmw;

extern "C" {

int cmw_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_drain();
    for (int i = 0; i < kPhases; ++i) {
        g_prof_ms[i] = 0;
        g_prof_counts[i] = 0;
    }
    g_prof_on = on != 0;
    return 0;
}

int cmw_profile_read(double* ms, int64_t* counts, int n) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_drain();
    for (int i = 0; i < n && i < kPhases; ++i) {
        if (ms) ms[i] = g_prof_ms[i];
        if (counts) counts[i] = g_prof_counts[i];
        g_prof_ms[i] = 0;
        g_prof_counts[i] = 0;
    }
    return 0;
}

size_t cmw_search_workspace_bytes(const cmw_store* h, int batch, int k, int mode) {
    if (!h || batch <= 0 || k <= 0) return 0;
    const Store* s = reinterpret_cast<const Store*>(h);
    (void)mode;
    (void)k;
    // sized for the largest K' so that option changes between the query and the call stay safe
    return ws_layout(s->dim, batch, kMaxKPrime, gemm_tf32_supported(s)).total;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// The search itself, in two stream-ordered halves:
//   filter half  -- query preparation, the filter slabs and the pool compactions; ends with every pool sorted
//                   by filter score and truncated to K'
//   finish half  -- F32_EXACT: fp64 rescoring + selection + certificate; BF16: emit the pool's best k
// cmw_search runs both back to back.  The row-sharded search (cmw_search_filter / cmw_shard_kth /
// cmw_search_finish / cmw_shard_merge) puts the cross-shard exchange of the k-th filter score between them.
// ---------------------------------------------------------------------------------------------
struct SearchPlan {
    Store* s;
    Options opt;  // snapshot: cmw_set_option during a search must not change what its later slabs do
    int batch, k, metric, mode, base_mode, kprime;
    bool gemm;
    bool tf32;  // the gemm filter reads the fp32 tiles through kind::tf32 MMAs
    WsLayout w;
    double *qn64, *q4, *qres, *exact;
    float* q_f32;
    float* q_tf32;
    __nv_bfloat16* q_bf16;
    Pool pool;
    uint8_t* ws;
};

static int make_plan(SearchPlan& pl, cmw_store* h, const float* queries_dev, int batch, int k, int metric, int mode,
                     void* ws_dev, size_t ws_bytes, const char* who) {
    CMW_REQUIRE(h != nullptr, "%s: store is NULL", who);
    Store* s = reinterpret_cast<Store*>(h);
    CMW_REQUIRE(batch > 0 && queries_dev, "%s: bad arguments", who);
    CMW_REQUIRE(k >= 1 && k <= kMaxKPrime, "%s: k must be in [1, %d], got %d", who, kMaxKPrime, k);
    CMW_REQUIRE(metric == CMW_METRIC_COSINE || metric == CMW_METRIC_IP, "%s: unknown metric %d", who, metric);
    const int base_mode = mode & 0xff;
    CMW_REQUIRE(base_mode == CMW_MODE_F32_EXACT || base_mode == CMW_MODE_BF16, "%s: unknown mode %d", who, base_mode);
    CMW_REQUIRE((reinterpret_cast<uintptr_t>(queries_dev) & 15) == 0, "%s: queries_dev must be 16-byte aligned", who);
    if (base_mode == CMW_MODE_F32_EXACT)
        CMW_REQUIRE(s->f32 != nullptr, "%s: CMW_MODE_F32_EXACT needs a CMW_STORE_F32 store", who);
    if (base_mode == CMW_MODE_BF16)
        CMW_REQUIRE(s->bf16 != nullptr, "%s: CMW_MODE_BF16 needs a store with 16-bit tiles (CMW_STORE_BF16 / _F16)", who);
    CMW_CUDA_OK(cudaSetDevice(s->device));
    pl.s = s;
    pl.opt = g_opt;
    pl.batch = batch;
    pl.k = k;
    pl.metric = metric;
    pl.mode = mode;
    pl.base_mode = base_mode;
    pl.gemm = use_gemm(s, pl.opt, batch, mode);
    pl.tf32 = pl.gemm && want_tf32(s, mode);
    if ((mode & 0xff00) == CMW_ALGO_GEMM || (mode & 0xff00) == CMW_ALGO_GEMM_TF32)
        CMW_REQUIRE(pl.gemm && (pl.tf32 || (mode & 0xff00) == CMW_ALGO_GEMM),
                    "%s: a tcgen05 filter was requested but is unavailable (store without the tiles it reads, "
                    "no TMA descriptor, or tf32 outside CMW_MODE_F32_EXACT)", who);
    pl.kprime = pick_kprime(pl.opt, k, mode, pl.gemm, pl.tf32 || (s->half_tiles && s->half_bits >= 10));
    pl.w = ws_layout(s->dim, batch, pl.kprime, gemm_tf32_supported(s));
    CMW_REQUIRE(ws_dev != nullptr && ws_bytes >= pl.w.total, "%s: workspace too small (%zu bytes given, %zu needed)",
                who, ws_bytes, pl.w.total);
    CMW_REQUIRE((reinterpret_cast<uintptr_t>(ws_dev) & 255) == 0, "%s: workspace must be 256-byte aligned", who);
    uint8_t* ws = reinterpret_cast<uint8_t*>(ws_dev);
    pl.ws = ws;
    pl.qn64 = reinterpret_cast<double*>(ws + pl.w.qn64);
    pl.q4 = reinterpret_cast<double*>(ws + pl.w.q4);
    pl.qres = reinterpret_cast<double*>(ws + pl.w.qres);
    pl.q_f32 = reinterpret_cast<float*>(ws + pl.w.q_f32);
    pl.q_bf16 = reinterpret_cast<__nv_bfloat16*>(ws + pl.w.q_bf16);
    pl.q_tf32 = reinterpret_cast<float*>(ws + pl.w.q_tf32);
    pl.pool.scores = reinterpret_cast<float*>(ws + pl.w.pool_scores);
    pl.pool.ids = reinterpret_cast<int32_t*>(ws + pl.w.pool_ids);
    pl.pool.cnt = reinterpret_cast<int32_t*>(ws + pl.w.pool_cnt);
    pl.pool.thr = reinterpret_cast<float*>(ws + pl.w.pool_thr);
    pl.pool.ovf = reinterpret_cast<int32_t*>(ws + pl.w.pool_ovf);
    pl.exact = reinterpret_cast<double*>(ws + pl.w.exact);
    return 0;
}

// live rows among the first `pos` rows in STORAGE order (pos a multiple of 256, or the end of the store); a lower
// bound when pos falls inside a block
static int64_t live_before(const Store* s, int64_t pos) {
    if (s->dead_prefix.empty()) return pos;
    const size_t nblk = s->dead_prefix.size() - 1;  // blocks the prefix covers; nothing is dead beyond them
    size_t blk = (size_t)(pos / 256);
    int64_t dead;
    if (blk >= nblk) {
        dead = s->dead_prefix[nblk];
    } else {
        dead = s->dead_prefix[blk];
        if (pos % 256) dead += s->dead_prefix[blk + 1] - s->dead_prefix[blk];
    }
    const int64_t live = pos - dead;
    return live > 0 ? live : 0;
}

static int run_filter_half(const SearchPlan& pl, const float* queries_dev, cudaStream_t stream) {
    Store* s = pl.s;
    const Options& opt = pl.opt;
    const int batch = pl.batch, kprime = pl.kprime, metric = pl.metric, mode = pl.mode;
    const bool gemm = pl.gemm;
    const WsLayout& w = pl.w;
    Pool pool = pl.pool;
    int rc;
    const int64_t rows = s->rows;
    // First slab: every row's score is kept (no threshold exists yet).  Large batches write it straight into
    // the pools (4096 rows = the pool capacity).  Small batches take a WIDE first slab of up to 131072 rows
    // whose scores go to a scratch matrix of up to 32 pool-sized segments per query, from which a two-level
    // selection fills the pools.  (The segments' survivors, kprime and ties each, must fit one pool: fewer
    // segments for a large kprime.)
    int wide_segs = (kPoolCap - 256) / kprime;
    if (wide_segs > kWideSegments) wide_segs = kWideSegments;
    const bool wide = opt.wide_dense != 0 && batch <= kWideDenseMaxBatch && rows > kDenseSlabRows && wide_segs >= 2 &&
                      !(mode & CMW_SLABS_SAFE);
    const int64_t wide_rows = (int64_t)wide_segs * kPoolCap;
    const int64_t slab0 = wide ? (rows < wide_rows ? rows : wide_rows)
                               : (rows < kDenseSlabRows ? rows : (int64_t)kDenseSlabRows);
    // Programmatic dependent launch of this search's kernels (ptx.cuh: pdl_wait) where the chain of small kernels IS
    // the search: a small batch over a collection that is one wide slab (6 kernels; 100k rows, k = 20: -2 % at batch
    // 1, -7 % at batch 16-32).  With a second, long filter launch behind the first slab it measured neutral at
    // batch 1 and 1-2 % slower at batch 4-32 (benchmarks/pdl_ab.py), so those searches launch plainly.
    t_pdl_search = (wide && slab0 == rows) ? 1 : 0;
    // What the schedule below reasons about is the number of LIVE rows the slabs so far have seen: the admission
    // threshold is the kprime-th best of those.  K2 scans a stride permutation of the tiles, so its slabs see the
    // store-wide live fraction; K1 (and small stores) scan in storage order, where the tombstones may sit in one
    // block -- a re-indexed collection is exactly that -- and the per-block counts give the exact number.
    const bool permuted = gemm && gemm_scan_permuted(s, opt, w.bpad, pl.tf32);
    const double live_frac = rows > 0 ? (double)(rows - s->dead) / (double)rows : 1.0;
    auto live_seen = [&](int64_t seen) -> double {
        if (s->dead == 0) return (double)seen;
        return permuted ? (double)seen * live_frac : (double)live_before(s, seen);
    };
    const double live0 = live_seen(slab0);
    const double live_all = (double)(rows - s->dead);
    // expected pool fill if everything after the first slab went through one launch
    const double rest_fill = (live_all - live0) * kprime / (live0 > 1.0 ? live0 : 1.0) + kprime;
    float* wide_scores = nullptr;
    int32_t* wide_ids = nullptr;
    Pool seg = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (wide) {
        uint8_t* base = pl.ws + w.wide;
        wide_scores = reinterpret_cast<float*>(base);
        wide_ids = reinterpret_cast<int32_t*>(base + (size_t)batch * kWideDenseRows * 4);
        uint8_t* tail = base + (size_t)batch * kWideDenseRows * 8;
        const size_t seg_bytes = align_up((size_t)batch * kWideSegments * 4, 256);
        seg.scores = wide_scores;
        seg.ids = wide_ids;
        seg.cnt = reinterpret_cast<int32_t*>(tail);
        seg.thr = reinterpret_cast<float*>(tail + seg_bytes);
        seg.ovf = reinterpret_cast<int32_t*>(tail + 2 * seg_bytes);
    }
    {
        PhaseTimer t(3, stream);
        if ((rc = launch_prep_queries(queries_dev, batch, w.bpad, s->dim, metric, pl.qn64, pl.q4, pl.qres, pl.q_f32,
                                      (gemm && !pl.tf32) ? pl.q_bf16 : nullptr, s->half_tiles ? s->half_bits : 0,
                                      pl.tf32 ? pl.q_tf32 : nullptr, pool, wide ? 0 : (int)slab0, seg, (int)slab0,
                                      stream)))
            return rc;
    }

    // which tiles the filter reads, and the per-row multiplier that goes with them
    const bool filter_bf16 = (gemm && !pl.tf32) || pl.base_mode == CMW_MODE_BF16;
    const void* tiles = filter_bf16 ? (const void*)s->bf16 : (const void*)s->f32;
    const float* row_mul;
    if (filter_bf16) row_mul = (metric == CMW_METRIC_COSINE) ? s->live : s->norm;  // rows pre-normalised
    else row_mul = (metric == CMW_METRIC_COSINE) ? s->inv_norm : s->live;          // raw rows

    auto run_filter = [&](int64_t r0, int64_t r1, int dense) -> int {
        PhaseTimer t(0, stream);
        if (gemm) {
            GemmArgs g;
            g.store = s;
            g.q_bf16 = pl.q_bf16;
            g.batch = batch;
            g.bpad = w.bpad;
            g.row_mul = row_mul;
            g.row_begin = r0;
            g.row_end = r1;
            g.pool = pool;
            g.dense = dense;
            g.wide_scores = dense ? wide_scores : nullptr;
            g.wide_ids = wide_ids;
            g.wide_stride = kWideDenseRows;
            g.opt = &opt;
            g.tf32 = pl.tf32 ? 1 : 0;
            g.q_tf32 = pl.q_tf32;
            return launch_gemm(g, stream);
        }
        for (int b0 = 0; b0 < batch; b0 += kScanMaxQueries) {
            ScanArgs a;
            a.rows = tiles;
            a.elt_bytes = filter_bf16 ? 2 : 4;
            a.half_tiles = s->half_tiles ? s->half_bits : 0;
            a.dim = s->dim;
            a.row_mul = row_mul;
            a.row_begin = r0;
            a.row_end = r1;
            a.q = pl.q_f32 + (size_t)b0 * s->dim;
            a.nq = (batch - b0 >= kScanMaxQueries) ? kScanMaxQueries : batch - b0;
            a.pool.scores = pool.scores + (size_t)b0 * kPoolCap;
            a.pool.ids = pool.ids + (size_t)b0 * kPoolCap;
            a.pool.cnt = pool.cnt + b0;
            a.pool.thr = pool.thr + b0;
            a.pool.ovf = pool.ovf + b0;
            a.dense = dense;
            a.wide_scores = (dense && wide_scores) ? wide_scores + (size_t)b0 * kWideDenseRows : nullptr;
            a.wide_ids = wide_ids ? wide_ids + (size_t)b0 * kWideDenseRows : nullptr;
            a.wide_stride = kWideDenseRows;
            a.sm_count = s->sm_count;
            int r = launch_scan(a, stream);
            if (r) return r;
        }
        return 0;
    };

    auto compact = [&](bool last) -> int {
        PhaseTimer t(1, stream);
        return launch_pool_compact(pool, batch, kprime, last ? 1 : 0, stream);
    };

    double growth = (double)(kPoolCap - kprime) / (3.0 * kprime);
    if (growth > 8.0) growth = 8.0;
    if (opt.slab_growth > 0 && opt.slab_growth < growth) growth = opt.slab_growth;
    if (growth < 1.0) growth = 1.0;
    // a slab of this many rows can never overflow a pool that holds at most kprime entries, whatever its threshold
    const int64_t safe_rows = (int64_t)((kPoolCap - kprime) & ~255);
    int64_t seen = 0;
    if (rows > 0) {
        if ((rc = run_filter(0, slab0, 1))) return rc;
        if (wide) {
            {
                PhaseTimer t(1, stream);
                // (the merge kernel also sorts and truncates when nothing else follows)
                if ((rc = launch_wide_select(seg, pool, batch, kprime, slab0 == rows ? 1 : 0, stream))) return rc;
            }
        } else if ((rc = compact(slab0 == rows))) {
            return rc;
        }
        seen = slab0;
        while (seen < rows) {
            int64_t m, end;
            // After a wide first slab the threshold is the kprime-th best of its 65536+ rows: the rest of the corpus
            // goes through ONE launch when its expected admissions, (live rows left) * kprime / (live rows seen),
            // plus the kprime survivors fill at most 85 % of the pool (up to ~1.05M rows at kprime = 224).  K2
            // scans the tiles in a stride permutation, so the slab was a representative sample and the estimate
            // holds whatever the row order; K1 scans in storage order and gets the single launch only up to 50 %.
            // A pool that overflows all the same is flagged and repaired like any other overflow.
            if (wide && seen == slab0 && live0 >= 2.0 * kprime && rest_fill <= (permuted ? 0.85 : 0.5) * kPoolCap) {
                if ((rc = run_filter(seen, rows, 0))) return rc;
                if ((rc = compact(true))) return rc;
                seen = rows;
                break;
            }
            const double ls = live_seen(seen);
            if ((mode & CMW_SLABS_SAFE) || ls < 2.0 * kprime) {
                // no (useful) threshold yet -- the slabs so far were mostly tombstones -- or the caller asked for
                // the overflow-proof schedule
                m = safe_rows;
                end = seen + m;
                if (end > rows) end = rows;
            } else {
                // expected admissions: m * kprime / ls = growth * kprime <= a third of the free pool
                m = (int64_t)(ls * growth);
                m &= ~(int64_t)255;
                if (m < kDenseSlabRows) m = kDenseSlabRows;
                end = seen + m;
                if (end > rows || rows - end < 4096) end = rows;
            }
            if ((rc = run_filter(seen, end, 0))) return rc;
            if ((rc = compact(end == rows))) return rc;
            seen = end;
        }
    } else {
        if ((rc = compact(true))) return rc;  // empty store: empty, consistent pools
    }
    return 0;
}

static CertParams make_cert(const SearchPlan& pl, const float* global_kth) {
    const Options& opt = pl.opt;
    CertParams cert;
    cert.kind = CERT_FIXED;
    // fp32 FMA filter: one lane's chain of D/32 FMAs + 5 shuffle adds + query scaling + row multiplier
    cert.eps_fixed = opt.f32_eps > 0 ? opt.f32_eps : (double)(pl.s->dim / 32 + 12) * 5.9604644775390625e-8;
    cert.sigmas = 0.0;
    cert.acc_slack = (double)pl.s->dim * 1.1920928955078125e-7 * 1.01;  // D * 2^-23
    cert.tile_u = pl.tf32 ? 1.0 / 1024.0 : (pl.s->half_tiles ? 1.0 / (double)(1 << pl.s->half_bits) : 1.0 / 256.0);
    cert.res_slot = pl.tf32 ? 3 : 2;
    if (pl.gemm) {
        if (opt.bf16_eps > 0) {
            cert.eps_fixed = opt.bf16_eps;
        } else if (opt.strict_certificate != 0) {
            cert.kind = CERT_RIGOROUS;
        } else {
            cert.kind = CERT_STATISTICAL;
            cert.sigmas = opt.bf16_sigmas;
        }
    }
    cert.q4 = pl.q4;
    cert.qres = pl.qres;
    cert.qn64 = pl.qn64;
    cert.norms = pl.s->maxnorm_bits;
    cert.metric = pl.metric;
    cert.global_kth = global_kth;
    return cert;
}

static int run_finish_half(const SearchPlan& pl, const float* queries_dev, const float* global_kth,
                           float* out_scores_dev, int64_t* out_ids_dev, double* out_scores64_dev,
                           int32_t* out_flags_dev, double* out_aux_dev, cudaStream_t fin_stream) {
    PhaseTimer tfin(2, fin_stream);
    if (pl.base_mode == CMW_MODE_F32_EXACT) {
        const CertParams cert = make_cert(pl, global_kth);
        return launch_rescore_select(pl.s, pl.pool, pl.batch, pl.k, pl.kprime, pl.metric, queries_dev, cert, pl.exact,
                                     out_scores_dev, out_ids_dev, out_scores64_dev, out_flags_dev, out_aux_dev,
                                     fin_stream);
    }
    return launch_pool_emit(pl.s, pl.pool, pl.batch, pl.k, pl.qn64, pl.metric, out_scores_dev, out_ids_dev,
                            out_scores64_dev, out_flags_dev, out_aux_dev, fin_stream);
}

// `fin_stream` == `stream`: everything in order on one stream (the public cmw_search).  Otherwise the finish
// half is forked onto `fin_stream` behind `fork_ev`, so that the caller may put the next search's filter on
// `stream` right away: the rescoring is an HBM gather that needs no shared memory or TMEM and runs next to the
// persistent tensor-core filter CTAs.
static int search_impl(cmw_store* h, const float* queries_dev, int batch, int k, int metric, int mode,
                       float* out_scores_dev, int64_t* out_ids_dev, double* out_scores64_dev,
                       int32_t* out_flags_dev, void* ws_dev, size_t ws_bytes, cudaStream_t stream,
                       cudaStream_t fin_stream, cudaEvent_t fork_ev) {
    CMW_REQUIRE(h != nullptr, "cmw_search: store is NULL");
    if (batch == 0) return 0;
    CMW_REQUIRE(out_scores_dev && out_ids_dev, "cmw_search: bad arguments");
    SearchPlan pl;
    int rc = make_plan(pl, h, queries_dev, batch, k, metric, mode, ws_dev, ws_bytes, "cmw_search");
    if (rc) return rc;
    if ((rc = run_filter_half(pl, queries_dev, stream))) return rc;
    if (fin_stream != stream) {
        CMW_CUDA_OK(cudaEventRecord(fork_ev, stream));
        CMW_CUDA_OK(cudaStreamWaitEvent(fin_stream, fork_ev, 0));
    }
    return run_finish_half(pl, queries_dev, nullptr, out_scores_dev, out_ids_dev, out_scores64_dev, out_flags_dev,
                           nullptr, fin_stream);
}

extern "C" {

int cmw_search(cmw_store* h, const float* queries_dev, int batch, int k, int metric, int mode,
               float* out_scores_dev, int64_t* out_ids_dev, double* out_scores64_dev,
               int32_t* out_flags_dev, void* ws_dev, size_t ws_bytes, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    return search_impl(h, queries_dev, batch, k, metric, mode, out_scores_dev, out_ids_dev, out_scores64_dev,
                       out_flags_dev, ws_dev, ws_bytes, stream, stream, nullptr);
}

// ---- row-sharded search: cmw_search in two halves around the exchange of the k-th filter score ----------
int cmw_search_filter(cmw_store* h, const float* queries_dev, int batch, int k, int metric, int mode,
                      float* out_filter_topk_dev, void* ws_dev, size_t ws_bytes, void* stream_v) {
    CMW_REQUIRE(h != nullptr, "cmw_search_filter: store is NULL");
    if (batch == 0) return 0;
    CMW_REQUIRE(out_filter_topk_dev != nullptr, "cmw_search_filter: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_v;
    SearchPlan pl;
    int rc = make_plan(pl, h, queries_dev, batch, k, metric, mode, ws_dev, ws_bytes, "cmw_search_filter");
    if (rc) return rc;
    if ((rc = run_filter_half(pl, queries_dev, stream))) return rc;
    PhaseTimer t(1, stream);
    return launch_pool_topk_scores(pl.pool, batch, k, out_filter_topk_dev, stream);
}

int cmw_search_finish(cmw_store* h, const float* queries_dev, int batch, int k, int metric, int mode,
                      const float* global_kth_dev, void* block_dev, void* ws_dev, size_t ws_bytes, void* stream_v) {
    CMW_REQUIRE(h != nullptr, "cmw_search_finish: store is NULL");
    if (batch == 0) return 0;
    CMW_REQUIRE(block_dev != nullptr && (reinterpret_cast<uintptr_t>(block_dev) & 15) == 0,
                "cmw_search_finish: block_dev must be a 16-byte aligned device pointer");
    SearchPlan pl;
    // the same arguments as the filter half -> the same workspace layout: the pools are where it left them
    int rc = make_plan(pl, h, queries_dev, batch, k, metric, mode, ws_dev, ws_bytes, "cmw_search_finish");
    if (rc) return rc;
    const ShardBlock lay = shard_block(batch, k);
    uint8_t* blk = reinterpret_cast<uint8_t*>(block_dev);
    return run_finish_half(pl, queries_dev, global_kth_dev, nullptr, reinterpret_cast<int64_t*>(blk + lay.ids),
                           reinterpret_cast<double*>(blk + lay.scores), reinterpret_cast<int32_t*>(blk + lay.flags),
                           reinterpret_cast<double*>(blk + lay.aux), (cudaStream_t)stream_v);
}

size_t cmw_shard_block_bytes(int batch, int k) {
    if (batch < 1 || k < 1) return 0;
    return shard_block(batch, k).total;
}

int cmw_shard_kth(const float* filter_topk_gathered_dev, int G, int B, int k, float* out_kth_dev, void* stream) {
    CMW_REQUIRE(filter_topk_gathered_dev && out_kth_dev, "cmw_shard_kth: NULL argument");
    CMW_REQUIRE(G >= 1 && B >= 0 && k >= 1, "cmw_shard_kth: bad sizes");
    if (B == 0) return 0;
    PhaseTimer t(4, (cudaStream_t)stream);
    return launch_shard_kth(filter_topk_gathered_dev, G, B, k, out_kth_dev, (cudaStream_t)stream);
}

int cmw_shard_merge(const void* blocks_dev, int G, int B, int k, int k_out, float* out_scores_dev,
                    int64_t* out_ids_dev, double* out_scores64_dev, int32_t* out_flags_dev, void* stream) {
    return cmw_shard_merge_ex(blocks_dev, G, B, k, k_out, out_scores_dev, out_ids_dev, out_scores64_dev, out_flags_dev,
                              nullptr, stream);
}

int cmw_shard_merge_ex(const void* blocks_dev, int G, int B, int k, int k_out, float* out_scores_dev,
                       int64_t* out_ids_dev, double* out_scores64_dev, int32_t* out_flags_dev,
                       const int32_t* peer_status_dev, void* stream) {
    CMW_REQUIRE(blocks_dev && out_scores_dev && out_ids_dev, "cmw_shard_merge: NULL argument");
    CMW_REQUIRE(G >= 1 && B >= 0 && k >= 1 && k_out >= 1, "cmw_shard_merge: bad sizes");
    if (B == 0) return 0;
    PhaseTimer t(5, (cudaStream_t)stream);
    return launch_shard_merge(blocks_dev, G, B, k, k_out, out_scores_dev, out_ids_dev, out_scores64_dev,
                              out_flags_dev, peer_status_dev, (cudaStream_t)stream);
}

int cmw_search_host(cmw_store* h, const float* queries_host, int batch, int k, int metric, int mode,
                    float* out_scores_host, int64_t* out_ids_host, int32_t* out_flags_host) {
    CMW_REQUIRE(h != nullptr, "cmw_search_host: store is NULL");
    Store* s = reinterpret_cast<Store*>(h);
    if (batch == 0) return 0;
    CMW_REQUIRE(batch > 0 && queries_host && out_scores_host && out_ids_host,
                "cmw_search_host: bad arguments");
    CMW_REQUIRE(k >= 1 && k <= kMaxKPrime, "cmw_search_host: k must be in [1, %d], got %d", kMaxKPrime, k);
    std::lock_guard<std::mutex> host_lock(s->host_mu);
    CMW_CUDA_OK(cudaSetDevice(s->device));
    cudaStream_t stream;
    int rc = get_stream(s, &stream);
    if (rc) return rc;
    const HostIo io = host_io(s, batch, k);
    if ((rc = ensure_blocking_buffers(h, s, io, batch, k, mode))) return rc;
    const bool out_pinned = is_pinned(out_scores_host) && is_pinned(out_ids_host);
    if ((rc = run_blocking(h, s, stream, io, queries_host, batch, k, metric, mode,
                           out_pinned ? out_scores_host : nullptr, out_pinned ? out_ids_host : nullptr)))
        return rc;
    uint8_t* pin = reinterpret_cast<uint8_t*>(s->pinned);
    if (!out_pinned) {
        memcpy(out_scores_host, pin + io.q_bytes, (size_t)batch * k * sizeof(float));
        memcpy(out_ids_host, pin + io.q_bytes + io.sc_bytes, (size_t)batch * k * sizeof(int64_t));
    }
    return repair_flagged(h, s, stream, io, queries_host, batch, k, metric, mode, out_scores_host, out_ids_host,
                          out_flags_host, reinterpret_cast<const int32_t*>(pin + io.q_bytes + io.sc_bytes + io.id_bytes));
}

int cmw_search_host_submit(cmw_store* h, const float* queries_host, int batch, int k, int metric, int mode,
                           float* out_scores_host, int64_t* out_ids_host, int32_t* out_flags_host,
                           int* ticket_out) {
    CMW_REQUIRE(h != nullptr, "cmw_search_host_submit: store is NULL");
    Store* s = reinterpret_cast<Store*>(h);
    CMW_REQUIRE(ticket_out != nullptr, "cmw_search_host_submit: ticket_out is NULL");
    *ticket_out = -1;
    CMW_REQUIRE(batch > 0 && queries_host && out_scores_host && out_ids_host,
                "cmw_search_host_submit: bad arguments");
    CMW_REQUIRE(k >= 1 && k <= kMaxKPrime, "cmw_search_host_submit: k must be in [1, %d], got %d", kMaxKPrime, k);
    std::lock_guard<std::mutex> host_lock(s->host_mu);
    CMW_CUDA_OK(cudaSetDevice(s->device));
    int slot_no = -1;
    for (int i = 0; i < kHostSlots; ++i)
        if (!s->slots[i].busy) {
            slot_no = i;
            break;
        }
    if (slot_no < 0) {
        set_error("cmw_search_host_submit: all %d slots are in flight; cmw_search_host_wait one first", kHostSlots);
        return -4;
    }
    HostSlot& sl = s->slots[slot_no];
    cudaStream_t compute;
    int rc = get_stream(s, &compute);
    if (rc) return rc;
    if (s->copy_in == nullptr) CMW_CUDA_OK(cudaStreamCreateWithFlags(&s->copy_in, cudaStreamNonBlocking));
    if (s->copy_out == nullptr) CMW_CUDA_OK(cudaStreamCreateWithFlags(&s->copy_out, cudaStreamNonBlocking));
    if (s->tail == nullptr) CMW_CUDA_OK(cudaStreamCreateWithFlags(&s->tail, cudaStreamNonBlocking));
    if (sl.ev_fork == nullptr) CMW_CUDA_OK(cudaEventCreateWithFlags(&sl.ev_fork, cudaEventDisableTiming));
    const bool overlap_tail = g_opt.host_overlap != 0;
    cudaStream_t fin = overlap_tail ? s->tail : compute;
    if (sl.ev_in == nullptr) CMW_CUDA_OK(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
    if (sl.ev_compute == nullptr) CMW_CUDA_OK(cudaEventCreateWithFlags(&sl.ev_compute, cudaEventDisableTiming));
    if (sl.ev_done == nullptr) CMW_CUDA_OK(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    const HostIo io = host_io(s, batch, k);
    const size_t ws_bytes = cmw_search_workspace_bytes(h, batch, k, mode);
    if ((rc = ensure_buffer(&sl.dev_io, &sl.dev_io_bytes, io.total, false))) return rc;
    if ((rc = ensure_buffer(&sl.ws, &sl.ws_bytes, ws_bytes, false))) return rc;
    const bool in_pinned = is_pinned(queries_host);
    const bool out_pinned = is_pinned(out_scores_host) && is_pinned(out_ids_host);
    if ((rc = ensure_buffer(&sl.pinned, &sl.pinned_bytes, io.total, true))) return rc;
    uint8_t* pin = reinterpret_cast<uint8_t*>(sl.pinned);
    uint8_t* dev = reinterpret_cast<uint8_t*>(sl.dev_io);

    auto enqueue = [&]() -> int {
        // copy-in stream: H2D of the queries (overlaps whatever the compute stream is doing for earlier tickets)
        const size_t qb = (size_t)batch * s->dim * sizeof(float);
        if (in_pinned) {
            CMW_CUDA_OK(cudaMemcpyAsync(dev, queries_host, qb, cudaMemcpyHostToDevice, s->copy_in));
        } else if (stage_h2d(s->device, pin, dev, reinterpret_cast<const uint8_t*>(queries_host), qb, s->copy_in)) {
            return -2;
        }
        CMW_CUDA_OK(cudaEventRecord(sl.ev_in, s->copy_in));
        CMW_CUDA_OK(cudaStreamWaitEvent(compute, sl.ev_in, 0));
        // filter phases on the compute stream (in order over all tickets and blocking calls of this store);
        // the finalisation on the tail stream, where it overlaps the next ticket's filter
        rc = search_impl(h, reinterpret_cast<const float*>(dev), batch, k, metric, mode,
                         reinterpret_cast<float*>(dev + io.q_bytes),
                         reinterpret_cast<int64_t*>(dev + io.q_bytes + io.sc_bytes), nullptr,
                         reinterpret_cast<int32_t*>(dev + io.q_bytes + io.sc_bytes + io.id_bytes), sl.ws, sl.ws_bytes,
                         compute, fin, sl.ev_fork);
        if (rc) return rc;
        CMW_CUDA_OK(cudaEventRecord(sl.ev_compute, fin));
        // copy-out stream: D2H of the results
        CMW_CUDA_OK(cudaStreamWaitEvent(s->copy_out, sl.ev_compute, 0));
        if (out_pinned) {
            CMW_CUDA_OK(cudaMemcpyAsync(out_scores_host, dev + io.q_bytes, (size_t)batch * k * sizeof(float),
                                        cudaMemcpyDeviceToHost, s->copy_out));
            CMW_CUDA_OK(cudaMemcpyAsync(out_ids_host, dev + io.q_bytes + io.sc_bytes, (size_t)batch * k * sizeof(int64_t),
                                        cudaMemcpyDeviceToHost, s->copy_out));
            CMW_CUDA_OK(cudaMemcpyAsync(pin + io.q_bytes + io.sc_bytes + io.id_bytes,
                                        dev + io.q_bytes + io.sc_bytes + io.id_bytes, io.fl_bytes, cudaMemcpyDeviceToHost,
                                        s->copy_out));
        } else {
            CMW_CUDA_OK(cudaMemcpyAsync(pin + io.q_bytes, dev + io.q_bytes, io.sc_bytes + io.id_bytes + io.fl_bytes,
                                        cudaMemcpyDeviceToHost, s->copy_out));
        }
        CMW_CUDA_OK(cudaEventRecord(sl.ev_done, s->copy_out));
        return 0;
    };
    if ((rc = enqueue())) {
        // part of the chain may already be queued on buffers this slot will hand out again: drain it
        cudaStreamSynchronize(s->copy_in);
        cudaStreamSynchronize(compute);
        cudaStreamSynchronize(s->tail);
        cudaStreamSynchronize(s->copy_out);
        return rc;
    }
    sl.q_host = queries_host;
    sl.batch = batch;
    sl.k = k;
    sl.metric = metric;
    sl.mode = mode;
    sl.out_scores = out_scores_host;
    sl.out_ids = out_ids_host;
    sl.out_flags = out_flags_host;
    sl.out_pinned = out_pinned;
    sl.busy = true;
    *ticket_out = slot_no;
    return 0;
}

int cmw_search_host_wait(cmw_store* h, int ticket) {
    CMW_REQUIRE(h != nullptr, "cmw_search_host_wait: store is NULL");
    Store* s = reinterpret_cast<Store*>(h);
    CMW_REQUIRE(ticket >= 0 && ticket < kHostSlots, "cmw_search_host_wait: bad ticket %d", ticket);
    HostSlot& sl = s->slots[ticket];
    cudaEvent_t done;
    {
        std::lock_guard<std::mutex> host_lock(s->host_mu);
        CMW_REQUIRE(sl.busy, "cmw_search_host_wait: ticket %d is not in flight", ticket);
        done = sl.ev_done;
    }
    // wait outside the lock: other threads may submit (or wait for other tickets) meanwhile
    CMW_CUDA_OK(cudaSetDevice(s->device));
    cudaError_t e = cudaEventSynchronize(done);
    std::lock_guard<std::mutex> host_lock(s->host_mu);
    sl.busy = false;
    if (e != cudaSuccess) {
        set_error("cmw_search_host_wait: %s", cudaGetErrorString(e));
        return -2;
    }
    const HostIo io = host_io(s, sl.batch, sl.k);
    const uint8_t* pin = reinterpret_cast<const uint8_t*>(sl.pinned);
    if (!sl.out_pinned) {
        memcpy(sl.out_scores, pin + io.q_bytes, (size_t)sl.batch * sl.k * sizeof(float));
        memcpy(sl.out_ids, pin + io.q_bytes + io.sc_bytes, (size_t)sl.batch * sl.k * sizeof(int64_t));
    }
    const int32_t* fl = reinterpret_cast<const int32_t*>(pin + io.q_bytes + io.sc_bytes + io.id_bytes);
    bool any = false;
    for (int b = 0; b < sl.batch && !any; ++b) any = (fl[b] & CMW_FLAG_UNCERTIFIED) != 0;
    if (!any || g_opt.repair < 1) {
        if (sl.out_flags) memcpy(sl.out_flags, fl, (size_t)sl.batch * sizeof(int32_t));
        return 0;
    }
    // rare: flagged queries go through the blocking repair chain on the compute stream (in order behind
    // whatever has been submitted since)
    cudaStream_t stream;
    int rc = get_stream(s, &stream);
    if (rc) return rc;
    if ((rc = ensure_blocking_buffers(h, s, io, sl.batch, sl.k, sl.mode))) return rc;
    return repair_flagged(h, s, stream, io, sl.q_host, sl.batch, sl.k, sl.metric, sl.mode, sl.out_scores, sl.out_ids,
                          sl.out_flags, fl);
}

}  // extern "C"
