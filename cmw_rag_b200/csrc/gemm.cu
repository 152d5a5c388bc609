// gemm.cu -- K2: tcgen05 / TMEM tensor-core filter with the top-k admission fused into the epilogue.
//
// Stands in for collection.query() (rag_engine/storage/vector_store.py:59-63 of the reference) for
// batches of query vectors: the S per-segment searches RAGRetriever gathers at
// rag_engine/retrieval/retriever.py:179-182, or thousands of independent queries.
//
// Shape: scores[row, query] = sum_d corpus_bf16[row, d] * query_bf16[query, d]
//   A operand = corpus tile, 128 rows (the MMA M dimension -> the 128 TMEM lanes)
//   B operand = query group, NT <= 256 queries (the MMA N dimension -> TMEM columns)
//   both K-major bf16 in shared memory, 128-byte swizzle, written by TMA (cp.async.bulk.tensor.2d);
//   fp32 accumulators in TMEM, two stages of 256 columns so the epilogue of item i overlaps the
//   MMAs of item i+1.
// Warp roles (192 threads, or 320 with 8 epilogue warps for 192/256-column query groups; one persistent CTA
// per SM):
//   warp 0   TMA producer (one elected lane): ring of [A 16 KB | B NT*128 B] stages, mbarrier tx
//   warp 1   TMEM allocator + MMA issuer: the whole warp walks the pipeline warp-uniformly (operand
//            descriptors stay in uniform registers), one elect.sync lane issues 4 back-to-back
//            tcgen05.mma.cta_group::1.kind::f16 (K = 16) per k-block; tcgen05.commit releases
//            shared-memory stages and publishes finished accumulators
//   warps 2-5 (2-9) epilogue (gemm_common.cuh): tcgen05.ld (32 lanes x 32 columns), multiply by the per-row
//            multiplier (NaN for tombstoned rows), compare with the per-query admission thresholds
//            (fetched as one batch while the TMEM load is in flight), stage the rare survivors in
//            shared memory (ballot + popc), hand the accumulator back, then flush the staged entries
//            to the queries' candidate pools with 32-64 global atomics in flight.  In the first
//            ("dense") slab every score is written to its own slot instead.
// The B x N score matrix is never written to memory.
//
// Work items: (row tile, query group), row-tile-major so that the CTAs running concurrently read the
// same corpus rows for different query groups (one HBM read, L2 hits for the rest).
// Algorithmic work per item: 2 * 128 * NT * D flop; HBM bytes per row tile: 128 * D * 2.
#include "gemm_common.cuh"

namespace cmw {

constexpr int kGemmThreads = 192;      // TMA warp, MMA warp, 4 epilogue warps
constexpr int kGemmThreadsWide = 320;  // ... 8 epilogue warps (query groups of 192 / 256 columns)

__global__ void __launch_bounds__(kGemmThreadsWide, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;

    // 128-byte-swizzled operand tiles must start on 1024-byte boundaries of the shared window
    uint8_t* stages = smem + ((1024u - (ptx::smem_u32(smem) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)p.nstages * p.stage_bytes);
    uint64_t* full = bars;                       // [kMaxStages]
    uint64_t* empty = bars + kMaxStages;         // [kMaxStages]
    uint64_t* tmem_full = bars + 2 * kMaxStages;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint2* stage_buf = reinterpret_cast<uint2*>(tmem_empty + 4);  // [4][kStageCap] or [8][kStageCap2] survivors

    const int n_items = p.n_tiles * p.n_groups;
    const int n_epi = (int)(blockDim.x >> 5) - 2;  // 4, or 8 = two warps per TMEM lane quarter
    int item_begin, item_end, item_step;
    item_range(n_items, p.n_groups, (int)blockIdx.x, (int)gridDim.x, item_begin, item_end, item_step);

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        for (int s = 0; s < p.nstages; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], (uint32_t)n_epi);  // one arrive per epilogue warp
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_base_smem, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    // programmatic dependent launch (small batches): everything above -- descriptor prefetch, barrier init, TMEM
    // allocation -- may run while the previous kernel of the search is still finishing; the query tile, the
    // admission thresholds and the pools are its outputs
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint64_t pol_a = (p.n_groups > 1) ? ptx::l2_policy_evict_last() : ptx::l2_policy_evict_first();
            const uint64_t pol_b = ptx::l2_policy_evict_last();
            const uint32_t tx_bytes = (uint32_t)(kABytes + p.nt * kBlockK * 2);
            int stage = 0;
            uint32_t phase = 0;
            for (int item = item_begin; item != item_end; item += item_step) {
                const int tile = item / p.n_groups;
                const int group = item - tile * p.n_groups;
                const int row0 = (int)(scan_tile(p, tile) * kTileM);
                const int q0 = group * p.nt;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1u);
                    uint8_t* sa = stages + (size_t)stage * p.stage_bytes;
                    uint8_t* sb = sa + kABytes;
                    ptx::mbar_arrive_expect_tx(&full[stage], tx_bytes);
                    ptx::tma_load_2d(sa, &tmap_a, kb * p.kb_elems, row0, &full[stage], pol_a);
                    ptx::tma_load_2d(sb, &tmap_b, kb * p.kb_elems, q0, &full[stage], pol_b);
                    if (++stage == p.nstages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp walks the pipeline (uniform control flow, operands in uniform registers); one
        // elected lane issues the MMAs and the commits.
        const uint32_t stage0_lo = (ptx::smem_u32(stages) >> 4);
        const uint32_t stage_step = (uint32_t)p.stage_bytes >> 4;
        constexpr uint32_t kDescHi = (uint32_t)(1024u >> 4) | (1u << 14) | (2u << 29);  // SBO, version, SW128
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int item = item_begin; item != item_end; item += item_step, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccStride);
            for (int kb = 0; kb < p.num_kb; ++kb) {
                ptx::mbar_wait(&full[stage], phase);
                ptx::tc_fence_after();
                const uint32_t a_lo = (stage0_lo + (uint32_t)stage * stage_step) & 0x3fffu;
                const uint32_t b_lo = (stage0_lo + (uint32_t)stage * stage_step + (kABytes >> 4)) & 0x3fffu;
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        // 32 bytes of every operand row per instruction in either format: 16 x 16-bit or 8 x tf32
                        const uint64_t da = ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + 2u * k);
                        const uint64_t db = ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + 2u * k);
                        if (p.tf32) ptx::umma_tf32(d_tmem, da, db, p.idesc, (kb | k) != 0 ? 1u : 0u);
                        else ptx::umma_bf16(d_tmem, da, db, p.idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty[stage]);  // frees the stage once these MMAs have read it
                    if (kb == p.num_kb - 1) ptx::umma_commit(&tmem_full[acc]);  // accumulator complete
                }
                __syncwarp();
                if (++stage == p.nstages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5, or 2..9) =====================
        // TMEM lanes [32*quarter, 32*quarter + 32) belong to this warp.  With 8 epilogue warps, two share a
        // quarter and split the query columns (see gemm2.cu: one epilogue warp per scheduler runs at ~0.2 IPC).
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int nt_local = (n_epi == 8) ? (p.nt >> 1) : p.nt;
        const int stage_cap = (n_epi == 8) ? kStageCap2 : kStageCap;
        uint2* stg = stage_buf + (size_t)(warp - 2) * stage_cap;
        int it = 0;
        for (int item = item_begin; item != item_end; item += item_step, ++it) {
            const int tile = item / p.n_groups;
            const int group = item - tile * p.n_groups;
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            const int64_t row_warp0 = scan_tile(p, tile) * kTileM + quarter * 32;
            const int64_t dense_slot0 = (int64_t)tile * kTileM + quarter * 32;  // position within the slab
            const int q0 = group * p.nt + half * nt_local;
            int ncols = p.batch - q0;  // real (unpadded) queries in this warp's columns
            if (ncols > nt_local) ncols = nt_local;
            if (ncols < 0) ncols = 0;
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                   (uint32_t)(acc * kAccStride + half * nt_local);
            uint64_t* rel = &tmem_empty[acc];
            epilogue_item(p, taddr, row_warp0, dense_slot0, lane, q0, ncols, nt_local, stg, stage_cap,
                          [rel, lane]() { if (lane == 0) ptx::mbar_arrive(rel); });
        }
    }

    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 2-D row-major [rows, dim] tensor of bf16 / fp16 / fp32 elements, box = 128 bytes x box_rows, 128-byte swizzle
static int encode_2d_typed(CUtensorMap* out, const void* base, int64_t rows, int dim, int box_rows,
                           CUtensorMapDataType dtype, int elt_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    CMW_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)dim * (cuuint64_t)elt_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(128 / elt_bytes), (cuuint32_t)box_rows};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = fn(out, dtype, 2,
                    const_cast<void*>(base), gdim, gstride, box,
                    estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CMW_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

int encode_2d(CUtensorMap* out, const void* base, int64_t rows, int dim, int box_rows, bool half_tiles) {
    return encode_2d_typed(out, base, rows, dim, box_rows,
                           half_tiles ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
}

int encode_bf16_tmap(Store* s) {
    if (s->bf16 == nullptr) return -1;
    return encode_2d(&s->tmap_bf16, s->bf16, s->capacity, s->dim, kTileM, s->half_tiles);
}

int encode_f32_tmap(Store* s) {
    if (s->f32 == nullptr) return -1;
    return encode_2d_typed(&s->tmap_f32, s->f32, s->capacity, s->dim, kTileM, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4);
}

bool gemm_supported(const Store* s) { return s->bf16 != nullptr && s->tmap_ok; }
bool gemm_tf32_supported(const Store* s) { return s->f32 != nullptr && s->tmap_f32_ok; }

int launch_gemm_2cta(const GemmArgs& a, cudaStream_t stream);  // gemm2.cu

// Stride permutation of the store's tiles (GemmParams::perm_mul): multiplier ~ tiles / golden ratio, coprime to
// the tile count.  Off for a launch that covers the whole store (nothing to be representative of) and for tiny
// stores.
void set_scan_order(GemmParams& p, const GemmArgs& a, int tile_rows) {
    const Store* s = a.store;
    p.tile_begin = a.row_begin / tile_rows;
    p.perm_mul = 0;
    p.perm_tiles = (s->rows + tile_rows - 1) / tile_rows;
    const bool whole = a.row_begin == 0 && a.row_end >= s->rows;
    if (a.opt->scan_permute == 0 || whole || p.perm_tiles < 64) return;
    auto gcd = [](int64_t x, int64_t y) {
        while (y) {
            const int64_t t = x % y;
            x = y;
            y = t;
        }
        return x;
    };
    int64_t m = (int64_t)((double)p.perm_tiles * 0.6180339887498949) | 1;
    while (gcd(m, p.perm_tiles) != 1) m += 2;
    p.perm_mul = m % p.perm_tiles;
    // rows are validated against the end of the STORE: a permuted tile can lie anywhere in it
    p.row_end = s->rows;
}

// which K2 kernel a padded batch runs on: CTA pairs (cta_group::2) for the tensor-bound batches, single CTAs
// for the HBM-bound ones
static bool use_cta_pairs(const Options& o, int bpad) {
    return o.gemm_2cta != 0 && bpad >= (int)o.gemm_2cta_min_batch && gemm_group_width(bpad) % 64 == 0;
}

// true when the slabs of a multi-slab search scan the store's tiles in the stride permutation (set_scan_order),
// i.e. every slab is a spread sample of the corpus; false = storage order
bool gemm_scan_permuted(const Store* s, const Options& o, int bpad, bool tf32) {
    const int tile_rows = (!tf32 && use_cta_pairs(o, bpad)) ? 2 * kTileM : kTileM;
    return o.scan_permute != 0 && (s->rows + tile_rows - 1) / tile_rows >= 64;
}

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
    const Store* s = a.store;
    CMW_REQUIRE(a.tf32 ? gemm_tf32_supported(s) : gemm_supported(s),
                "launch_gemm: the store lacks the tiles / TMA descriptor this filter reads");
    if (a.row_end <= a.row_begin) return 0;
    if (a.dense)
        CMW_REQUIRE(a.row_end - a.row_begin <= (a.wide_scores ? a.wide_stride : kPoolCap),
                    "launch_gemm: dense slab larger than its destination");
    CMW_REQUIRE((a.row_begin % (2 * kTileM)) == 0, "launch_gemm: slab start must be a multiple of %d rows", 2 * kTileM);
    if (!a.tf32 && use_cta_pairs(*a.opt, a.bpad)) return launch_gemm_2cta(a, stream);
    GemmParams p;
    p.dim = s->dim;
    p.tf32 = a.tf32;
    p.kb_elems = a.tf32 ? kBlockK / 2 : kBlockK;
    p.num_kb = (s->dim + p.kb_elems - 1) / p.kb_elems;
    p.nt = gemm_group_width(a.bpad);
    CMW_REQUIRE(p.nt % 16 == 0 && a.bpad % p.nt == 0, "launch_gemm: bad query padding %d", a.bpad);
    p.n_groups = a.bpad / p.nt;
    p.batch = a.batch;
    p.row_begin = a.row_begin;
    p.row_end = a.row_end;
    p.n_tiles = (int)((a.row_end - a.row_begin + kTileM - 1) / kTileM);
    CMW_REQUIRE(a.row_begin % kTileM == 0, "launch_gemm: slab start must be a multiple of %d rows", kTileM);
    set_scan_order(p, a, kTileM);
    p.stage_bytes = kABytes + p.nt * kBlockK * 2;
    const size_t tail = (2 * kMaxStages + 4) * sizeof(uint64_t) + 64 + 4 * kStageCap * sizeof(uint2);
    int nst = (int)((220 * 1024 - tail - 1024) / (size_t)p.stage_bytes);
    if (nst > kMaxStages) nst = kMaxStages;
    p.nstages = nst;
    p.dense = a.dense;
    p.dense_scores = a.wide_scores ? a.wide_scores : a.pool.scores;
    p.dense_ids = a.wide_scores ? a.wide_ids : a.pool.ids;
    p.dense_stride = a.wide_scores ? a.wide_stride : kPoolCap;
    p.dynamic = 0;
    // instruction descriptor: D = f32 (bit 4), A / B format at bits 7 / 10 (0 = fp16, 1 = bf16), both K-major,
    // N >> 3 at bit 17, M >> 4 at bit 24
    // (2 = tf32: the fp32 rows themselves, narrowed by the tensor core)
    const uint32_t fmt = a.tf32 ? 2u : (s->half_tiles ? 0u : 1u);
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.nt >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    p.row_mul = a.row_mul;
    p.pool_scores = a.pool.scores;
    p.pool_ids = a.pool.ids;
    p.pool_cnt = a.pool.cnt;
    p.pool_thr = a.pool.thr;
    CUtensorMap tmap_b;
    int rc = a.tf32 ? encode_2d_typed(&tmap_b, a.q_tf32, a.bpad, s->dim, p.nt, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4)
                    : encode_2d(&tmap_b, a.q_bf16, a.bpad, s->dim, p.nt, s->half_tiles);
    if (rc) return rc;
    const size_t smem = (size_t)nst * p.stage_bytes + tail + 1024;  // + slack for 1024-byte alignment
    static SmemAttrCache smem_set;
    if (smem_set.needs(smem)) {
        CMW_CUDA_OK(cudaFuncSetAttribute(gemm_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        smem_set.done(smem);
    }
    const int n_items = p.n_tiles * p.n_groups;
    const int grid = n_items < s->sm_count ? n_items : s->sm_count;
    // 8 epilogue warps for the widest query groups (192 or 256 columns): there the 4-warp epilogue of a tile takes as long as streaming its rows from HBM; narrower groups measured no gain
    const int threads = (p.nt >= 192 && p.nt % 64 == 0) ? kGemmThreadsWide : kGemmThreads;
    CMW_CUDA_OK(launch_kernel(gemm_topk_kernel, dim3(grid), dim3(threads), smem, stream, a.batch <= kWideDenseMaxBatch,
                              a.tf32 ? s->tmap_f32 : s->tmap_bf16, tmap_b, p));
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cmw
